"""mt19937_oracle.py -- TEST INFRASTRUCTURE.  numpy restatement of how torch's CPU default generator produces
``torch.rand(n)`` (float32): at::mt19937 (ATen/core/MT19937RNGEngine.h: next_state / operator()) and
at::uniform_real_distribution<float> ((r & (2^24 - 1)) * 2^-24, ATen/core/TransformationHelper.h), plus the layout of the
5056-byte state tensor of torch.get_rng_state() (CPUGeneratorImplStateLegacy: seed u64 @0, left i32 @8, seeded i32 @12,
next u64 @16, state u64[624] @24).  The reference draws its Cranley-Patterson shift with exactly that call
(utils/sample_util.py:102: torch.rand(b, 1, 2)).  Pinned against torch itself in tests/test_oracle.py; the device kernel
texir_mt19937_uniform is checked against torch.rand on the GPU box."""
import numpy as np

N, M = 624, 397
OFF_LEFT, OFF_NEXT, OFF_STATE = 8, 16, 24


def parse_state(state_bytes: np.ndarray):
    a = np.asarray(state_bytes, np.uint8)
    left = int(a[OFF_LEFT:OFF_LEFT + 4].view(np.int32)[0])
    nxt = int(a[OFF_NEXT:OFF_NEXT + 8].view(np.uint64)[0])
    words = a[OFF_STATE:OFF_STATE + 8 * N].view(np.uint64).astype(np.uint32)
    return words, left, nxt


def store_state(state_bytes: np.ndarray, words: np.ndarray, left: int, nxt: int) -> np.ndarray:
    a = np.array(state_bytes, np.uint8, copy=True)
    a[OFF_LEFT:OFF_LEFT + 4] = np.array([left], np.int32).view(np.uint8)
    a[OFF_NEXT:OFF_NEXT + 8] = np.array([nxt], np.uint64).view(np.uint8)
    a[OFF_STATE:OFF_STATE + 8 * N] = words.astype(np.uint64).view(np.uint8)
    return a


def _twist(u, v):
    return (((u & np.uint32(0x80000000)) | (v & np.uint32(0x7FFFFFFF))) >> np.uint32(1)) ^ np.where(
        v & np.uint32(1), np.uint32(0x9908B0DF), np.uint32(0))


def next_state(s: np.ndarray) -> np.ndarray:
    old = s.copy()
    s = s.copy()
    s[0:N - M] = old[M:N] ^ _twist(old[0:N - M], old[1:N - M + 1])
    s[N - M:2 * (N - M)] = s[0:N - M] ^ _twist(old[N - M:2 * (N - M)], old[N - M + 1:2 * (N - M) + 1])
    s[2 * (N - M):N - 1] = s[N - M:M - 1] ^ _twist(old[2 * (N - M):N - 1], old[2 * (N - M) + 1:N])
    s[N - 1] = s[M - 1] ^ _twist(old[N - 1:N], s[0:1])[0]
    return s


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


def rand_float32(words: np.ndarray, left: int, nxt: int, n: int):
    """-> (values float32 [n], new words, new left, new next): n calls of the engine's operator()"""
    s = np.asarray(words, np.uint32)
    out = []
    k = min(n, max(left - 1, 0))
    out.append(s[nxt:nxt + k])
    left, nxt, n = left - k, nxt + k, n - k
    while n > 0:
        s = next_state(s)
        m = min(n, N)
        out.append(s[:m])
        n, left, nxt = n - m, N + 1 - m, m
    y = temper(np.concatenate(out))
    return (y & np.uint32(0xFFFFFF)).astype(np.float32) * np.float32(2.0 ** -24), s, left, nxt
