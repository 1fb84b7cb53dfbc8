// texir_rng.cu -- the reference's Cranley-Patterson shift stream, regenerated on the device.
//
// The reference draws the shift of every generate_dir call as  torch.rand(b, 1, 2)  on the CPU default generator and
// moves it to the GPU (utils/sample_util.py:102).  At the Mat stage's 98 304 pixels that is 196 608 serial mt19937 draws,
// 0.3 ... 0.7 ms of host time per optimisation step -- more than the host spends on everything else of a 16-spp step.
// k_mt19937_uniform produces THE SAME float32 numbers from the same generator state: at::mt19937 (624-word state;
// next_state() = three data-parallel phases of 227 + 227 + 170 words; tempering) and
// at::uniform_real_distribution<float> ((r & (2^24 - 1)) * 2^-24), one CTA, the state in shared memory; it also
// returns the advanced state so that the host generator can be set to where torch.rand would have left it.
// Restated (and pinned against torch itself) in oracle/mt19937_oracle.py.
#include <cuda_runtime.h>

#include <string>

#include "../../include/texir_b200.h"
#include "texir_host.h"

using texir_host::fail;
using texir_host::finish_launch;

namespace {
constexpr int kN = 624, kM = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7FFFFFFFu)) >> 1) ^ ((v & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ float mt_uniform(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return (float)(y & 0xFFFFFFu) * 5.9604644775390625e-8f;  // 2^-24
}

// Two copies of the state: a twist reads block b from one and writes block b + 1 into the other, so a phase needs no
// barrier between its reads and its writes (three barriers per 624 numbers instead of seven).  Warps 0-7 carry the
// twist -- a chain of 3 dependent phases per block, 945 for a cube map's 196 608 numbers: the latency of that chain IS
// the kernel -- and warps 8-15 temper and store block b + 1 while the first phase of the next twist runs (both only
// read it).
constexpr int kThreads = 512, kTwist = 256;
__global__ void __launch_bounds__(kThreads) k_mt19937_uniform(const uint32_t* __restrict__ state_in, int left, int next, int64_t n,
                                                              float* __restrict__ out, uint32_t* __restrict__ state_out) {
    __shared__ uint32_t buf[2][kN];
    constexpr int kA = kN - kM;  // 227 words per phase
    const int t = threadIdx.x;
    const bool twister = t < kTwist;
    for (int i = t; i < kN; i += kThreads) buf[0][i] = state_in[i];
    if (left < 0) {
        // CUDA-graph form: the position inside the block travels with the state (words 624, 625), because kernel
        // arguments are frozen at capture; an inconsistent pair produces zeros and marks the returned state
        left = (int)state_in[kN];
        next = (int)state_in[kN + 1];
        if (left < 0 || left > kN + 1 || next < 0 || next > kN || (left > 1 && next + left - 1 > kN)) {
            for (int64_t i = t; i < n; i += kThreads) out[i] = 0.f;
            for (int i = t; i < kN; i += kThreads) state_out[i] = buf[0][i];
            if (t == 0) state_out[kN] = state_out[kN + 1] = 0xFFFFFFFFu;
            return;
        }
    }
    __syncthreads();
    int64_t produced = 0;
    {  // what is left of the current block of 624 words
        const int avail = left > 1 ? left - 1 : 0;
        const int k = (int)(n < avail ? n : avail);
        for (int i = t; i < k; i += kThreads) out[i] = mt_uniform(buf[0][next + i]);
        produced = k;
        left -= k;
        next += k;
    }
    int cur = 0;
    int64_t pend_at = 0;  // block waiting for its tempering + store: `pend_m` numbers of buf[cur] to out + pend_at
    int pend_m = 0;
    while (produced < n) {
        const uint32_t* o = buf[cur];
        uint32_t* w = buf[cur ^ 1];
        // next_state(): new[i] = x[(i + 397) mod 624] ^ twist(old[i], old[i + 1]), x = new once i + 397 wraps
        if (twister) {
            if (t < kA) w[t] = o[t + kM] ^ mt_twist(o[t], o[t + 1]);                        // i in [0, 227): old words only
        } else {
            for (int i = t - kTwist; i < pend_m; i += kThreads - kTwist) out[pend_at + i] = mt_uniform(o[i]);  // previous block
        }
        __syncthreads();
        if (t < kA) w[t + kA] = w[t] ^ mt_twist(o[t + kA], o[t + kA + 1]);                  // i in [227, 454)
        __syncthreads();
        if (t < kN - 2 * kA - 1) w[t + 2 * kA] = w[t + kA] ^ mt_twist(o[t + 2 * kA], o[t + 2 * kA + 1]);   // i in [454, 623)
        if (t == 255) w[kN - 1] = w[kM - 1] ^ mt_twist(o[kN - 1], w[0]);                    // i = 623 (wraps onto new[0])
        __syncthreads();
        cur ^= 1;
        pend_m = (int)(n - produced < kN ? n - produced : kN);
        pend_at = produced;
        produced += pend_m;
        left = kN + 1 - pend_m;
        next = pend_m;
    }
    for (int i = t; i < pend_m; i += kThreads) out[pend_at + i] = mt_uniform(buf[cur][i]);
    for (int i = t; i < kN; i += kThreads) state_out[i] = buf[cur][i];
    if (t == 0) {
        state_out[kN] = (uint32_t)left;
        state_out[kN + 1] = (uint32_t)next;
    }
}
}  // namespace

// Host twin of the kernel's bookkeeping: where n draws leave the engine (no numbers are produced).  The graph form of the
// shift draw uses it so that torch's generator can be advanced without waiting for the device.
extern "C" int texir_mt19937_advance_host(uint32_t* state, int* left_io, int* next_io, int64_t n) {
    if (!state || !left_io || !next_io || n < 0) return fail("texir_mt19937_advance_host: bad argument");
    int left = *left_io, next = *next_io;
    if (left < 0 || left > kN + 1 || next < 0 || next > kN || (left > 1 && next + left - 1 > kN))
        return fail("texir_mt19937_advance_host: inconsistent generator state");
    auto twist = [](uint32_t u, uint32_t v) {
        return (((u & 0x80000000u) | (v & 0x7FFFFFFFu)) >> 1) ^ ((v & 1u) ? 0x9908B0DFu : 0u);
    };
    int64_t produced = 0;
    {
        const int avail = left > 1 ? left - 1 : 0;
        const int k = (int)(n < avail ? n : avail);
        produced = k;
        left -= k;
        next += k;
    }
    constexpr int kA = kN - kM;
    while (produced < n) {
        uint32_t* p = state;
        for (int i = 0; i < kA; ++i) p[i] = p[i + kM] ^ twist(p[i], p[i + 1]);
        for (int i = kA; i < kN - 1; ++i) p[i] = p[i - kA] ^ twist(p[i], p[i + 1]);
        p[kN - 1] = p[kM - 1] ^ twist(p[kN - 1], p[0]);
        const int m = (int)(n - produced < kN ? n - produced : kN);
        produced += m;
        left = kN + 1 - m;
        next = m;
    }
    *left_io = left;
    *next_io = next;
    return 0;
}

extern "C" int texir_mt19937_uniform(const uint32_t* state_in, int left, int next, int64_t n, float* out, uint32_t* state_out,
                                     void* stream) {
    if (!state_in || !out || !state_out || n <= 0) return fail("texir_mt19937_uniform: bad argument");
    if (left != -1 && (left < 0 || left > kN + 1 || next < 0 || next > kN || (left > 1 && next + left - 1 > kN)))
        return fail("texir_mt19937_uniform: inconsistent generator state");
    texir_host::ScopedDevice guard(texir_host::device_of(out));
    k_mt19937_uniform<<<1, kThreads, 0, (cudaStream_t)stream>>>(state_in, left, next, n, out, state_out);
    return finish_launch("k_mt19937_uniform");
}
