"""Direction sampler -- GPU mirror of the reference's utils/sample_util.py (same names, same arguments).

    RadicalInverse / Hammersley      utils/sample_util.py:28-41
    generate_fixed_samples           utils/sample_util.py:44-60
    generate_dir                     utils/sample_util.py:63-146

generate_dir is one sm_100a kernel (texir_generate_dir) instead of ~51 ATen launches and a Python
Hammersley loop; it stays differentiable w.r.t. ``roughness`` (importance mode) through
texir_generate_dir_bwd.  The Cranley-Patterson shift is drawn exactly as the reference draws it --
``torch.rand(b, 1, 2)`` on the CPU default generator (utils/sample_util.py:102) -- so seeding the CPU
generator reproduces the reference's sample set.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from .scene import _ptr, _stream_ptr

TINY_NUMBER = 1e-6


def RadicalInverse(bits: int) -> float:
    """van der Corput radical inverse in base 2 (utils/sample_util.py:28-38)"""
    bits = int(bits) & 0xFFFFFFFF
    rev = int("{:032b}".format(bits)[::-1], 2)
    return float(rev) * 2.3283064365386963e-10


def Hammersley(i: int, N: int):
    """utils/sample_util.py:40-41"""
    return [float(i) / float(N), RadicalInverse(i)]


def hammersley_table(n: int) -> np.ndarray:
    """[n,2] float32 table of Hammersley(i, n), i < n"""
    i = np.arange(n, dtype=np.uint32)
    b = i.copy()
    b = (b << np.uint32(16)) | (b >> np.uint32(16))
    b = ((b & np.uint32(0x55555555)) << np.uint32(1)) | ((b & np.uint32(0xAAAAAAAA)) >> np.uint32(1))
    b = ((b & np.uint32(0x33333333)) << np.uint32(2)) | ((b & np.uint32(0xCCCCCCCC)) >> np.uint32(2))
    b = ((b & np.uint32(0x0F0F0F0F)) << np.uint32(4)) | ((b & np.uint32(0xF0F0F0F0)) >> np.uint32(4))
    b = ((b & np.uint32(0x00FF00FF)) << np.uint32(8)) | ((b & np.uint32(0xFF00FF00)) >> np.uint32(8))
    out = np.empty((n, 2), np.float32)
    out[:, 0] = (i.astype(np.float64) / float(n)).astype(np.float32)
    out[:, 1] = (b.astype(np.float64) * 2.3283064365386963e-10).astype(np.float32)
    return out


def generate_fixed_samples(b: int, num_sample_dir: int) -> torch.Tensor:
    """[b, num_sample_dir, 2] Hammersley samples (utils/sample_util.py:44-60), on the current CUDA device"""
    s = torch.from_numpy(hammersley_table(num_sample_dir)).unsqueeze(0)
    return s.cuda().repeat(b, 1, 1).detach()


_shift_ring = {}   # (b, device index) -> [next slot, [(pinned buffer, copy-done event) ...]]
_MT_N, _OFF_LEFT, _OFF_NEXT, _OFF_STATE = 624, 8, 16, 24   # layout of torch.get_rng_state() (CPUGeneratorImplStateLegacy)
_mt_dev = {}       # device index -> dict(stream, pinned in/out, device in/out) of the device-side generator path
_mt_ok = None      # None: not checked yet; True / False: the device stream reproduced torch.rand in this process
MT_MIN_DRAWS = 1 << 15   # below this many numbers the host draws them faster than a kernel round trip


def _draw_cpu(b: int, dev: torch.device, idx: int) -> torch.Tensor:
    """torch.rand(b, 1, 2) into a pinned staging buffer, asynchronous copy (three buffers per size rotate)"""
    ring = _shift_ring.get((b, idx))
    if ring is None:
        if len(_shift_ring) > 16:
            _shift_ring.clear()
        ring = _shift_ring[(b, idx)] = [0, [(torch.empty(b, 1, 2).pin_memory(), None) for _ in range(3)]]
    slot = ring[0]
    ring[0] = (slot + 1) % len(ring[1])
    buf, ev = ring[1][slot]
    if ev is not None:
        ev.synchronize()
    torch.rand(b, 1, 2, out=buf)          # same values as torch.rand(b, 1, 2) for the same generator state
    with torch.cuda.device(idx):
        out = buf.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(idx))
    ring[1][slot] = (buf, ev)
    return out


_mt_pending = None  # (done event, h_out, state tensor, its numpy view): a device draw whose state is not back in torch yet
_defer_depth = 0    # > 0 inside deferred_rng_commit(): draw_shift leaves the hand-back of the state to the end of the block


def finish_pending() -> None:
    """hand the generator state of a deferred device draw back to torch (waits for that draw's kernel)"""
    global _mt_pending
    if _mt_pending is None:
        return
    done, h_out, st, a = _mt_pending
    _mt_pending = None
    done.synchronize()
    _mt_store_state(a, h_out)
    torch.set_rng_state(st)


class deferred_rng_commit:
    """``with deferred_rng_commit():`` -- device draws inside the block do not make the host wait for their kernel; the
    CPU generator receives its advanced state when the block ends (or at the next draw).  Only for blocks that do not
    touch torch's CPU generator themselves: MaterialModel.forward wraps its own body in it, so the state is in place
    before the caller's code runs again."""

    def __enter__(self):
        global _defer_depth
        _defer_depth += 1
        return self

    def __exit__(self, *exc):
        global _defer_depth
        _defer_depth -= 1
        if _defer_depth == 0:
            finish_pending()
        return False


def _mt_store_state(a, h_out) -> None:
    w = h_out.numpy().view(np.uint32)
    if int(w[_MT_N]) == 0xFFFFFFFF:
        raise _lib.TexirError("texir_mt19937_uniform: the generator state handed to the device was inconsistent")
    a[_OFF_STATE:_OFF_STATE + 8 * _MT_N] = w[:_MT_N].astype(np.uint64).view(np.uint8)
    a[_OFF_LEFT:_OFF_LEFT + 4] = np.array([int(w[_MT_N])], np.int32).view(np.uint8)
    a[_OFF_NEXT:_OFF_NEXT + 8] = np.array([int(w[_MT_N + 1])], np.uint64).view(np.uint8)


_graph_rng = None   # the GraphRng whose capture is in progress: draw_shift routes every draw through it


class GraphRng:
    """The CPU generator's stream inside a CUDA graph (graphed.GraphedMatStep).  Kernel arguments are frozen at capture,
    so the whole engine state -- 624 words, left, next -- travels through device memory, and every draw of the captured
    step advances it there (texir_mt19937_uniform, left = -1).  Per replay
    the host copies torch's current state into the device buffer the graph's first draw reads (an ordinary asynchronous copy
    on the replaying stream, from a ring of pinned buffers) and advances torch's generator ITSELF by the step's draws
    (texir_mt19937_advance_host: the twists only, no numbers -- about 0.1 ms of host time that overlaps the previous
    replay), so nothing waits for the device and replays queue back to back."""

    def __init__(self, device):
        dev = torch.device(device)
        self.dev = dev
        self.idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.h_ring = [[torch.empty(_MT_N + 2, dtype=torch.int32).pin_memory(), None] for _ in range(4)]  # [buffer, copy-done event]
        self.slot = 0
        with torch.cuda.device(self.idx):
            self.d = [torch.empty(_MT_N + 2, dtype=torch.int32, device=dev) for _ in range(2)]
            self.stream = torch.cuda.Stream(device=self.idx, priority=-1)
        self.cur, self.counts = 0, []

    def __enter__(self):            # wrap the capture in it
        global _graph_rng
        finish_pending()
        _graph_rng, self.cur, self.counts = self, 0, []
        return self

    def __exit__(self, *exc):
        global _graph_rng
        _graph_rng = None
        return False

    def draw(self, n: int) -> torch.Tensor:
        main = torch.cuda.current_stream(self.idx)
        with torch.cuda.device(self.idx):
            out = torch.empty(n, dtype=torch.float32, device=self.dev)
            self.stream.wait_stream(main)     # a capture only admits streams that branch off the capturing one
            with torch.cuda.stream(self.stream):
                _lib.check(_lib.load().texir_mt19937_uniform(_ptr(self.d[self.cur]), -1, 0, n, _ptr(out), _ptr(self.d[self.cur ^ 1]),
                                                             _stream_ptr(self.dev)), "texir_mt19937_uniform")
                self.cur ^= 1
                self.counts.append(int(n))
            main.wait_stream(self.stream)
        return out

    def before_replay(self) -> None:
        """torch's generator state into the graph's input buffer; torch's generator advanced by the step's draws"""
        if not self.counts:
            return
        finish_pending()
        st = torch.get_rng_state()
        a = st.numpy()
        buf, ev = self.h_ring[self.slot]
        if ev is not None:
            ev.synchronize()                 # four replays ago: long done
        w = buf.numpy().view(np.uint32)
        w[:_MT_N] = a[_OFF_STATE:_OFF_STATE + 8 * _MT_N].view(np.uint64).astype(np.uint32)
        left = ctypes.c_int(int(a[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0]))
        nxt = ctypes.c_int(int(a[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64)[0]))
        w[_MT_N], w[_MT_N + 1] = np.uint32(left.value & 0xFFFFFFFF), np.uint32(nxt.value)
        with torch.cuda.device(self.idx):
            self.d[0].copy_(buf, non_blocking=True)          # stream-ordered before the replay that follows
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.idx))
        self.h_ring[self.slot][1] = ev
        self.slot = (self.slot + 1) % len(self.h_ring)
        adv = w[:_MT_N].copy()
        lib = _lib.load()
        for n in self.counts:
            _lib.check(lib.texir_mt19937_advance_host(adv.ctypes.data, ctypes.byref(left), ctypes.byref(nxt), n),
                       "texir_mt19937_advance_host")
        a[_OFF_STATE:_OFF_STATE + 8 * _MT_N] = adv.astype(np.uint64).view(np.uint8)
        a[_OFF_LEFT:_OFF_LEFT + 4] = np.array([left.value], np.int32).view(np.uint8)
        a[_OFF_NEXT:_OFF_NEXT + 8] = np.array([nxt.value], np.uint64).view(np.uint8)
        torch.set_rng_state(st)

    def after_replay(self) -> None:
        pass


def _draw_mt_device(n: int, dev: torch.device, idx: int, state: Optional[torch.Tensor] = None, commit: bool = True):
    """n values of torch.rand (CPU default generator, or the generator state `state`) produced ON THE DEVICE by
    texir_mt19937_uniform from the generator's own mt19937 state, and the host generator advanced to where torch.rand
    would have left it (commit).  -> (float32 [n] on dev, new state tensor)"""
    global _mt_pending
    finish_pending()
    st = torch.get_rng_state() if state is None else state.clone()
    a = st.numpy()
    left = int(a[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0])
    nxt = int(a[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64)[0])
    ctx = _mt_dev.get(idx)
    if ctx is None:
        with torch.cuda.device(idx):
            ctx = _mt_dev[idx] = {
                "stream": torch.cuda.Stream(device=idx, priority=-1),   # not behind the work queued on the caller's stream
                "h_in": torch.empty(_MT_N, dtype=torch.int32).pin_memory(), "h_out": torch.empty(_MT_N + 2, dtype=torch.int32).pin_memory(),
                "d_in": torch.empty(_MT_N, dtype=torch.int32, device=dev), "d_out": torch.empty(_MT_N + 2, dtype=torch.int32, device=dev)}
    ctx["h_in"].numpy().view(np.uint32)[:] = a[_OFF_STATE:_OFF_STATE + 8 * _MT_N].view(np.uint64).astype(np.uint32)
    main = torch.cuda.current_stream(idx)
    with torch.cuda.stream(ctx["stream"]):     # (the stream context also makes its device current)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        ctx["d_in"].copy_(ctx["h_in"], non_blocking=True)
        _lib.check(_lib.load().texir_mt19937_uniform(_ptr(ctx["d_in"]), left, nxt, n, _ptr(out), _ptr(ctx["d_out"]),
                                                     _stream_ptr(dev)), "texir_mt19937_uniform")
        ctx["h_out"].copy_(ctx["d_out"], non_blocking=True)
        done = torch.cuda.Event()
        done.record(ctx["stream"])
    main.wait_event(done)
    out.record_stream(main)
    if commit and state is None and _defer_depth > 0:
        _mt_pending = (done, ctx["h_out"], st, a)   # the caller's block ends with finish_pending()
        return out, st
    done.synchronize()                     # about one kernel (0.1 ms): the generator state comes back to the host
    _mt_store_state(a, ctx["h_out"])
    if commit:
        torch.set_rng_state(st)
    return out, st


def _mt_device_selfcheck(dev: torch.device, idx: int) -> bool:
    """once per process: the device stream against torch.rand itself, on a copy of the current generator state (values and
    the state torch ends in), from mid-block; a torch build that lays its generator out differently simply keeps the host draw"""
    try:
        st = torch.get_rng_state()
        g = torch.Generator()
        g.set_state(st)
        torch.rand(1000, generator=g)                      # move off the block boundary
        mid = g.get_state()
        ref = torch.rand(5000, generator=g)
        got, st2 = _draw_mt_device(5000, dev, idx, state=mid, commit=False)
        return bool(torch.equal(got.cpu(), ref)) and bool(torch.equal(st2, g.get_state()))
    except Exception:
        return False


def draw_shift(b: int, device) -> torch.Tensor:
    """the reference's ``shift = torch.rand(b, 1, 2).cuda()`` (utils/sample_util.py:102): the same numbers from the same
    CPU default generator, which ends in the same state.  Small draws run torch.rand on the host into a pinned staging
    buffer with an asynchronous copy (a pageable ``.cuda()`` makes the host wait for everything queued on the stream);
    large ones (the 98 304-pixel cube map of a Mat-stage step: 0.3 ... 0.7 ms of serial mt19937 on the host) regenerate
    the generator's stream on the device (texir_mt19937_uniform) and hand the advanced state back to torch.
    TEXIR_SHIFT_DEVICE=0 keeps everything on the host."""
    global _mt_ok
    if _graph_rng is not None:
        return _graph_rng.draw(2 * b).view(b, 1, 2)
    finish_pending()
    dev = torch.device(device)
    if dev.type != "cuda":
        return torch.rand(b, 1, 2).to(dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if 2 * b >= MT_MIN_DRAWS and _mt_ok is not False and os.environ.get("TEXIR_SHIFT_DEVICE", "1") != "0":
        if _mt_ok is None:
            _mt_ok = _mt_device_selfcheck(dev, idx)
        if _mt_ok:
            return _draw_mt_device(2 * b, dev, idx)[0].view(b, 1, 2)
    return _draw_cpu(b, dev, idx)


class _GenerateDir(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normals, roughness, shift, N, mode, per_sample):
        dev = normals.device
        P = normals.shape[0]
        nrm = normals.detach().to(torch.float32).contiguous()
        sh = shift.detach().to(device=dev, dtype=torch.float32).contiguous()
        rough = None
        if mode == _lib.MODE_IMPORTANCE:
            rough = roughness.detach().to(torch.float32).reshape(-1).contiguous()
            if rough.numel() != P:
                raise ValueError("roughness must be [b,1]")
        out = torch.empty(P, N, 3, dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            fn = lib.texir_generate_dir_samples if per_sample else lib.texir_generate_dir
            _lib.check(fn(_ptr(nrm), _ptr(rough), _ptr(sh), P, N, mode, _ptr(out), _stream_ptr(dev)),
                       "texir_generate_dir")
        ctx.mode, ctx.N, ctx.per_sample = mode, N, per_sample
        if mode == _lib.MODE_IMPORTANCE and roughness.requires_grad:
            if per_sample:
                raise NotImplementedError("roughness gradient with pre_mode != 'Hammersley' is not provided")
            ctx.save_for_backward(nrm, rough, sh)
            ctx.rshape = roughness.shape
        return out

    @staticmethod
    def backward(ctx, g_out):
        if ctx.mode != _lib.MODE_IMPORTANCE or not ctx.saved_tensors:
            return None, None, None, None, None, None
        nrm, rough, sh = ctx.saved_tensors
        dev = nrm.device
        P = nrm.shape[0]
        g = g_out.to(torch.float32).contiguous()
        g_rough = torch.empty(P, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().texir_generate_dir_bwd(_ptr(nrm), _ptr(rough), _ptr(sh), P, ctx.N, _ptr(g),
                                                          _ptr(g_rough), _stream_ptr(dev)), "texir_generate_dir_bwd")
        return None, g_rough.reshape(ctx.rshape), None, None, None, None


def generate_dir(normals: torch.Tensor, num_sample_dir: int, samples=None, mode: str = "uniform",
                 roughness: Optional[torch.Tensor] = None, pre_mode: str = "Hammersley",
                 shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same signature and semantics as utils/sample_util.py:63 (``samples`` is ignored there too).

    normals [b,3] CUDA float32; roughness [b,1] for mode='importance'; returns L [b,N,3].
    Extra keyword ``shift`` [b,1,2]: feed the Cranley-Patterson shift instead of drawing it.
    """
    if mode not in _lib.MODES:
        raise ValueError("unknown sampling mode %r" % (mode,))
    if not normals.is_cuda:
        raise _lib.TexirError("generate_dir: normals must be a CUDA tensor (there is no CPU path)")
    b = normals.shape[0]
    m = _lib.MODES[mode]
    if m == _lib.MODE_IMPORTANCE and roughness is None:
        raise ValueError("mode='importance' needs roughness")
    if pre_mode == "Hammersley":
        if shift is None:
            shift = draw_shift(b, normals.device)
        return _GenerateDir.apply(normals, roughness, shift.reshape(b, 2), int(num_sample_dir), m, False)
    smp = torch.rand((b, num_sample_dir, 2)).to(normals.device)  # utils/sample_util.py:111
    return _GenerateDir.apply(normals, roughness, smp, int(num_sample_dir), m, True)


# north-star alias (BASELINE.json names it sample_hemisphere)
sample_hemisphere = generate_dir
