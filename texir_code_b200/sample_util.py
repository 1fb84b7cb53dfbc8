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


def draw_shift(b: int, device) -> torch.Tensor:
    """the reference's ``shift = torch.rand(b, 1, 2).cuda()`` (utils/sample_util.py:102): the same draw from the CPU
    default generator, then moved.  The draw lands in a pinned staging buffer and the copy is asynchronous: a pageable
    ``.cuda()`` would make the host wait for everything queued on the stream first, once per generate_dir call.  Three
    staging buffers per size rotate; one is reused only after the copy that read it has finished."""
    dev = torch.device(device)
    if dev.type != "cuda":
        return torch.rand(b, 1, 2).to(dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
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
