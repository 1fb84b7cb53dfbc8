"""IrT post-processing on the GPU: what sits between IrT generation and material training.

    pad_texture          the array form of tools/padding_texture.py:49-80 -- empty texels take their nearest
                         non-empty texel (scipy's distance_transform_edt indices + grid_sample 'nearest')
    padding_texture      the reference tool itself: read `path`, pad, write the 't' file, call the denoiser hook
    imwrite_hdr/imread_hdr   cv2.imwrite / cv2.imread(path, -1) for Radiance .hdr (trainer/generate_ir_texture.py:82,
                         models/mat_nvdiffrast.py:73,89): RGBE pixel codec on the device, run-length container in
                         host C++ (libtexir_b200.so), byte-identical to OpenCV's files
    save_irr_texture     trainer/generate_ir_texture.py:80-82

No cv2 / scipy is imported here and nothing falls back to them.
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, Optional

import torch

from . import _lib
from .scene import _ptr, _stream_ptr


def _need_cuda(name: str, t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.TexirError("%s: tensors must live on a CUDA device (there is no CPU path)" % name)


def pad_texture(img: torch.Tensor, return_indices: bool = False):
    """img [H,W,C] float32 CUDA (C <= 4) -> padded copy; with return_indices also scipy's `indices` as
    int32 [H,W,2] (row, col of the nearest non-empty texel)."""
    _need_cuda("pad_texture", img)
    if img.dim() != 3 or img.shape[2] > 4:
        raise ValueError("pad_texture expects [H,W,C] with C <= 4, got %s" % (tuple(img.shape),))
    src = img.detach().to(torch.float32).contiguous()
    H, W, C = (int(s) for s in src.shape)
    out = torch.empty_like(src)
    scratch = torch.empty((H, W), dtype=torch.int32, device=src.device)
    index = torch.empty((H, W, 2), dtype=torch.int32, device=src.device) if return_indices else None
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().texir_pad_texture(_ptr(src), H, W, C, _ptr(out), _ptr(index) if index is not None else None,
                                                 _ptr(scratch), _stream_ptr(src.device)), "texir_pad_texture")
    return (out, index) if return_indices else out


def rgbe_encode(img: torch.Tensor, bgr: bool = True) -> torch.Tensor:
    """img [...,3] float32 CUDA (cv2's BGR order by default) -> uint8 [...,4] RGBE in file order"""
    _need_cuda("rgbe_encode", img)
    src = img.detach().to(torch.float32).contiguous()
    out = torch.empty(src.shape[:-1] + (4,), dtype=torch.uint8, device=src.device)
    n = src.numel() // 3
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().texir_rgbe_encode(_ptr(src), n, int(bgr), _ptr(out), _stream_ptr(src.device)),
                   "texir_rgbe_encode")
    return out


def rgbe_decode(rgbe: torch.Tensor, bgr: bool = True) -> torch.Tensor:
    """uint8 [...,4] CUDA -> float32 [...,3] (BGR by default, as cv2.imread returns it)"""
    _need_cuda("rgbe_decode", rgbe)
    src = rgbe.contiguous()
    if src.dtype != torch.uint8 or src.shape[-1] != 4:
        raise ValueError("rgbe_decode expects uint8 [...,4]")
    out = torch.empty(src.shape[:-1] + (3,), dtype=torch.float32, device=src.device)
    n = src.numel() // 4
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().texir_rgbe_decode(_ptr(src), n, int(bgr), _ptr(out), _stream_ptr(src.device)),
                   "texir_rgbe_decode")
    return out


def imwrite_hdr(path: str, img: torch.Tensor) -> None:
    """cv2.imwrite(path, img) for a float32 BGR [H,W,3] CUDA tensor and a .hdr path"""
    if img.dim() != 3 or img.shape[2] != 3:
        raise ValueError("imwrite_hdr expects [H,W,3], got %s" % (tuple(img.shape),))
    H, W = int(img.shape[0]), int(img.shape[1])
    host = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)
    host.copy_(rgbe_encode(img, bgr=True))
    _lib.check(_lib.load().texir_hdr_write_host(os.fsencode(path), host.data_ptr(), H, W), "texir_hdr_write_host")


def imread_hdr(path: str, device="cuda") -> torch.Tensor:
    """cv2.imread(path, -1) for a .hdr file -> float32 BGR [H,W,3] on `device`"""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.TexirError("imread_hdr: the pixel codec runs on a CUDA device (there is no CPU path)")
    lib = _lib.load()
    h, w = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.texir_hdr_read_info(os.fsencode(path), ctypes.byref(h), ctypes.byref(w)), "texir_hdr_read_info")
    host = torch.empty((h.value, w.value, 4), dtype=torch.uint8, pin_memory=True)
    _lib.check(lib.texir_hdr_read_host(os.fsencode(path), host.data_ptr(), h.value, w.value), "texir_hdr_read_host")
    return rgbe_decode(host.to(device, non_blocking=True), bgr=True)


def imread_hdr_rgbe(path: str) -> torch.Tensor:
    """the raw RGBE texels of a .hdr file: uint8 [H,W,4] in file order (R, G, B mantissas, shared exponent), pinned host
    memory -- what Scene.set_texture_rgbe takes (no float decode at all)"""
    lib = _lib.load()
    h, w = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.texir_hdr_read_info(os.fsencode(path), ctypes.byref(h), ctypes.byref(w)), "texir_hdr_read_info")
    pin = torch.cuda.is_available()
    host = torch.empty((h.value, w.value, 4), dtype=torch.uint8, pin_memory=pin)
    _lib.check(lib.texir_hdr_read_host(os.fsencode(path), host.data_ptr(), h.value, w.value), "texir_hdr_read_host")
    return host


def save_irr_texture(irr_texture: torch.Tensor, path: str) -> None:
    """trainer/generate_ir_texture.py:80-82: the model's RGB [H,W,3] texture is written channel-reversed (BGR)"""
    imwrite_hdr(path, irr_texture.flip(-1))


def padding_texture(path: str, denoise: Optional[Callable[[str, str], None]] = None, device="cuda") -> str:
    """tools/padding_texture.py:49-87 with the same file naming: reads `path` (…0_irr_texture.hdr), writes the
    padded texture to path.replace('0_irr_texture', 't') and returns that path.  The reference then shells out to
    an external oidnDenoise binary; here `denoise(src_path, dst_path)` is an in-process hook called with the same
    two paths (dst = path.replace('0_irr_texture', '0_irr_texture_denoised_padding')), skipped when None."""
    img = imread_hdr(path, device)
    padded = pad_texture(img)
    dst = path.replace("0_irr_texture", "t")
    imwrite_hdr(dst, padded)
    if denoise is not None:
        denoise(dst, path.replace("0_irr_texture", "0_irr_texture_denoised_padding"))
    return dst
