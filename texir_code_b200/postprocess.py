"""IrT post-processing on the GPU: what sits between IrT generation and material training.

    pad_texture          the array form of tools/padding_texture.py:49-80 -- empty texels take their nearest
                         non-empty texel (scipy's distance_transform_edt indices + grid_sample 'nearest')
    padding_texture      the reference tool itself: read `path`, pad, write the 't' file, call the denoiser hook
    imwrite_hdr/imread_hdr   cv2.imwrite / cv2.imread(path, -1) for Radiance .hdr (trainer/generate_ir_texture.py:82,
                         models/mat_nvdiffrast.py:73,89): RGBE pixel codec on the device, run-length container in
                         host C++ (libtexir_b200.so), byte-identical to OpenCV's files
    save_irr_texture     trainer/generate_ir_texture.py:80-82
    padding_index_texture    tools/padding_texture.py:8-43: the same padding for the 16-bit PNG index texture
    imread_png / imwrite_png  cv2.imread(path, -1) / cv2.imwrite for 8/16-bit PNG (chunks + zlib here, scanline filters
                         in libtexir_b200.so); channel order as cv2 returns it (BGR / BGRA)

No cv2 / scipy is imported here and nothing falls back to them.
"""
from __future__ import annotations

import ctypes
import os
import struct
import zlib
from typing import Callable, Optional

import numpy as np

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


_PNG_SIG = b"\x89PNG\r\n\x1a\n"
_PNG_CHANNELS = {0: 1, 2: 3, 4: 2, 6: 4}     # colour type -> samples per pixel (palette images are not index textures)


def imread_png(path: str) -> np.ndarray:
    """cv2.imread(path, -1) for a non-interlaced 8- or 16-bit grey / RGB / RGBA PNG: uint8 / uint16 [H,W] or [H,W,C]
    with cv2's channel order (BGR, BGRA)."""
    data = open(path, "rb").read()
    if data[:8] != _PNG_SIG:
        raise _lib.TexirError("imread_png: %s is not a PNG file" % path)
    pos, idat, head = 8, [], None
    while pos + 8 <= len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if len(body) != n:
            raise _lib.TexirError("imread_png: truncated chunk in %s" % path)
        if kind == b"IHDR":
            head = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
        pos += 12 + n
    if head is None or not idat:
        raise _lib.TexirError("imread_png: %s has no IHDR / IDAT" % path)
    W, H, depth, ctype, _, _, interlace = head
    if depth not in (8, 16) or ctype not in _PNG_CHANNELS or interlace != 0:
        raise _lib.TexirError("imread_png: unsupported PNG (bit depth %d, colour type %d, interlace %d)" % (depth, ctype, interlace))
    C = _PNG_CHANNELS[ctype]
    bpp = C * depth // 8
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8)
    if raw.size != H * (1 + W * bpp):
        raise _lib.TexirError("imread_png: %s: IDAT holds %d bytes, expected %d" % (path, raw.size, H * (1 + W * bpp)))
    out = np.empty((H, W * bpp), np.uint8)
    _lib.check(_lib.load().texir_png_unfilter_host(raw.ctypes.data, H, W * bpp, bpp, out.ctypes.data), "texir_png_unfilter_host")
    img = out.view(">u2").astype(np.uint16) if depth == 16 else out
    img = img.reshape(H, W, C)
    if C >= 3:
        img = img[:, :, [2, 1, 0] + ([3] if C == 4 else [])]
    return np.ascontiguousarray(img[:, :, 0] if C == 1 else img)


def imwrite_png(path: str, img: np.ndarray) -> None:
    """cv2.imwrite(path, img) for uint8 / uint16 [H,W], [H,W,3] (BGR) or [H,W,4] (BGRA): the same pixels in the file
    (the compressed bytes differ from OpenCV's: filter type 0 on every scanline)"""
    a = np.asarray(img)
    if a.dtype not in (np.uint8, np.uint16) or a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] not in (1, 3, 4)):
        raise ValueError("imwrite_png expects uint8 / uint16 [H,W], [H,W,3] or [H,W,4]")
    if a.ndim == 2:
        a = a[:, :, None]
    H, W, C = a.shape
    if C >= 3:
        a = a[:, :, [2, 1, 0] + ([3] if C == 4 else [])]
    depth = 16 if a.dtype == np.uint16 else 8
    rows = np.ascontiguousarray(a.astype(">u2") if depth == 16 else a).view(np.uint8).reshape(H, -1)
    body = np.concatenate([np.zeros((H, 1), np.uint8), rows], axis=1).tobytes()

    def chunk(kind, payload):
        return struct.pack(">I", len(payload)) + kind + payload + struct.pack(">I", zlib.crc32(kind + payload) & 0xFFFFFFFF)

    ctype = {1: 0, 3: 2, 4: 6}[C]
    with open(path, "wb") as f:
        f.write(_PNG_SIG + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(body, 1)) + chunk(b"IEND", b""))


def padding_index_texture(path: str, device="cuda") -> str:
    """tools/padding_texture.py:8-43 with the same file naming: reads the 16-bit index texture `path` (….png), gives every
    empty texel (channel sum 0) the index of its nearest non-empty texel, writes path.replace('.png', '_padding.png')
    and returns that path.  The indices go through float32 exactly as in the reference (they are < 2^16)."""
    img = imread_png(path)
    if img.ndim != 3:
        raise ValueError("padding_index_texture expects a colour PNG (the reference unpacks h, w, c)")
    out_dtype = np.uint16                                             # the tool always casts to uint16 (:39)
    src = torch.from_numpy(img.astype(np.float32)).to(device)
    if src.shape[2] > 3:
        # channels beyond the first three do not enter the mask (:12) but are sampled like the others (:36-37): pad an
        # image of (row + 1, col + 1) with the same mask -- it comes back holding the texel grid_sample picked -- and gather
        H, W = int(src.shape[0]), int(src.shape[1])
        full = (src[:, :, 0] + src[:, :, 1] + src[:, :, 2] != 0.0).to(torch.float32)
        rows = torch.arange(1, H + 1, device=src.device, dtype=torch.float32).view(H, 1).expand(H, W)
        cols = torch.arange(1, W + 1, device=src.device, dtype=torch.float32).view(1, W).expand(H, W)
        where = pad_texture(torch.stack([rows * full, cols * full, torch.zeros_like(full)], -1))
        picked = (where[..., 0] > 0).unsqueeze(-1)                    # grid_sample's zero padding: nothing picked
        got = src[(where[..., 0].long() - 1).clamp_(min=0), (where[..., 1].long() - 1).clamp_(min=0)] * picked
        padded = torch.where(full.bool().unsqueeze(-1), src, got)
    else:
        padded = pad_texture(src)
    dst = path.replace(".png", "_padding.png")
    imwrite_png(dst, padded.cpu().numpy().astype(out_dtype))
    return dst
