"""texir_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracle for the TexIR hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(texir_code_b200) never does.

What is here
  * OracleScene        -- ctypes front-end of oracle/cast_rays_oracle.c (the restated Open3D/Embree
                          closest-hit contract); its cast_rays() honours Open3D's output contract
                          (dict of objects with .numpy()), so the REAL reference query_irf can run
                          on top of it (oracle/make_golden.py does exactly that).
  * torch restatements of the reference's Python, device-agnostic, same operation order so that
    float32 rounding is the reference's own:
        generate_dir            utils/sample_util.py:63-146
        query_irf               models/mat_nvdiffrast.py:292-320, models/tracer_o3d_irt.py:240-269
        specular_reflectance    models/mat_nvdiffrast.py:260-279 (eps 1e-14) /
                                models/test_nvdiffrast.py:318-338 (eps 1e-6)
        diffuse_reflectance     models/mat_nvdiffrast.py:252-258
        render                  models/mat_nvdiffrast.py:201-249, relighting branch
                                models/test_nvdiffrast.py:275-281
        irt_forward             models/tracer_o3d_irt.py:156-173
    The restatements are pinned against the reference's own functions executed in the build
    container (tests/golden/*.npz, written by oracle/make_golden.py; tests/test_oracle.py).

Parity status: everything except the ray/triangle closest hit is pinned to the reference's own
code.  The closest hit is "parity unpinned" (open3d/Embree cannot be installed here and the
reference has no golden vectors for it); see cast_rays_oracle.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libcast_rays_oracle.so")
_SRC = os.path.join(_HERE, "cast_rays_oracle.c")

TINY_NUMBER = 1e-6
TINY_TINY_NUMBER = 1e-14


def build_oracle_lib(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off; x86-64-v3 so the file built here also runs on the GPU box."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(_BUILD, exist_ok=True)
        cmd = ["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC,
               "-lm", "-lpthread"]
        subprocess.check_call(cmd)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build_oracle_lib()
        lib = ctypes.CDLL(_SO)
        lib.oracle_scene_create.restype = ctypes.c_void_p
        lib.oracle_scene_create.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        lib.oracle_scene_destroy.argtypes = [ctypes.c_void_p]
        lib.oracle_cast_rays.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


class _O3DTensor:
    """stand-in for open3d.core.Tensor: the reference only calls .numpy() on the results"""

    def __init__(self, a: np.ndarray):
        self._a = a

    def numpy(self) -> np.ndarray:
        return self._a


class OracleScene:
    """Closest-hit scene with Open3D's RaycastingScene.cast_rays contract."""

    def __init__(self, vertices: np.ndarray, triangles: np.ndarray, brute: bool = False, threads: int = 0):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32)
        self.brute = bool(brute)
        self.threads = int(threads) if threads > 0 else (os.cpu_count() or 1)
        lib = _load()
        self._h = lib.oracle_scene_create(self.vertices.ctypes.data, self.vertices.shape[0],
                                          self.triangles.ctypes.data, self.triangles.shape[0])
        if not self._h:
            raise ValueError("oracle scene needs at least one triangle")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.oracle_scene_destroy(h)

    def cast_rays_np(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        shape = rays.shape[:-1]
        flat = rays.reshape(-1, 6)
        n = flat.shape[0]
        t = np.empty(n, np.float32)
        ids = np.empty(n, np.uint32)
        uv = np.empty((n, 2), np.float32)
        if n:
            _load().oracle_cast_rays(self._h, flat.ctypes.data, n, int(self.brute), self.threads, t.ctypes.data,
                                     ids.ctypes.data, uv.ctypes.data)
        return t.reshape(shape), ids.reshape(shape), uv.reshape(shape + (2,))

    def cast_rays(self, rays):
        if isinstance(rays, _O3DTensor):
            rays = rays.numpy()
        t, ids, uv = self.cast_rays_np(np.asarray(rays))
        return {"t_hit": _O3DTensor(t), "primitive_ids": _O3DTensor(ids), "primitive_uvs": _O3DTensor(uv)}


# --------------------------------------------------------------------------------------------------
# Hammersley (utils/sample_util.py:28-41)
# --------------------------------------------------------------------------------------------------
def radical_inverse(bits: int) -> float:
    bits = ((bits << 16) | (bits >> 16)) & 0xFFFFFFFF
    bits = ((bits & 0x55555555) << 1) | ((bits & 0xAAAAAAAA) >> 1)
    bits = ((bits & 0x33333333) << 2) | ((bits & 0xCCCCCCCC) >> 2)
    bits = ((bits & 0x0F0F0F0F) << 4) | ((bits & 0xF0F0F0F0) >> 4)
    bits = ((bits & 0x00FF00FF) << 8) | ((bits & 0xFF00FF00) >> 8)
    return float(bits) * 2.3283064365386963e-10


def hammersley_table(n: int) -> np.ndarray:
    """[n,2] float32, vectorised but value-identical to the reference's Python loop (:94-98)"""
    i = np.arange(n, dtype=np.uint64)
    b = i.copy()
    b = ((b << np.uint64(16)) | (b >> np.uint64(16))) & np.uint64(0xFFFFFFFF)
    b = ((b & np.uint64(0x55555555)) << np.uint64(1)) | ((b & np.uint64(0xAAAAAAAA)) >> np.uint64(1))
    b = ((b & np.uint64(0x33333333)) << np.uint64(2)) | ((b & np.uint64(0xCCCCCCCC)) >> np.uint64(2))
    b = ((b & np.uint64(0x0F0F0F0F)) << np.uint64(4)) | ((b & np.uint64(0xF0F0F0F0)) >> np.uint64(4))
    b = ((b & np.uint64(0x00FF00FF)) << np.uint64(8)) | ((b & np.uint64(0xFF00FF00)) >> np.uint64(8))
    out = np.empty((n, 2), np.float32)
    out[:, 0] = (i.astype(np.float64) / float(n)).astype(np.float32)
    out[:, 1] = (b.astype(np.float64) * 2.3283064365386963e-10).astype(np.float32)
    return out


# --------------------------------------------------------------------------------------------------
# generate_dir (utils/sample_util.py:63-146)
# --------------------------------------------------------------------------------------------------
def generate_dir(normals: torch.Tensor, num_sample_dir: int, samples=None, mode: str = "uniform",
                 roughness: Optional[torch.Tensor] = None, pre_mode: str = "Hammersley",
                 shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`shift` [b,1,2] replaces the reference's internal torch.rand(b,1,2) draw when given."""
    dev = normals.device
    b = normals.shape[0]
    nrm = normals.unsqueeze(1).expand(b, num_sample_dir, 3)
    x_axis = torch.where(torch.abs(nrm[:, :, 0:1]) > 0.99, torch.tensor([0, 1.0, 0.0], device=dev),
                         torch.tensor([1.0, 0.0, 0.0], device=dev))

    def norm_axis(x):
        return x / (torch.norm(x, dim=-1, keepdim=True) + TINY_NUMBER)

    nrm = norm_axis(nrm)
    U = norm_axis(torch.cross(x_axis, nrm, dim=-1))
    V = norm_axis(torch.cross(nrm, U, dim=-1))

    if pre_mode == "Hammersley":
        smp = torch.from_numpy(hammersley_table(num_sample_dir)).unsqueeze(0).to(dev)
        smp = smp.repeat(b, 1, 1).detach()
        if shift is None:
            shift = torch.rand(b, 1, 2)
        smp = smp + shift.to(dev)
        over = smp > 1.0
        smp[over] = smp[over] - 1.0
        under = smp < 0.0
        smp[under] = smp[under] + 1
        smp = torch.clamp(smp, 0 + TINY_NUMBER, 1 - TINY_NUMBER)
    else:
        smp = torch.rand((b, num_sample_dir, 2)).to(dev)

    phi = 2 * np.pi * smp[:, :, 1:2] - np.pi
    if mode == "uniform":
        cos_t = 1.0 - smp[:, :, 0:1]
        sin_t = torch.sqrt(1.0 - cos_t * cos_t)
    elif mode == "cosine":
        cos_t = torch.sqrt(1.0 - smp[:, :, 0:1])
        sin_t = torch.sqrt(1.0 - cos_t * cos_t)
    elif mode == "importance":
        a = roughness * roughness
        a = a.unsqueeze(1).expand(b, num_sample_dir, 1)
        cos_t = torch.sqrt((1.0 - smp[:, :, 0:1]) / (1.0 + (a * a - 1) * smp[:, :, 0:1]))
        cos_t = torch.clamp(cos_t, min=-1.0 + TINY_NUMBER, max=1.0 - TINY_NUMBER)
        sin_t = torch.clamp(torch.sqrt(1.0 - cos_t * cos_t), min=-1.0 + TINY_NUMBER, max=1.0 - TINY_NUMBER)
    else:
        raise ValueError(mode)
    return V * (torch.sin(phi) * sin_t) + nrm * cos_t + U * -(torch.cos(phi) * sin_t)


# --------------------------------------------------------------------------------------------------
# query_irf (models/tracer_o3d_irt.py:240-269 -- the shape-agnostic copy)
# --------------------------------------------------------------------------------------------------
def query_irf(scene, triangle_uvs: np.ndarray, texture: torch.Tensor, points: torch.Tensor,
              directions: torch.Tensor, num_sample: int) -> torch.Tensor:
    """scene: anything with Open3D's cast_rays; triangle_uvs float64 [3T,2]; texture [1,3,H,W] CPU.
    points [P,N,3], directions [P,N,1,3] -> [P,N,3] on points.device."""
    batch = points.shape[0]
    rays = torch.cat([points.unsqueeze(-2).expand_as(directions), directions], dim=-1)
    hits = scene.cast_rays(rays.detach().cpu().numpy())
    t_hit = hits["t_hit"].numpy()
    mask = np.logical_and(t_hit > 1e-4, np.isfinite(t_hit))
    prim = hits["primitive_ids"].numpy().copy()
    bary = np.clip(hits["primitive_uvs"].numpy(), 0.0, 1.0)
    prim[~mask] = 0
    corner = np.stack([prim * 3 + 0, prim * 3 + 1, prim * 3 + 2], axis=0).reshape(-1)
    cuv = triangle_uvs[corner].reshape(3, batch, num_sample, 1, 2)
    grid = cuv[0] * (1 - bary[..., 0:1] - bary[..., 1:2]) + cuv[1] * bary[..., 0:1] + cuv[2] * bary[..., 1:2]
    grid = torch.from_numpy(grid).float()
    grid[..., 0:1] = grid[..., 0:1] * 2 - 1
    grid[..., 1:2] = -(1 - grid[..., 1:2] * 2)
    rad = F.grid_sample(texture.expand([batch] + list(texture.shape[1:])), grid, mode="bilinear",
                        padding_mode="border", align_corners=False).permute(0, 2, 3, 1)
    rad[~mask, :] = 0
    return rad.reshape(batch, num_sample, 3).to(points.device)


# --------------------------------------------------------------------------------------------------
# BRDF reductions (models/mat_nvdiffrast.py:252-279)
# --------------------------------------------------------------------------------------------------
def diffuse_reflectance(lighting, l, n, albedo, sample_type="uniform"):
    ndl = torch.clamp(torch.sum(n.unsqueeze(1) * l, dim=-1, keepdim=True), 0.0, 1.0)
    brdf = albedo.unsqueeze(1) / np.pi
    if sample_type == "cosine":
        return torch.sum(lighting * brdf * np.pi, dim=1)
    return torch.sum(lighting * brdf * ndl * 2 * np.pi, dim=1)


def specular_reflectance(lighting, h, n, v, l, roughness, albedo=None, eps: float = TINY_TINY_NUMBER):
    vdh = torch.clamp(torch.sum(h * v.unsqueeze(1), dim=-1, keepdim=True), 0.0, 1.0)
    ndl = torch.clamp(torch.sum(n.unsqueeze(1) * l, dim=-1, keepdim=True), 0.0, 1.0)
    ndh = torch.clamp(torch.sum(n.unsqueeze(1) * h, dim=-1, keepdim=True), 0.0, 1.0)
    ndv = torch.clamp(torch.sum(n.unsqueeze(1) * v.unsqueeze(1), dim=-1, keepdim=True), 0.0, 1.0)
    f = 0.04 + 0.96 * torch.pow(2.0, ((-5.55472 * vdh - 6.98316) * vdh))
    k = (roughness.unsqueeze(1) + 1) * (roughness.unsqueeze(1) + 1) / 8
    g1_ndv = ndv / torch.clamp(ndv * (1 - k) + k, min=eps)
    g1_ndl = ndl / torch.clamp(ndl * (1 - k) + k, min=eps)
    g = g1_ndl * g1_ndv
    brdf = f * g / torch.clamp(4 * ndl * ndv, min=eps)
    return torch.sum(lighting * brdf * ndl * 4 * vdh / torch.clamp(ndh, eps), dim=1)


# --------------------------------------------------------------------------------------------------
# render (models/mat_nvdiffrast.py:201-249; relighting branch models/test_nvdiffrast.py:275-281)
# --------------------------------------------------------------------------------------------------
def render(scene, triangle_uvs, texture, sample_l, sample_type, normal, albedo, roughness, points, cam_position,
           irr, shift_spec=None, shift_diff=None, relighting: bool = False, eps: float = TINY_TINY_NUMBER):
    shp = normal.shape[:-1]
    normal = normal.reshape(-1, 3)
    albedo = albedo.reshape(-1, 3)
    roughness = roughness.reshape(-1, 1)
    points = points.reshape(-1, 3)
    view = F.normalize(cam_position.unsqueeze(0) - points, eps=1e-4)
    if relighting:
        l_diff = generate_dir(normal, sample_l[0], None, mode=sample_type[0], shift=shift_diff)
        rad = query_irf(scene, triangle_uvs, texture, points.unsqueeze(1).expand_as(l_diff), l_diff.unsqueeze(-2),
                        sample_l[0])
        diffuse = diffuse_reflectance(rad, l_diff, normal, albedo, sample_type[0]) / sample_l[0]
    else:
        diffuse = irr.reshape(-1, 3) * albedo / np.pi
    h = generate_dir(normal, sample_l[1], None, sample_type[1], roughness, shift=shift_spec)
    vdh = torch.clamp(torch.sum(h * view.unsqueeze(1), dim=-1, keepdim=True), 0.0, 1.0)
    l_spec = 2 * vdh * h - view.unsqueeze(1)
    rad = query_irf(scene, triangle_uvs, texture, points.unsqueeze(1).expand_as(l_spec),
                    l_spec.unsqueeze(-2).detach(), sample_l[1])
    specular = specular_reflectance(rad, h, normal, view, l_spec, roughness, eps=eps) / sample_l[1]
    return {
        "rgb": (diffuse + specular).reshape(*shp, 3),
        "albedo": albedo.reshape(*shp, 3),
        "normal": normal.reshape(*shp, 3).detach(),
        "position": (points + 2e-2 * normal).reshape(*shp, 3).detach(),
    }


# --------------------------------------------------------------------------------------------------
# IrT hot loop (models/tracer_o3d_irt.py:156-173)
# --------------------------------------------------------------------------------------------------
def irt_forward(scene, triangle_uvs, texture, position, normal, num_samples: int, sample_type: str = "uniform",
                batch: int = 512, shifts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """position, normal [P,3] with P % batch == 0; shifts [P,1,2] (else drawn per batch like the reference)."""
    P = position.shape[0]
    assert P % batch == 0
    out = torch.zeros_like(position)
    for i in range(P // batch):
        sl = slice(i * batch, (i + 1) * batch)
        pos, nrm = position[sl], normal[sl]
        sh = None if shifts is None else shifts[sl]
        l = generate_dir(nrm, num_samples, None, sample_type, None, shift=sh)
        rad = query_irf(scene, triangle_uvs, texture, pos.unsqueeze(1).expand_as(l), l.unsqueeze(-2), num_samples)
        ndl = torch.clamp(torch.sum(nrm.unsqueeze(1) * l.reshape(batch, -1, 3), dim=-1, keepdim=True), 0.0, 1.0)
        out[sl] = torch.sum(rad.reshape(batch, -1, 3).to(ndl.device) * ndl, dim=1) * 2 * np.pi / num_samples
    return out


# ------------------------------------------------------------------------------------------------
# IrT post-processing: tools/padding_texture.py:49-80 on arrays, and OpenCV's RGBE pixel codec
# (cv2.imwrite / cv2.imread of .hdr; trainer/generate_ir_texture.py:80-82, models/mat_nvdiffrast.py:73,89).
# Pinned by tests/golden/post_*.npz, which oracle/make_golden_post.py wrote by running the reference's own
# padding_texture() and cv2 (tests/test_oracle.py).
# ------------------------------------------------------------------------------------------------
def padding_texture_array(img: np.ndarray):
    """tools/padding_texture.py:50-77 without the file I/O: img float32 [H,W,3] -> (padded [H,W,3], indices [2,H,W])"""
    from scipy import ndimage
    img = np.asarray(img, np.float32)
    h, w, c = img.shape
    intensity = img[:, :, 0] + img[:, :, 1] + img[:, :, 2]
    mask = np.asarray(intensity == 0.0, dtype=np.uint8)
    _, indices_np = ndimage.distance_transform_edt(mask, return_indices=True)
    indices = torch.from_numpy(indices_np).permute(1, 2, 0).reshape(-1, 2)
    img_torch = torch.from_numpy(img).permute(2, 0, 1).unsqueeze(0)
    uv_init = torch.zeros((h * w, 2), dtype=torch.float32)
    sel = torch.from_numpy(mask.reshape(-1).astype(bool))   # the reference indexes with the uint8 mask (legacy bool)
    uv_init[sel] = indices[sel][:, [1, 0]].float() / torch.tensor([w, h]).unsqueeze(0) * 2. - 1.
    uv_init = uv_init.reshape(h, w, 2).unsqueeze(0)
    res = F.grid_sample(img_torch, uv_init, mode="nearest", align_corners=False)[0].permute(1, 2, 0).numpy()
    res = res * np.asarray(mask, np.float32)[:, :, np.newaxis] + img * np.asarray(1 - mask, np.float32)[:, :, np.newaxis]
    return res, indices_np.astype(np.int32)


def float2rgbe(img_bgr: np.ndarray) -> np.ndarray:
    """OpenCV rgbe.cpp float2rgbe on a BGR float32 image -> bytes [..., 4] in file order (R, G, B, exponent)"""
    img_bgr = np.asarray(img_bgr, np.float32)
    r, g, b = img_bgr[..., 2], img_bgr[..., 1], img_bgr[..., 0]
    v = np.maximum(np.maximum(r, g), b)
    m, e = np.frexp(v.astype(np.float64))
    ok = v.astype(np.float64) >= 1e-32
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(ok, m * 256.0 / v.astype(np.float64), 0.0).astype(np.float32)
    out = np.zeros(img_bgr.shape[:-1] + (4,), np.uint8)
    for k, ch in enumerate((r, g, b)):
        out[..., k] = np.where(ok, (ch * scale).astype(np.int32) & 255, 0)
    out[..., 3] = np.where(ok, e + 128, 0)
    return out


def rgbe2float(rgbe: np.ndarray) -> np.ndarray:
    """OpenCV rgbe.cpp rgbe2float: bytes [..., 4] -> BGR float32 image (what cv2.imread(path, -1) returns)"""
    rgbe = np.asarray(rgbe, np.uint8)
    f = np.ldexp(1.0, rgbe[..., 3].astype(np.int32) - 136).astype(np.float32)
    out = np.stack([rgbe[..., 2] * f, rgbe[..., 1] * f, rgbe[..., 0] * f], -1).astype(np.float32)
    out[rgbe[..., 3] == 0] = 0.0
    return out


# ------------------------------------------------------------------------------------------------
# Material-texture fetch of the Mat stage: dr.texture(..., filter_mode='linear' |
# 'linear-mipmap-linear', max_mip_level=M) at models/mat_nvdiffrast.py:131-139, and the texd that
# dr.interpolate(..., rast_db, diff_attrs='all') feeds it (:128).
# PARITY UNPINNED: nvdiffrast is a third-party CUDA extension, absent from /root/reference and not installable
# here; this restates its published algorithm (Laine et al. 2020, section 3.4: box-filtered pyramid, level of
# detail from the major axis of the pixel footprint, trilinear blend, 'wrap' boundary, texel centres at
# (i + 0.5) / size) in plain differentiable torch, so torch autograd provides the reference gradients.
# ------------------------------------------------------------------------------------------------
def mip_pyramid(tex: torch.Tensor, max_mip_level: Optional[int] = None):
    """tex [H,W,C] -> list of levels, level l+1 = 2x2 box average of level l (2x1 once a side is 1)"""
    levels = [tex]
    while max(levels[-1].shape[0], levels[-1].shape[1]) > 1 and (max_mip_level is None or len(levels) - 1 < max_mip_level):
        t = levels[-1]
        h, w = t.shape[0], t.shape[1]
        if (h > 1 and h % 2) or (w > 1 and w % 2):
            raise ValueError("odd side")
        if h > 1:
            t = 0.5 * (t[0::2] + t[1::2])
        if w > 1:
            t = 0.5 * (t[:, 0::2] + t[:, 1::2])
        levels.append(t)
    return levels


def texture_linear(tex: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """bilinear, wrap: tex [H,W,C], uv [P,2] -> [P,C]"""
    H, W = tex.shape[0], tex.shape[1]
    x = uv[:, 0] * W - 0.5
    y = uv[:, 1] * H - 0.5
    x0, y0 = torch.floor(x), torch.floor(y)
    ax, ay = (x - x0).unsqueeze(-1), (y - y0).unsqueeze(-1)
    x0, y0 = x0.long(), y0.long()
    x1, y1 = (x0 + 1) % W, (y0 + 1) % H
    x0, y0 = x0 % W, y0 % H
    return (tex[y0, x0] * (1 - ax) * (1 - ay) + tex[y0, x1] * ax * (1 - ay) + tex[y1, x0] * (1 - ax) * ay
            + tex[y1, x1] * ax * ay)


def texture_lod(uv_da: torch.Tensor, H: int, W: int, max_level: int) -> torch.Tensor:
    """fractional level of detail per pixel from uv_da [P,4] = (du/dX, du/dY, dv/dX, dv/dY)"""
    dsdx, dsdy, dtdx, dtdy = uv_da[:, 0] * W, uv_da[:, 1] * W, uv_da[:, 2] * H, uv_da[:, 3] * H
    A, B, C = dsdx * dsdx + dtdx * dtdx, dsdy * dsdy + dtdy * dtdy, dsdx * dsdy + dtdx * dtdy
    major2 = 0.5 * (A + B) + torch.sqrt(0.25 * (A - B) * (A - B) + C * C)
    lvl = 0.5 * torch.log2(major2)
    lvl = torch.nan_to_num(lvl, nan=0.0, posinf=float(max_level), neginf=0.0)
    return torch.clamp(lvl, 0.0, float(max_level))


def texture_mipmap_linear(tex: torch.Tensor, uv: torch.Tensor, uv_da: torch.Tensor,
                          max_mip_level: Optional[int] = None) -> torch.Tensor:
    """'linear-mipmap-linear': trilinear over the pyramid; differentiable in tex"""
    levels = mip_pyramid(tex, max_mip_level)
    L = len(levels) - 1
    lod = texture_lod(uv_da.detach(), tex.shape[0], tex.shape[1], L)
    l0 = torch.floor(lod).long()
    l1 = torch.clamp(l0 + 1, max=L)
    frac = torch.where(lod > 0, lod - l0.float(), torch.zeros_like(lod)).unsqueeze(-1)
    out = torch.zeros(uv.shape[0], tex.shape[2], dtype=tex.dtype)
    for l, t in enumerate(levels):
        s = texture_linear(t, uv)
        w = torch.where(l0 == l, 1.0 - frac[:, 0], torch.zeros_like(lod)) + torch.where(
            (l1 == l) & (frac[:, 0] > 0), frac[:, 0], torch.zeros_like(lod))
        out = out + s * w.unsqueeze(-1)
    return out


def gbuffer_uv_da(vertices, triangles, triangle_uvs, rays, ddir, prim_ids) -> np.ndarray:
    """float64 ray differentials of the interpolated uv on the plane of the hit triangle -> [R,4]
    (du/dX, du/dY, dv/dX, dv/dY); rows with prim_ids == 0xFFFFFFFF are 0"""
    v = np.asarray(vertices, np.float64)[np.asarray(triangles, np.int64)]
    tuv = np.asarray(triangle_uvs, np.float64).reshape(-1, 3, 2)
    rays, ddir = np.asarray(rays, np.float64), np.asarray(ddir, np.float64)
    hit = np.asarray(prim_ids) != 0xFFFFFFFF
    ids = np.where(hit, prim_ids, 0).astype(np.int64)
    v0, E1, E2 = v[ids, 0], v[ids, 1] - v[ids, 0], v[ids, 2] - v[ids, 0]
    n = np.cross(E1, E2)
    o, d = rays[:, :3], rays[:, 3:]
    nd = np.einsum("ij,ij->i", n, d)
    nn = np.einsum("ij,ij->i", n, n)
    ok = hit & (nd != 0) & (nn > 0)
    nd, nn = np.where(ok, nd, 1.0), np.where(ok, nn, 1.0)
    t = np.einsum("ij,ij->i", n, v0 - o) / nd
    out = np.zeros((len(rays), 4))
    for k in range(2):
        dd = ddir[:, 3 * k:3 * k + 3]
        dt = -t * np.einsum("ij,ij->i", n, dd) / nd
        dx = dt[:, None] * d + t[:, None] * dd
        db1 = np.einsum("ij,ij->i", np.cross(dx, E2), n) / nn
        db2 = np.einsum("ij,ij->i", np.cross(E1, dx), n) / nn
        duv = db1[:, None] * (tuv[ids, 1] - tuv[ids, 0]) + db2[:, None] * (tuv[ids, 2] - tuv[ids, 0])
        out[:, k] = duv[:, 0]
        out[:, 2 + k] = duv[:, 1]
    out[~ok] = 0.0
    return out
