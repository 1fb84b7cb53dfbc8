"""Scene objects of the hot path: the GPU replacement of the reference's Open3D ray-casting state.

Reference state replaced (models/mat_nvdiffrast.py:86-101, models/tracer_o3d_irt.py:74-89):
    self.scene        o3d.t.geometry.RaycastingScene  (Embree, CPU only)      -> Scene / RaycastingScene
    self.triangle_uvs np.float64 [3T,2]                                        -> uploaded with the scene
    self.texture      torch.float32 [1,3,H,W] on the CPU, flipped and exposed  -> Scene.set_texture (RGBA in HBM)

Scene is the native handle (BVH2 + compressed BVH8 + triangles + UVs + texture resident in HBM).
RaycastingScene mimics Open3D's class closely enough that every reference model runs unmodified with
``self.scene`` swapped: ``add_triangles`` and ``cast_rays(rays) -> {'t_hit','primitive_ids',
'primitive_uvs','primitive_normals','geometry_ids'}`` whose values expose ``.numpy()``.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib

INVALID_ID = 0xFFFFFFFF


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_NO_GUARD = contextlib.nullcontext()


def _on(device: torch.device):
    """device guard for a block of launches: nothing at all when `device` is already the current one (the usual case; a
    torch.cuda.device() context costs several microseconds per step and call site)"""
    idx = device.index
    if idx is None or torch.cuda.current_device() == idx:
        return _NO_GUARD
    return torch.cuda.device(device)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _as_f32(x, device) -> torch.Tensor:
    """dense float32 tensor on `device` (accepts numpy arrays and CPU/CUDA tensors)"""
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32))
    return x.detach().to(device=device, dtype=torch.float32).contiguous()


class Scene:
    """BVH + triangles + per-corner UVs (+ radiance texture) resident on one GPU."""

    def __init__(self, vertices, triangles, triangle_uvs=None, device=None, build_threads: int = 0,
                 bvh: int = _lib.BVH2, builder: str = "auto"):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TexirError("texir_code_b200.Scene needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TexirError("Scene device must be a CUDA device, got %s" % self.device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        v = np.ascontiguousarray(_to_numpy(vertices), dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(_to_numpy(triangles), dtype=np.int32).reshape(-1, 3)
        uv_ptr = None
        if triangle_uvs is not None:
            uv = np.ascontiguousarray(_to_numpy(triangle_uvs), dtype=np.float64).reshape(-1, 2)
            if uv.shape[0] != 3 * t.shape[0]:
                raise ValueError("triangle_uvs must be [3*T,2] (got %s for T=%d)" % (uv.shape, t.shape[0]))
            uv_ptr = uv.ctypes.data
        torch.cuda.init()
        h = ctypes.c_void_p()
        # builder: "auto" (the device for >= 4096 triangles), "host" or "device" -- the same tree either way
        _lib.check(lib.texir_scene_create_ex(v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], uv_ptr,
                                             self.device.index, build_threads, _lib.BUILDERS[builder], ctypes.byref(h)),
                   "texir_scene_create")
        self._h = h
        self.num_tris = int(t.shape[0])
        self.has_uvs = triangle_uvs is not None
        self._tex_key = None
        self._tex_ref = None
        if bvh != _lib.BVH2:
            self.set_bvh(bvh)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().texir_scene_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self.close()

    @property
    def handle(self) -> ctypes.c_void_p:
        if not self._h:
            raise _lib.TexirError("scene has been closed")
        return self._h

    # -- configuration ----------------------------------------------------------------------------
    def set_bvh(self, kind: int):
        _lib.check(_lib.load().texir_scene_set_bvh(self.handle, int(kind)), "texir_scene_set_bvh")

    def info(self) -> dict:
        inf = _lib.SceneInfo()
        _lib.check(_lib.load().texir_scene_get_info(self.handle, ctypes.byref(inf)), "texir_scene_get_info")
        return {k: getattr(inf, k) for k, _ in _lib.SceneInfo._fields_}

    def set_texture(self, texture) -> None:
        """texture: the reference's ``self.texture`` -- float32 [1,3,H,W] / [3,H,W] (or [H,W,3]), already
        flipped and exposed (mat_nvdiffrast.py:89-93); CPU or CUDA.  Re-uploads only when the tensor's
        identity or in-place version changed (stage -1 swaps self.texture and restores it,
        mat_nvdiffrast.py:143-150)."""
        if not isinstance(texture, torch.Tensor):
            texture = torch.from_numpy(np.ascontiguousarray(np.asarray(texture), dtype=np.float32))
        key = (texture.data_ptr(), texture._version, tuple(texture.shape), str(texture.device))
        if key == self._tex_key:
            return
        t = texture.detach()
        if t.dim() == 4:
            if t.shape[0] != 1:
                raise ValueError("texture batch must be 1")
            t = t[0]
        if t.dim() != 3:
            raise ValueError("texture must be [1,3,H,W], [3,H,W] or [H,W,3]")
        if t.shape[0] == 3:
            layout, H, W = _lib.TEX_CHW, int(t.shape[1]), int(t.shape[2])
        elif t.shape[2] == 3:
            layout, H, W = _lib.TEX_HWC, int(t.shape[0]), int(t.shape[1])
        else:
            raise ValueError("texture needs 3 channels")
        t = t.to(torch.float32).contiguous()
        on_dev = t.is_cuda
        if on_dev and t.device != self.device:
            t = t.to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().texir_scene_set_texture(self.handle, _ptr(t), H, W, layout, int(on_dev),
                                                           _stream_ptr(self.device)), "texir_scene_set_texture")
            if on_dev:
                torch.cuda.current_stream(self.device).synchronize()  # `t` may be a temporary
        self._tex_key = key
        self._tex_ref = texture  # keeps data_ptr from being recycled by another tensor
        self.tex_hw = (H, W)

    def set_texture_rgbe(self, rgbe, exposure: int = 0, flip: bool = True) -> None:
        """The radiance texture as the RGBE texels of its .hdr file: rgbe uint8 [H,W,4] in file order (R, G, B, E --
        ``postprocess.imread_hdr_rgbe(path)``), CPU or CUDA; ``flip`` = the reference's cv2.flip(texture, 0) and
        ``exposure`` its integer train.hdr_exposure (mat_nvdiffrast.py:89-93).  Bit-identical results to set_texture of
        the decoded float texture, 4 bytes per texel resident instead of 16."""
        if int(exposure) != exposure:
            raise ValueError("the RGBE form needs an integer hdr_exposure (a power-of-two factor folds into the exponent); "
                             "use set_texture for %r" % (exposure,))
        t = rgbe if isinstance(rgbe, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(rgbe), dtype=np.uint8))
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 4:
            raise ValueError("rgbe must be uint8 [H,W,4]")
        t = t.contiguous()
        H, W = int(t.shape[0]), int(t.shape[1])
        key = ("rgbe", t.data_ptr(), t._version, (H, W), int(exposure), bool(flip))
        if key == self._tex_key:
            return
        on_dev = t.is_cuda
        if on_dev and t.device != self.device:
            t = t.to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().texir_scene_set_texture_rgbe(self.handle, _ptr(t), H, W, int(bool(flip)), int(exposure),
                                                                int(on_dev), _stream_ptr(self.device)),
                       "texir_scene_set_texture_rgbe")
            if on_dev:
                torch.cuda.current_stream(self.device).synchronize()
        self._tex_key = key
        self._tex_ref = (rgbe, t)
        self.tex_hw = (H, W)

    def texture_key(self):
        """identity of the radiance texture that is resident now (what a ray-recomputing backward must find again)"""
        return self._tex_key

    # -- closest hit ------------------------------------------------------------------------------
    def cast_rays(self, rays: torch.Tensor, normals: bool = False, stats: bool = False):
        """rays [...,6] float32 (origin, direction) -> dict of CUDA tensors shaped like rays[...,0]:
        t_hit f32 (inf on miss), primitive_ids int32 (bit pattern of Open3D's uint32; -1 on miss),
        primitive_uvs f32 [...,2], optionally primitive_normals [...,3] or traversal counters."""
        r = _as_f32(rays, self.device)
        shape = tuple(r.shape[:-1])
        if r.shape[-1] != 6:
            raise ValueError("rays must be [...,6]")
        r = r.reshape(-1, 6)
        R = r.shape[0]
        t = torch.empty(R, dtype=torch.float32, device=self.device)
        ids = torch.empty(R, dtype=torch.int32, device=self.device)
        uv = torch.empty(R, 2, dtype=torch.float32, device=self.device)
        out = {}
        lib = _lib.load()
        with torch.cuda.device(self.device):
            if stats:
                nv = torch.empty(R, dtype=torch.int32, device=self.device)
                nt = torch.empty(R, dtype=torch.int32, device=self.device)
                _lib.check(lib.texir_cast_rays_stats(self.handle, _ptr(r), R, _ptr(t), _ptr(ids), _ptr(nv), _ptr(nt),
                                                     _stream_ptr(self.device)), "texir_cast_rays_stats")
                out["nodes_visited"] = nv.reshape(shape)
                out["tris_tested"] = nt.reshape(shape)
            else:
                nrm = torch.empty(R, 3, dtype=torch.float32, device=self.device) if normals else None
                _lib.check(lib.texir_cast_rays(self.handle, _ptr(r), R, _ptr(t), _ptr(ids), _ptr(uv), _ptr(nrm),
                                               _stream_ptr(self.device)), "texir_cast_rays")
                out["primitive_uvs"] = uv.reshape(shape + (2,))
                if normals:
                    out["primitive_normals"] = nrm.reshape(shape + (3,))
        out["t_hit"] = t.reshape(shape)
        out["primitive_ids"] = ids.reshape(shape)
        return out


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        return x.numpy()
    return np.asarray(x)


class _HostTensor:
    """what Open3D returns from cast_rays as far as the reference is concerned: ``.numpy()``"""

    def __init__(self, a: np.ndarray):
        self._a = a

    def numpy(self) -> np.ndarray:
        return self._a

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    @property
    def shape(self):
        return self._a.shape


class RaycastingScene:
    """Drop-in for ``o3d.t.geometry.RaycastingScene`` at the reference's call sites
    (models/mat_nvdiffrast.py:98-101,294-295): CPU arrays in, CPU arrays out, rays traced on the GPU."""

    INVALID_ID = INVALID_ID

    def __init__(self, device=None, nthreads: int = 0):
        self._device = device
        self._threads = nthreads
        self._verts = []
        self._tris = []
        self._geom_first_prim = []
        self._scene: Optional[Scene] = None
        self._nv = 0

    def add_triangles(self, vertex_positions, triangle_indices=None) -> int:
        """add_triangles(mesh) or add_triangles(vertex_positions, triangle_indices); returns the geometry id"""
        if triangle_indices is None:
            mesh = vertex_positions
            if hasattr(mesh, "vertex") and hasattr(mesh, "triangle"):  # o3d.t.geometry.TriangleMesh
                vtx, tri = mesh.vertex, mesh.triangle
                v = vtx["positions"] if not hasattr(vtx, "positions") else vtx.positions
                t = tri["indices"] if not hasattr(tri, "indices") else tri.indices
            else:  # legacy mesh
                v, t = mesh.vertices, mesh.triangles
        else:
            v, t = vertex_positions, triangle_indices
        v = np.ascontiguousarray(_to_numpy(v), dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(_to_numpy(t), dtype=np.int64).reshape(-1, 3)
        gid = len(self._verts)
        self._geom_first_prim.append(sum(x.shape[0] for x in self._tris))
        self._tris.append((t + self._nv).astype(np.int32))
        self._verts.append(v)
        self._nv += v.shape[0]
        self._scene = None
        return gid

    @property
    def native(self) -> Scene:
        if self._scene is None:
            if not self._verts:
                raise _lib.TexirError("RaycastingScene has no geometry")
            self._scene = Scene(np.concatenate(self._verts), np.concatenate(self._tris), None, device=self._device,
                                build_threads=self._threads)
        return self._scene

    def cast_rays(self, rays, nthreads: int = 0):
        r = np.ascontiguousarray(_to_numpy(rays), dtype=np.float32)
        sc = self.native
        res = sc.cast_rays(torch.from_numpy(r), normals=True)
        ids = res["primitive_ids"].cpu().numpy().view(np.uint32)
        geom = np.full(ids.shape, INVALID_ID, np.uint32)
        if len(self._verts) == 1:
            geom[ids != INVALID_ID] = 0
            prim = ids
        else:  # Open3D numbers primitives per geometry
            first = np.asarray(self._geom_first_prim, np.int64)
            hit = ids != INVALID_ID
            g = np.searchsorted(first, ids[hit].astype(np.int64), side="right") - 1
            geom[hit] = g.astype(np.uint32)
            prim = ids.copy()
            prim[hit] = (ids[hit].astype(np.int64) - first[g]).astype(np.uint32)
        return {
            "t_hit": _HostTensor(res["t_hit"].cpu().numpy()),
            "primitive_ids": _HostTensor(prim),
            "geometry_ids": _HostTensor(geom),
            "primitive_uvs": _HostTensor(res["primitive_uvs"].cpu().numpy()),
            "primitive_normals": _HostTensor(res["primitive_normals"].cpu().numpy()),
        }
