"""The reference's config-level plug-in seam.

The reference's runners build their model as
``utils.general.get_class(conf.get_string('train.model_class'))(conf=..., ids=..., extrinsics=..., optim_cam=...)``
(trainer/train_material.py:110-111, trainer/generate_ir_texture.py:64-65, utils/general.py:12-18).  The two classes
here take exactly those constructor arguments, read the same files next to ``train.path_mesh_open3d`` the reference
classes read, and serve the same ``forward`` calls -- so a ``.conf`` that names

    train.model_class = texir_code_b200.conf_models.MaterialModel      # Mat / MatSyn stages
    train.model_class = texir_code_b200.conf_models.TracerO3d          # IrrT stage

runs under the unchanged trainers with no Open3D, nvdiffrast or pyredner in the process:

    reference ctor line (models/mat_nvdiffrast.py)                  here
    :194 pyredner.load_obj(path, obj_group=True)                    load_obj (positions, per-corner uvs, vertex normals)
    :73  cv2.imread(irt.hdr)[:,:,::-1]                              postprocess.imread_hdr (RGBE decode on the device)
    :87-93 o3d.io.read_triangle_mesh + cv2.imread(hdr_texture.hdr), flip, * 2**hdr_exposure      same files, same order
    :98-101 o3d RaycastingScene().add_triangles                     Scene (BVH built on the device)
    :82  dr.RasterizeGLContext()                                    not needed: primary rays through the same BVH

``conf`` is anything with pyhocon's accessors (get_string / get_list / get_float / get_int / get_bool with
``default=``); ConfigLite parses the HOCON subset the reference's .conf files use, for processes without pyhocon.
"""
from __future__ import annotations

import os
import re
from typing import Any, List, Optional

import numpy as np
import torch

from . import models


# ---------------------------------------------------------------------------------------------------
# a HOCON-subset reader: nested `name { ... }` blocks, `key = value`, numbers / bare words / quoted strings / lists, # comments
# ---------------------------------------------------------------------------------------------------
_MISSING = object()


class ConfigLite(dict):
    """pyhocon-style access to a nested dict: ``conf.get_string('train.path_mesh_open3d')``"""

    @staticmethod
    def _scalar(tok: str) -> Any:
        tok = tok.strip()
        if len(tok) >= 2 and tok[0] == tok[-1] and tok[0] in "\"'":
            return tok[1:-1]
        low = tok.lower()
        if low in ("true", "false"):
            return low == "true"
        try:
            return int(tok)
        except ValueError:
            pass
        try:
            return float(tok)
        except ValueError:
            return tok

    @classmethod
    def parse_string(cls, text: str) -> "ConfigLite":
        lines = []
        for raw in text.splitlines():
            line = re.sub(r"(#|//).*$", "", raw).strip()
            if line:
                lines.append(line)
        src = "\n".join(lines)
        # tokens: braces, or a `key = value` / `key` line
        root = cls()
        stack: List[dict] = [root]
        pending: Optional[str] = None
        for line in src.split("\n"):
            while line:
                line = line.strip()
                if not line:
                    break
                if line[0] == "{":
                    if pending is None:
                        raise ValueError("'{' without a block name")
                    child = stack[-1].setdefault(pending, cls())
                    stack.append(child)
                    pending = None
                    line = line[1:]
                    continue
                if line[0] == "}":
                    if len(stack) == 1:
                        raise ValueError("unbalanced '}'")
                    stack.pop()
                    line = line[1:]
                    continue
                m = re.match(r"([A-Za-z_][\w\.\-]*)\s*(=|:)\s*(.*)$", line)
                if m:
                    key, val = m.group(1), m.group(3).strip()
                    if val.startswith("{"):
                        pending = key
                        line = val
                        continue
                    if val.startswith("["):
                        close = val.rfind("]")
                        if close < 0:
                            raise ValueError("list must close on its own line: %r" % line)
                        items = [x for x in (t.strip() for t in val[1:close].split(",")) if x]
                        stack[-1][key] = [cls._scalar(x) for x in items]
                        line = val[close + 1:]
                        continue
                    brace = val.find("}")
                    if brace >= 0:
                        stack[-1][key] = cls._scalar(val[:brace])
                        line = val[brace:]
                    else:
                        stack[-1][key] = cls._scalar(val)
                        line = ""
                    continue
                m = re.match(r"([A-Za-z_][\w\.\-]*)\s*(\{.*)?$", line)
                if not m:
                    raise ValueError("cannot parse %r" % line)
                pending = m.group(1)
                line = (m.group(2) or "")
        if len(stack) != 1:
            raise ValueError("unbalanced '{'")
        return root

    @classmethod
    def parse_file(cls, path: str) -> "ConfigLite":
        with open(path) as f:
            return cls.parse_string(f.read())

    def _lookup(self, key: str, default=_MISSING):
        node: Any = self
        for part in key.split("."):
            if not isinstance(node, dict) or part not in node:
                if default is _MISSING:
                    raise KeyError(key)
                return default
            node = node[part]
        return node

    def get(self, key, default=None):  # dotted keys as well
        return self._lookup(key, default) if isinstance(key, str) and "." in key else dict.get(self, key, default)

    def get_string(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is default and default is not _MISSING else str(v)

    def get_int(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is default and default is not _MISSING else int(v)

    def get_float(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is default and default is not _MISSING else float(v)

    def get_bool(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is default and default is not _MISSING else bool(v)

    def get_list(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is default and default is not _MISSING else list(v)

    def get_config(self, key, default=_MISSING):
        return self._lookup(key, default)


# ---------------------------------------------------------------------------------------------------
# Wavefront OBJ: what the two loaders of the reference take from the file
# ---------------------------------------------------------------------------------------------------
def load_obj(path: str) -> dict:
    """vertices [V,3] f32, triangles [T,3] i32, triangle_uvs [3T,2] f64 (RAW v, per corner -- what Open3D's
    ``mesh.triangle_uvs`` holds, mat_nvdiffrast.py:97), vertex_normals [V,3] f32 or None (the `vn` of each corner stored at
    its position vertex -- pyredner's per-vertex ``normals``, :61).  Polygons are fanned; negative indices resolved."""
    vs: List[List[float]] = []
    vts: List[List[float]] = []
    vns: List[List[float]] = []
    tri_v: List[int] = []
    tri_t: List[int] = []
    tri_n: List[int] = []
    with open(path) as f:
        for line in f:
            if not line or line[0] not in "vf":
                continue
            parts = line.split()
            if not parts:
                continue
            tag = parts[0]
            if tag == "v":
                vs.append([float(x) for x in parts[1:4]])
            elif tag == "vt":
                vts.append([float(parts[1]), float(parts[2]) if len(parts) > 2 else 0.0])
            elif tag == "vn":
                vns.append([float(x) for x in parts[1:4]])
            elif tag == "f":
                corners = []
                for tok in parts[1:]:
                    ids = tok.split("/")
                    vi = int(ids[0])
                    ti = int(ids[1]) if len(ids) > 1 and ids[1] else 0
                    ni = int(ids[2]) if len(ids) > 2 and ids[2] else 0
                    corners.append((vi - 1 if vi > 0 else len(vs) + vi, ti - 1 if ti > 0 else (len(vts) + ti if ti < 0 else -1),
                                    ni - 1 if ni > 0 else (len(vns) + ni if ni < 0 else -1)))
                for k in range(1, len(corners) - 1):
                    for c in (corners[0], corners[k], corners[k + 1]):
                        tri_v.append(c[0])
                        tri_t.append(c[1])
                        tri_n.append(c[2])
    if not vs or not tri_v:
        raise ValueError("%s: no triangles" % path)
    vertices = np.asarray(vs, np.float32)
    triangles = np.asarray(tri_v, np.int32).reshape(-1, 3)
    uv_idx = np.asarray(tri_t, np.int64)
    if vts and (uv_idx >= 0).all():
        triangle_uvs = np.asarray(vts, np.float64)[uv_idx]
    else:
        triangle_uvs = np.zeros((len(tri_v), 2), np.float64)
    normals = None
    n_idx = np.asarray(tri_n, np.int64)
    if vns and (n_idx >= 0).all():
        normals = np.zeros_like(vertices)
        normals[np.asarray(tri_v, np.int64)] = np.asarray(vns, np.float32)[n_idx]
    return {"vertices": vertices, "triangles": triangles, "triangle_uvs": triangle_uvs, "vertex_normals": normals}


def _mip_level(n: int) -> int:
    """utils/general.py:88-93 get_mip_level"""
    count = 0
    while not (n & 1 or n == 1):
        n >>= 1
        count += 1
    return count


def camera_centre(mvp_face: torch.Tensor) -> torch.Tensor:
    """the eye point of a perspective mvp in the reference's row-vector convention (clip = [p,1] @ mvp): the one point
    whose clip x, y and w vanish.  The IrT generator rasterises from the matrices alone (tracer_o3d_irt.py:101-103);
    primary rays need the eye as well."""
    m = mvp_face.detach().double().cpu()
    a = m[:3, [0, 1, 3]].t()                 # three equations  p . M[:3,k] = -M[3,k]
    b = -m[3, [0, 1, 3]]
    return torch.linalg.solve(a, b).float()


def _radiance_texture(path_mesh: str, exposure: float, device) -> torch.Tensor:
    """mat_nvdiffrast.py:89-93: hdr_texture.hdr -> RGB, flipped vertically, times 2**hdr_exposure, [1,3,H,W]"""
    from .postprocess import imread_hdr
    t = imread_hdr(path_mesh.replace("out1.obj", "hdr_texture.hdr"), device=device)   # cv2 order (BGR), [H,W,3]
    t = t.flip(-1).flip(0)                                                              # [:,:,::-1], cv2.flip(., 0)
    return (t.permute(2, 0, 1).unsqueeze(0).float() * (2 ** exposure)).contiguous()


def _packed_texture(path_mesh: str, exposure: float, float_texture: torch.Tensor):
    """the same texture as the file's own 4-byte RGBE texels for Scene.set_texture_rgbe (exact for an integer exposure)"""
    if int(exposure) != exposure:
        return None
    from .postprocess import imread_hdr_rgbe
    return (imread_hdr_rgbe(path_mesh.replace("out1.obj", "hdr_texture.hdr")), int(exposure), float_texture)


class MaterialModel(models.MaterialModel):
    """models.mat_nvdiffrast.MaterialModel with the reference's constructor (mat_nvdiffrast.py:36-101)."""

    def __init__(self, conf, ids, extrinsics, optim_cam=False, gt_irf=True, gt_irrt=True, device=None, **kw):
        from .postprocess import imread_hdr
        path = conf.get_string("train.path_mesh_open3d")
        pano_res = conf.get_list("train.pano_img_res", default=[1000, 2000])
        mesh = load_obj(path)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        texture = _radiance_texture(path, conf.get_float("train.hdr_exposure"), dev) if gt_irf else torch.zeros(1, 3, 2, 2, device=dev)
        irrt = None
        if gt_irrt:
            irrt = imread_hdr(path.replace("out1.obj", "irt.hdr"), device=dev).flip(-1).contiguous()      # :73, NOT flipped
        super().__init__(mesh["vertices"], mesh["triangles"], mesh["triangle_uvs"], texture,
                         sample_l=conf.get_list("train.sample_light", default=[64, 64]),
                         sample_type=conf.get_list("models.render.sample_type", default=["uniform", "importance"]),
                         device=dev, materials_a_res=2048, materials_r_res=4096, cube_res=int(pano_res[1] / 4), irrt=irrt,
                         max_mip_level=_mip_level(8192), hdr_exposure=conf.get_float("train.hdr_exposure"),
                         vertex_normals=mesh["vertex_normals"], flip_tex_coords=True, **kw)
        if gt_irf:  # the kernels read the .hdr's own RGBE texels (bit-identical radiance, a quarter of the bytes)
            self.texture_rgbe = _packed_texture(path, conf.get_float("train.hdr_exposure"), self.texture)
        self.conf, self.ids, self.extrinsics, self.optim_cam = conf, ids, extrinsics, optim_cam
        self.path_traced_mesh, self.pano_res, self.gt_irrt = path, pano_res, gt_irrt
        self.resolution = conf.get_list("train.env_res", default=[8, 16])


class TracerO3d(models.TracerO3d):
    """models.tracer_o3d_irt.TracerO3d with the reference's constructor (tracer_o3d_irt.py:36-96) and its argument-free
    ``forward()`` (trainer/generate_ir_texture.py:77): position / normal panoramas of every view (primary rays through
    the BVH for the six cube faces, then the reference's own utils.Cube2Pano resampling, :99-112), the index-texture
    lookup (:115-142), the fused IrT launch and the seam masking (:145-180)."""

    def __init__(self, conf, ids, extrinsics, optim_cam=False, gt_irf=True, device=None):
        import cv2  # the index texture is a 16-bit PNG; the reference reads and resizes it with cv2 as well (:91-95)
        path = conf.get_string("train.path_mesh_open3d")
        mesh = load_obj(path)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        texture = _radiance_texture(path, conf.get_float("train.hdr_exposure"), dev)
        super().__init__(mesh["vertices"], mesh["triangles"], mesh["triangle_uvs"], texture,
                         sample_l=conf.get_list("train.sample_light", default=[64, 64]),
                         sample_type=conf.get_list("models.render.sample_type", default=["uniform", "importance"]), device=dev)
        self.texture_rgbe = _packed_texture(path, conf.get_float("train.hdr_exposure"), self.texture)
        self.conf, self.ids, self.extrinsics, self.optim_cam = conf, ids, extrinsics, optim_cam
        self.path_traced_mesh = path
        self.cube_res = 256
        self.vertex_normals = None if mesh["vertex_normals"] is None else torch.from_numpy(mesh["vertex_normals"]).to(dev)
        self.mesh_triangles = torch.from_numpy(mesh["triangles"].astype(np.int64)).to(dev)
        index_texture = cv2.imread(path.replace("out1.obj", "0.png"), -1)
        if index_texture is None:
            raise FileNotFoundError(path.replace("out1.obj", "0.png"))
        # the reference's call, argument for argument (:95): the third positional parameter of cv2.resize is `dst`, so
        # cv2.INTER_NEAREST is not the interpolation there either
        self.index_texture = cv2.resize(index_texture, (1024, 1024), cv2.INTER_NEAREST)

    def generate_positions(self):
        """tracer_o3d_irt.py:99-112: per view a [512,1024,6] panorama of (position + 1e-2 normal, normal)"""
        try:
            from utils.Cube2Pano import Cube2Pano          # the reference's own resampler (caller side of the boundary)
        except ImportError as exc:
            raise ImportError("TracerO3d.generate_positions needs the reference's utils.Cube2Pano on sys.path "
                              "(it runs inside the reference's trainer); otherwise pass position / normal textures to "
                              "forward() directly") from exc
        from . import ops
        c2p = Cube2Pano(pano_width=1024, pano_height=512, cube_lenth=self.cube_res, cube_channel=6, is_cuda=True)
        dev = self.scene.device
        self.position_normal_list = []
        for i in range(len(self.ids)):
            mvp = self.extrinsics[i].to(dev)
            gb = ops.gbuffer_mvp(self.scene, mvp, camera_centre(mvp[0]).to(dev), self.cube_res)
            normal = gb["normal"]
            if self.vertex_normals is not None:
                tri = self.mesh_triangles[gb["primitive_ids"].long().clamp_(min=0) % self.mesh_triangles.shape[0]]
                b = gb["primitive_uvs"]
                n = (self.vertex_normals[tri[..., 0]] * (1 - b[..., 0:1] - b[..., 1:2]) + self.vertex_normals[tri[..., 1]] * b[..., 0:1]
                     + self.vertex_normals[tri[..., 2]] * b[..., 1:2])
                normal = torch.where(gb["mask"] > 0, n, torch.tensor([1.0, 0.0, 0.0], device=dev))
            g = torch.cat([gb["position"] + 1e-2 * normal, normal], -1)                     # [6,h,w,6]
            pano = c2p.ToPano(g.permute(0, 3, 1, 2).reshape(1, -1, self.cube_res, self.cube_res))[0].permute(1, 2, 0)
            self.position_normal_list.append(pano)
        return self.position_normal_list

    def forward(self, position_texture=None, normal_texture=None, index_texture=None, shift=None):
        if position_texture is None and self.position_texture is None:
            self.generate_positions()
            self.calcute_position_normal_texture(self.position_normal_list, self.index_texture)
        return super().forward(position_texture, normal_texture, index_texture, shift)
