"""Host-side mirror of the reference's model classes, hot-path part only.

    HotPathMixin      the four methods the reference models define on this path, same names / arguments:
                      render, query_irf, specular_reflectance, diffuse_reflectance
                      (models/mat_nvdiffrast.py:201-320; relighting variant models/test_nvdiffrast.py:256-365)
    MaterialModel     state + the mixin, built from arrays instead of a HOCON conf (the conf-driven ctor of
                      the reference needs pyredner / nvdiffrast / open3d, which sit on the CALLER side of
                      the boundary; INTEGRATION.md shows the 6-line patch for the reference's own ctor)
    TracerO3d         the IrT generator's hot loop, models/tracer_o3d_irt.py:145-180, as one launch

State names follow the reference: ``scene``, ``triangle_uvs``, ``texture``, ``sample_l``, ``sample_type``,
``relighting``.  ``sample_l`` is read at call time because the trainer mutates it
(trainer/train_material_syn.py:735).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops, sample_util
from .sample_util import draw_shift, generate_dir
from .scene import Scene


def _upload_texture(model, sc: Scene) -> None:
    """makes ``model.texture`` the scene's resident radiance texture.  ``model.texture_rgbe = (rgbe uint8 [H,W,4],
    exposure, float_texture)`` (opt-in) keeps it as the 4-byte RGBE texels of the .hdr file it came from -- used as long
    as model.texture still IS that float texture (stage -1 swaps in a thresholded copy, mat_nvdiffrast.py:143-150, which
    goes up as floats)."""
    packed = getattr(model, "texture_rgbe", None)
    if packed is not None and model.texture is packed[2]:
        sc.set_texture_rgbe(packed[0], packed[1], flip=True)
    else:
        sc.set_texture(model.texture)


class HotPathMixin:
    """Expects on ``self``: scene (texir_code_b200.Scene built WITH triangle_uvs), texture (the reference's
    [1,3,H,W] tensor; re-uploaded when it is swapped or modified in place), sample_l, sample_type, and
    optionally relighting (bool), brdf_eps (1e-14 for mat_nvdiffrast, 1e-6 for test_nvdiffrast and the
    baseline variants), fused (bool, default True) and backward_mode ('stash' | 'recompute')."""

    # -- helpers ----------------------------------------------------------------------------------
    def _scene(self) -> Scene:
        sc = self.scene
        if not isinstance(sc, Scene):
            raise TypeError("HotPathMixin needs self.scene to be a texir_code_b200.Scene built with triangle_uvs")
        _upload_texture(self, sc)
        return sc

    def _draw_shift(self, b: int, device) -> torch.Tensor:
        """the Cranley-Patterson shift of one generate_dir call: torch.rand(b, 1, 2) on the CPU generator, as the
        reference draws it (utils/sample_util.py:102).  ``self.shift_rng = 'device'`` (opt-in) draws it from the CUDA
        generator instead: a different, equally distributed sample set, without the ~0.5 ms of host time that
        torch.rand(98304, 1, 2) costs every step -- no longer the reference's stream for a given seed."""
        if getattr(self, "shift_rng", "reference") == "device":
            return torch.rand(b, 1, 2, device=device)
        return draw_shift(b, device)

    # -- reference methods ------------------------------------------------------------------------
    def query_irf(self, points, directions, num_sample):
        """mat_nvdiffrast.py:292 -- points [P,N,3], directions [P,N,1,3] -> [P,N,3] (CUDA)"""
        return ops.query_irf(self._scene(), points, directions, num_sample)

    def diffuse_reflectance(self, lighting, l, n, albedo, sample_type="uniform"):
        """mat_nvdiffrast.py:252"""
        return ops.diffuse_reflectance(lighting, l, n, albedo, sample_type)

    def specular_reflectance(self, lighting, h, n, v, l, roughness, albedo=None):
        """mat_nvdiffrast.py:260"""
        return ops.specular_reflectance(lighting, h, n, v, l, roughness, albedo, eps=getattr(self, "brdf_eps", 1e-14))

    def render(self, normal: torch.Tensor, albedo: torch.Tensor, roughness: torch.Tensor, points: torch.Tensor,
               cam_position: torch.Tensor, irr: Optional[torch.Tensor]):
        """mat_nvdiffrast.py:201 -- inputs [6,h,w,c] (any leading shape), cam_position [3].
        Returns the reference's dict: rgb, albedo, normal (detached), position = p + 2e-2 n (detached)."""
        lead = normal.shape[:-1]
        normal = normal.reshape(-1, 3)
        albedo = albedo.reshape(-1, 3)
        roughness = roughness.reshape(-1, 1)
        points = points.reshape(-1, 3)
        irr = irr.reshape(-1, 3) if irr is not None else None
        relight = bool(getattr(self, "relighting", False))
        eps = getattr(self, "brdf_eps", 1e-14)
        n_diff, n_spec = int(self.sample_l[0]), int(self.sample_l[1])
        if getattr(self, "fused", True):
            b, dev = normal.shape[0], normal.device
            # RNG order of the reference: the diffuse generate_dir (if any) draws first
            shift_diff = self._draw_shift(b, dev) if relight else None
            shift_spec = self._draw_shift(b, dev)
            rgb = ops.shade(self._scene(), normal, albedo, roughness, points, cam_position, irr, shift_spec,
                            num_spec=n_spec, shift_diff=shift_diff, num_diff=n_diff if relight else 0,
                            mode_spec=self.sample_type[1], mode_diff=self.sample_type[0], eps=eps,
                            backward=getattr(self, "backward_mode", "stash"))
        else:  # op-by-op composition, exactly the reference's sequence of calls
            view = torch.nn.functional.normalize(cam_position.unsqueeze(0) - points, eps=1e-4)
            if relight:
                l_diff = generate_dir(normal, n_diff, None, mode=self.sample_type[0])
                rad = self.query_irf(points.unsqueeze(1).expand_as(l_diff), l_diff.unsqueeze(-2), n_diff)
                diffuse = self.diffuse_reflectance(rad, l_diff, normal, albedo, self.sample_type[0]) / n_diff
            else:
                diffuse = irr * albedo / np.pi
            h = generate_dir(normal, n_spec, None, self.sample_type[1], roughness)
            vdh = torch.clamp(torch.sum(h * view.unsqueeze(1), dim=-1, keepdim=True), 0.0, 1.0)
            l_spec = 2 * vdh * h - view.unsqueeze(1)
            rad = self.query_irf(points.unsqueeze(1).expand_as(l_spec), l_spec.unsqueeze(-2).detach(), n_spec)
            rgb = diffuse + self.specular_reflectance(rad, h, normal, view, l_spec, roughness) / n_spec
        return {
            "rgb": rgb.reshape(*lead, 3),
            "albedo": albedo.reshape(*lead, 3),
            "normal": normal.reshape(*lead, 3).detach(),
            "position": (points + 2e-2 * normal).reshape(*lead, 3).detach(),
        }

    # north-star aliases (BASELINE.json names them trace / shade)
    trace = query_irf
    shade = render


def rays_from_mvp(mvp: torch.Tensor, cam_position: torch.Tensor, res: int):
    """Primary rays of the views the reference rasterises (models/mat_nvdiffrast.py:116-119: clip = [v,1] @ mvp[i],
    dr.rasterize at res x res), one per pixel centre, row 0 at ndc y = -1 as in nvdiffrast's images.
    mvp [V,4,4] (datasets/dataset.py:461-463: transpose(proj @ w2c)), cam_position [3]  ->
    rays [V,res,res,6] (origin, unit direction) and ray differentials [V,res,res,6] = (d dir/dX, d dir/dY) per
    pixel step (X = column, Y = row).  Plain torch in float64: the readable statement of what the one-launch kernel
    behind ops.gbuffer_mvp computes per pixel (the model uses that kernel; tests compare the two)."""
    dev = mvp.device
    inv = torch.linalg.inv(mvp.detach().double())                      # world_h = clip_h @ inv
    cam = cam_position.detach().double().to(dev)
    lin = (torch.arange(res, dtype=torch.float64, device=dev) + 0.5) / res * 2.0 - 1.0
    y, x = torch.meshgrid(lin, lin, indexing="ij")
    clip = torch.stack([x, y, torch.ones_like(x), torch.ones_like(x)], -1)            # far plane, w = 1
    wh = torch.einsum("hwk,vkl->vhwl", clip, inv)
    H3, w = wh[..., :3], wh[..., 3:4]
    d = H3 / w - cam
    ln = torch.linalg.norm(d, dim=-1, keepdim=True)
    dh = d / ln
    cols = []
    for axis in (0, 1):                                                # d/d(ndc x), d/d(ndc y)
        row = inv[:, axis, :].reshape(-1, 1, 1, 4)
        dP = (row[..., :3] * w - H3 * row[..., 3:4]) / (w * w) * (2.0 / res)
        cols.append((dP - dh * (dh * dP).sum(-1, keepdim=True)) / ln)
    rays = torch.cat([cam.expand_as(dh), dh], -1).float().contiguous()
    return rays, torch.cat(cols, -1).float().contiguous()


def rgb_to_intensity(t: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """utils/general.py:95-111 (dim = -1 and dim = 1 are the forms the hot path uses)"""
    sl = lambda k: t.narrow(dim if dim >= 0 else t.dim() - 1, k, 1)
    return 0.29900 * sl(0) + 0.58700 * sl(1) + 0.11400 * sl(2)


class MaterialModel(HotPathMixin, nn.Module):
    """models.mat_nvdiffrast.MaterialModel (and, with relighting=True / brdf_eps=1e-6, the hot path of
    models.test_nvdiffrast.MaterialModel), built from arrays.  ``forward(mvp, id, cam_position, stage)`` is the
    reference's (mat_nvdiffrast.py:105-191) with the OpenGL rasteriser and dr.interpolate replaced by primary rays
    through the scene BVH (ops.gbuffer) and dr.texture by ops.texture_fetch."""

    def __init__(self, vertices, triangles, triangle_uvs, texture, sample_l: Sequence[int] = (32, 64),
                 sample_type: Sequence[str] = ("uniform", "importance"), relighting: bool = False,
                 brdf_eps: float = 1e-14, device=None, materials_a_res: int = 0, materials_r_res: int = 0,
                 fused: bool = True, backward_mode: str = "stash", cube_res: int = 128, irrt=None,
                 max_mip_level: Optional[int] = 13, hdr_exposure: float = 5.0, vertex_normals=None,
                 scene: Optional[Scene] = None, flip_tex_coords: bool = True, shard=None, cache_gbuffers: bool = False,
                 shift_rng: str = "reference", grad_sync: str = "allreduce"):
        super().__init__()
        if grad_sync not in ("allreduce", "gather"):
            raise ValueError("grad_sync must be 'allreduce' or 'gather'")
        # sharded model only.  'allreduce': the backward leaves this rank's partial texture gradients and the caller runs
        # parallel.allreduce_gradients (ONE all-reduce of 112 MiB; replicas stay bit-identical).  'gather': the backward
        # all-gathers the per-pixel gradients (4.3 MB) and every rank scatters all of them -- the full gradients without an
        # all-reduce of the textures (float atomics arrive in a different order on each rank: equal up to rounding).
        self.grad_sync = grad_sync
        if shift_rng not in ("reference", "device"):
            raise ValueError("shift_rng must be 'reference' (CPU generator, the reference's stream) or 'device'")
        self.shift_rng = shift_rng
        # `scene`: share an already built Scene (BVH + uploads) instead of building another one from the arrays
        self.scene = scene if scene is not None else Scene(vertices, triangles, triangle_uvs, device=device)
        self.triangle_uvs = np.asarray(triangle_uvs)
        self.texture = texture if isinstance(texture, torch.Tensor) else torch.from_numpy(np.asarray(texture))
        self.sample_l = list(sample_l)
        self.sample_type = list(sample_type)
        self.relighting = relighting
        self.brdf_eps = brdf_eps
        self.fused = fused
        self.backward_mode = backward_mode
        self.cube_res = int(cube_res)
        self.max_mip_level = max_mip_level      # get_mip_level(8192) = 13, mat_nvdiffrast.py:65
        self.hdr_exposure = float(hdr_exposure)
        # The reference addresses materials_a / materials_r / irrt through self.uvs, which pyredner.load_obj delivers
        # with flip_tex_coords=True (v' = 1 - v; mat_nvdiffrast.py:194,128-139), while the Open3D ray caster keeps the
        # OBJ's raw v against a vertically flipped radiance texture (:89-90).  `scene` holds the raw per-corner UVs, so
        # the texture fetches of forward() see (u, 1 - v); irt.hdr is loaded unflipped (:73) for the same reason.
        self.flip_tex_coords = bool(flip_tex_coords)
        # (rank, world): render only this rank's pixels (parallel.local_pixels gives the matching slice of an image);
        # forward() then returns flat [P_local, C] tensors and parallel.allreduce_gradients completes the step
        self.shard = None if shard is None else (int(shard[0]), int(shard[1]))
        self.mat_ws = ops.MatWorkspace()
        self._warned_normals = False
        # opt-in: keep the G-buffer of a view between steps.  The reference rasterises every step although geometry and
        # cameras are constant (optim_cam is off in every shipped conf); a cached entry is reused only when forward() is
        # called with the SAME mvp / cam_position tensors (identity and in-place version) for the same `id`, i.e. when
        # the caller uploads the per-view matrices once -- nothing is ever assumed about tensors it re-creates
        self.cache_gbuffers = bool(cache_gbuffers)
        self._gb_cache = {}
        dev = self.scene.device
        # trainable material textures, same names / initial values as mat_nvdiffrast.py:68-69
        if materials_a_res > 0:
            self.materials_a = nn.Parameter(torch.full((materials_a_res, materials_a_res, 3), 0.5, device=dev))
        if materials_r_res > 0:
            self.materials_r = nn.Parameter(torch.full((materials_r_res, materials_r_res, 1), 0.1, device=dev))
        if irrt is not None:                    # the padded / denoised IrT, mat_nvdiffrast.py:72-76
            self.irrt = nn.Parameter(torch.as_tensor(irrt, dtype=torch.float32).to(dev), requires_grad=False)
        if vertex_normals is not None:          # the reference interpolates the OBJ's vertex normals (:120)
            self.normals = nn.Parameter(torch.as_tensor(vertex_normals, dtype=torch.float32).to(dev), requires_grad=False)
            self.triangles = nn.Parameter(torch.as_tensor(np.asarray(triangles), dtype=torch.int64).to(dev),
                                          requires_grad=False)

    def _draw_shift(self, b: int, device) -> torch.Tensor:
        """sharded model: every rank draws the shift of ALL 6 * cube_res^2 pixels (so the CPU generator advances as on one
        GPU and a pixel gets the same shift whatever the world size) and keeps its own pixels'"""
        total = 6 * self.cube_res * self.cube_res
        if self.shard is None or self.shard[1] == 1 or b * self.shard[1] != total:
            return HotPathMixin._draw_shift(self, b, device)
        from .parallel import local_pixels
        if getattr(self, "shift_rng", "reference") == "device":   # (ranks seeded alike draw the same full set)
            full = torch.rand(total, 1, 2, device=device)
            return local_pixels(full.reshape(6, -1, 2), self.shard[0], self.shard[1], 6).reshape(-1, 1, 2)
        full = draw_shift(total, device)
        return local_pixels(full.reshape(6, -1, 2), self.shard[0], self.shard[1], 6).reshape(-1, 1, 2)

    def forward_gbuffer(self, normal, points, cam_position, irr, albedo=None, roughness=None, texc=None, texd=None):
        """G-buffer in, shaded colour out.  Materials come either as per-pixel tensors or, given texc [P,2] (and texd
        [P,4] for the mip-mapped filter), from the trainable textures through the fetch whose backward scatters
        into them."""
        if texc is not None:
            flat = texc.reshape(-1, 2)
            da = texd.reshape(-1, 4) if texd is not None else None
            albedo = ops.texture_fetch(self.materials_a, flat, da, max_mip_level=self.max_mip_level)
            roughness = ops.texture_fetch(self.materials_r, flat, da, max_mip_level=self.max_mip_level)
            albedo = albedo.reshape(*normal.shape[:-1], 3)
            roughness = roughness.reshape(*normal.shape[:-1], 1)
        return self.render(normal, albedo, roughness, points, cam_position, irr)

    def _render_textured(self, normal, albedo, roughness, roughness_womipmap, points, cam_position, irr, texc, texd, stage,
                         shift_spec=None):
        """stages 1 and 2 of forward(): the result dict of render() plus roughness / roughness_womipmap.  Fused: one
        autograd node does the four fetches, the render and -- in its backward -- the scatter of every output's gradient
        into the material textures (ops.mat_stage).  Not fused (or relighting): the reference's own composition."""
        lead = normal.shape[:-1]
        if albedo is not None:
            if stage == 1:
                res = self.render(normal, albedo.detach(), roughness_womipmap, points, cam_position, irr)
            else:
                res = self.render(normal, albedo, roughness, points, cam_position, irr)
            return res, roughness, roughness_womipmap
        n, p = normal.reshape(-1, 3), points.reshape(-1, 3)
        if shift_spec is None:
            shift_spec = self._draw_shift(n.shape[0], n.device)
        elif shift_spec.shape[0] != n.shape[0]:
            raise RuntimeError("the shift drawn ahead of the G-buffer covers %d pixels, the G-buffer has %d"
                               % (shift_spec.shape[0], n.shape[0]))
        rgb, alb, rough, rough_w, _ = ops.mat_stage(
            self._scene(), self.mat_ws, self.materials_a, self.materials_r, self.irrt, n, p, cam_position, texc, texd,
            shift_spec, stage=stage, num_spec=int(self.sample_l[1]), max_mip_level=self.max_mip_level,
            mode_spec=self.sample_type[1], eps=getattr(self, "brdf_eps", 1e-14),
            backward=getattr(self, "backward_mode", "stash"),
            gather_group=(True, None) if (self.shard is not None and self.shard[1] > 1 and self.grad_sync == "gather") else None)
        alb = alb.reshape(*lead, 3)
        res = {"rgb": rgb.reshape(*lead, 3), "albedo": alb.detach() if stage == 1 else alb, "normal": normal.detach(),
               "position": (points + 2e-2 * normal).detach()}
        return res, rough.reshape(*lead, 1), rough_w.reshape(*lead, 1)

    def gbuffers(self, mvp: torch.Tensor, cam_position: torch.Tensor, id=None) -> dict:
        """mat_nvdiffrast.py:116-128: position / normal / texc / texd / mask of the 6 cube faces"""
        if self.cache_gbuffers and id is not None:
            try:
                hash(id)
            except TypeError:
                id = int(id)
            key = (mvp.data_ptr(), mvp._version, cam_position.data_ptr(), cam_position._version, self.shard, self.cube_res)
            hit = self._gb_cache.get(id)
            if hit is not None and hit[0] == key:
                return hit[2]
            gb = self._gbuffers(mvp, cam_position)
            if len(self._gb_cache) > 4096:
                self._gb_cache.clear()
            self._gb_cache[id] = (key, (mvp, cam_position), gb)   # the tensors are kept so their addresses stay theirs
            return gb
        return self._gbuffers(mvp, cam_position)

    def _gbuffers(self, mvp: torch.Tensor, cam_position: torch.Tensor) -> dict:
        dev = self.scene.device
        gb = ops.gbuffer_mvp(self._scene(), mvp.to(dev), cam_position.to(dev), self.cube_res, shard=self.shard,
                             flip_v=self.flip_tex_coords)
        if not hasattr(self, "normals") and not self._warned_normals:
            import warnings
            warnings.warn("MaterialModel was built without vertex_normals: the G-buffer normal is the geometric face normal "
                          "turned towards the camera, whereas the reference interpolates the OBJ's vertex normals "
                          "(mat_nvdiffrast.py:120); pass vertex_normals= for meshes with smooth shading normals")
            self._warned_normals = True
        if hasattr(self, "normals"):
            tri = self.triangles[gb["primitive_ids"].long().clamp_(min=0) % self.triangles.shape[0]]
            b = gb["primitive_uvs"]
            n = (self.normals[tri[..., 0]] * (1 - b[..., 0:1] - b[..., 1:2]) + self.normals[tri[..., 1]] * b[..., 0:1]
                 + self.normals[tri[..., 2]] * b[..., 1:2])
            fixed = torch.tensor([1.0, 0.0, 0.0], device=dev)
            gb["normal"] = torch.where(gb["mask"] > 0, n, fixed)
        return gb

    def forward(self, mvp, id, cam_position, stage=1):
        """mat_nvdiffrast.py:105-191, same arguments and result keys (rgb, albedo, normal, position, empty_mask,
        roughness_womipmap, roughness); stage -1 / 0 / 1 / 2 as in the reference."""
        # a device-side shift draw hands torch's CPU generator its advanced state when this call returns, not in the
        # middle of the launches (nothing in here reads that generator in between)
        with sample_util.deferred_rng_commit():
            return self._forward(mvp, id, cam_position, stage)

    def _forward(self, mvp, id, cam_position, stage):
        for name in ("materials_a", "materials_r", "irrt"):
            if not hasattr(self, name):
                raise ValueError("MaterialModel.forward needs %s (constructor: materials_a_res / materials_r_res / irrt)" % name)
        dev = self.scene.device
        fused = getattr(self, "fused", True) and not getattr(self, "relighting", False)
        shift_spec = None
        if stage in (1, 2) and fused:
            # neither the shift of the step's one generate_dir call nor the pyramids of the textures depend on the view:
            # the shift kernel (its own stream) and the pyramid builds (side stream) run beside the G-buffer walk.  Nothing
            # else draws from the generator before the render, so drawing first leaves the reference's stream intact.
            world = self.shard[1] if self.shard is not None else 1
            shift_spec = self._draw_shift(int(mvp.shape[0]) * self.cube_res * self.cube_res // world, dev)
            self.mat_ws.prefetch({"a": self.materials_a, "r": self.materials_r, "i": self.irrt}, self.max_mip_level)
        gb = self.gbuffers(mvp, cam_position, id)       # (keyed on the caller's own tensors when cache_gbuffers is on)
        cam_position = cam_position.to(dev)
        normal, pos, mask = gb["normal"].detach(), gb["position"].detach(), gb["mask"]
        shp = normal.shape[:-1]
        texc, texd = gb["texc"].reshape(-1, 2), gb["texd"].reshape(-1, 4)
        if stage in (1, 2) and fused:
            # fetches, render and the texture-space backward are ONE autograd node (ops.mat_stage)
            albedo = roughness = roughness_womipmap = irr = None
        else:
            mip = dict(filter_mode="linear-mipmap-linear", max_mip_level=self.max_mip_level)
            albedo = ops.texture_fetch(self.materials_a, texc, texd, **mip).reshape(*shp, 3)
            roughness_womipmap = ops.texture_fetch(self.materials_r, texc, None, "linear").reshape(*shp, 1)
            roughness = ops.texture_fetch(self.materials_r, texc, texd, **mip).reshape(*shp, 1)
            irr = ops.texture_fetch(self.irrt, texc, texd, **mip).reshape(*shp, 3)
        if stage == -1:
            # "a simple trick to locate the light source": only texels brighter than 0.5 (before exposure) emit
            source = self.texture
            t = source.to(dev) if isinstance(source, torch.Tensor) else torch.as_tensor(source).to(dev)
            intensity = rgb_to_intensity(t * (2 ** -self.hdr_exposure), dim=1)
            self.texture = torch.where(intensity >= 0.5, t, torch.zeros_like(t))
            try:
                res = self.render(normal, torch.zeros_like(albedo), torch.ones_like(roughness) * 0.01, pos + 1e-2 * normal,
                                  cam_position, irr)
            finally:
                self.texture = source
        elif stage == 0:
            res = {"rgb": irr * albedo / np.pi, "albedo": albedo, "normal": normal, "position": pos + 1e-1 * normal}
        elif stage in (1, 2):
            res, roughness, roughness_womipmap = self._render_textured(normal, albedo, roughness, roughness_womipmap,
                                                                       pos + 1e-2 * normal, cam_position, irr, texc, texd,
                                                                       stage, shift_spec)
        else:
            raise ValueError("stage must be -1, 0, 1 or 2")
        res.update({"empty_mask": mask, "roughness_womipmap": roughness_womipmap, "roughness": roughness})
        return res


def _seam_mask(index_texture, device) -> torch.Tensor:
    """tracer_o3d_irt.py:138,177: texels whose index entry sums to zero -- summed in the array's OWN dtype, like the
    reference's numpy does (the index texture is a 16-bit PNG, so the sum wraps modulo 65536 there too)"""
    if isinstance(index_texture, torch.Tensor):
        idx = index_texture
        return ((idx[..., 0] + idx[..., 1] + idx[..., 2]) == 0).to(device)
    idx = np.asarray(index_texture)
    return torch.from_numpy(np.ascontiguousarray((idx[:, :, 0] + idx[:, :, 1] + idx[:, :, 2]) == 0)).to(device)


class TracerO3d(nn.Module):
    """IrT generator hot loop (models/tracer_o3d_irt.py:156-180).  The reference walks the 1024^2 texels in
    2048 batches of 512; here all texels go through ONE launch of the fused kernel.  The shift stream is
    the reference's: consecutive ``torch.rand(512,1,2)`` draws equal one ``torch.rand(P,1,2)`` draw."""

    batch = 512

    def __init__(self, vertices, triangles, triangle_uvs, texture, sample_l: Sequence[int] = (2048, 64),
                 sample_type: Sequence[str] = ("uniform", "importance"), device=None, build_threads: int = 0):
        super().__init__()
        self.scene = Scene(vertices, triangles, triangle_uvs, device=device, build_threads=build_threads)
        self.triangle_uvs = np.asarray(triangle_uvs)
        self.texture = texture if isinstance(texture, torch.Tensor) else torch.from_numpy(np.asarray(texture))
        self.sample_l = list(sample_l)
        self.sample_type = list(sample_type)
        self.position_texture = None    # set by calcute_position_normal_texture, read by forward()
        self.normal_texture = None
        self.index_texture = None

    def calcute_position_normal_texture(self, position_normal_list, index_texture):
        """tracer_o3d_irt.py:115-142 (the reference's spelling): every texel of the index texture [H,W,3] = (row, col,
        view) in units of 1/50000 of the panorama picks its position and normal from that view's [h,w,6]
        position-normal panorama; texels whose entry is all zero (seams) get zeros.  Same integer arithmetic as the
        reference's numpy, done on the device."""
        dev = self.scene.device
        idx = torch.as_tensor(np.asarray(index_texture).astype(np.int64), device=dev)
        H, W = idx.shape[:2]
        pos = torch.zeros(H, W, 3, dtype=torch.float32, device=dev)
        nrm = torch.zeros(H, W, 3, dtype=torch.float32, device=dev)
        for view in torch.unique(idx[..., 2]).tolist():
            pano = torch.as_tensor(position_normal_list[view], dtype=torch.float32, device=dev)
            h, w = pano.shape[:2]
            sel = idx[..., 2] == view
            # (index / 50000 * size).astype(int): float64 division, truncation, then np.clip
            col = (idx[..., 1][sel].double() / 50000 * w).long().clamp_(0, w - 1)
            row = (idx[..., 0][sel].double() / 50000 * h).long().clamp_(0, h - 1)
            pos[sel] = pano[row, col, 0:3]
            nrm[sel] = pano[row, col, 3:6]
        seam = _seam_mask(index_texture, dev)
        pos[seam] = 0.0
        nrm[seam] = 0.0
        self.position_texture, self.normal_texture, self.index_texture = pos, nrm, index_texture
        return pos, nrm

    def query_irf(self, points, directions, num_sample):
        """tracer_o3d_irt.py:240"""
        _upload_texture(self, self.scene)
        return ops.query_irf(self.scene, points, directions, num_sample)

    def irradiance(self, position: torch.Tensor, normal: torch.Tensor, shift: Optional[torch.Tensor] = None):
        """position, normal [P,3] -> E [P,3]; one fused launch"""
        _upload_texture(self, self.scene)
        P = position.shape[0]
        if shift is None:
            shift = draw_shift(P, position.device)
        mode = self.sample_type[0]
        if mode == "importance":
            raise ValueError("IrT sampling must be 'uniform' or 'cosine'")
        return ops.irradiance(self.scene, position, normal, shift, int(self.sample_l[0]), mode)

    def forward(self, position_texture: Optional[torch.Tensor] = None, normal_texture: Optional[torch.Tensor] = None,
                index_texture=None, shift: Optional[torch.Tensor] = None):
        """position_texture, normal_texture [H,W,3] (CUDA) -> irradiance texture [H,W,3]; texels whose
        index-texture entry is all zero (seams) are zeroed (tracer_o3d_irt.py:176-178).  Called without arguments
        like the reference's runner does (trainer/generate_ir_texture.py:77: ``self.model()``) it uses the textures
        calcute_position_normal_texture() left on the model."""
        if position_texture is None or normal_texture is None:
            if self.position_texture is None or self.normal_texture is None:
                raise ValueError("TracerO3d.forward(): no position / normal textures -- pass them, or call "
                                 "calcute_position_normal_texture(position_normal_list, index_texture) first")
            position_texture, normal_texture = self.position_texture, self.normal_texture
            if index_texture is None:
                index_texture = self.index_texture
        shp = position_texture.shape
        pos = position_texture.reshape(-1, 3)
        nrm = normal_texture.reshape(-1, 3)
        if pos.shape[0] % self.batch != 0:
            raise AssertionError("texel count must be a multiple of 512 (tracer_o3d_irt.py:161)")
        irr = self.irradiance(pos, nrm, shift).reshape(shp)
        if index_texture is not None:
            irr = irr.masked_fill(_seam_mask(index_texture, irr.device).unsqueeze(-1), 0.0)
        return irr
