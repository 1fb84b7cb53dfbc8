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

from . import ops
from .sample_util import draw_shift, generate_dir
from .scene import Scene


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
        sc.set_texture(self.texture)
        return sc

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
            shift_diff = draw_shift(b, dev) if relight else None
            shift_spec = draw_shift(b, dev)
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


class MaterialModel(HotPathMixin, nn.Module):
    """Hot-path state of models.mat_nvdiffrast.MaterialModel (and, with relighting=True / brdf_eps=1e-6,
    of models.test_nvdiffrast.MaterialModel), built from arrays."""

    def __init__(self, vertices, triangles, triangle_uvs, texture, sample_l: Sequence[int] = (32, 64),
                 sample_type: Sequence[str] = ("uniform", "importance"), relighting: bool = False,
                 brdf_eps: float = 1e-14, device=None, materials_a_res: int = 0, materials_r_res: int = 0,
                 fused: bool = True, backward_mode: str = "stash"):
        super().__init__()
        self.scene = Scene(vertices, triangles, triangle_uvs, device=device)
        self.triangle_uvs = np.asarray(triangle_uvs)
        self.texture = texture if isinstance(texture, torch.Tensor) else torch.from_numpy(np.asarray(texture))
        self.sample_l = list(sample_l)
        self.sample_type = list(sample_type)
        self.relighting = relighting
        self.brdf_eps = brdf_eps
        self.fused = fused
        self.backward_mode = backward_mode
        dev = self.scene.device
        # trainable material textures, same names / initial values as mat_nvdiffrast.py:68-69
        if materials_a_res > 0:
            self.materials_a = nn.Parameter(torch.full((materials_a_res, materials_a_res, 3), 0.5, device=dev))
        if materials_r_res > 0:
            self.materials_r = nn.Parameter(torch.full((materials_r_res, materials_r_res, 1), 0.1, device=dev))

    def forward(self, normal, points, cam_position, irr, albedo=None, roughness=None, texc=None):
        """G-buffer in, shaded colour out.  Materials come either as per-pixel tensors or, given texc [P,2],
        from the trainable textures through the bilinear fetch whose backward scatters into them."""
        if texc is not None:
            flat = texc.reshape(-1, 2)
            albedo = ops.texture_fetch(self.materials_a, flat).reshape(*normal.shape[:-1], 3)
            roughness = ops.texture_fetch(self.materials_r, flat).reshape(*normal.shape[:-1], 1)
        return self.render(normal, albedo, roughness, points, cam_position, irr)


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

    def query_irf(self, points, directions, num_sample):
        """tracer_o3d_irt.py:240"""
        self.scene.set_texture(self.texture)
        return ops.query_irf(self.scene, points, directions, num_sample)

    def irradiance(self, position: torch.Tensor, normal: torch.Tensor, shift: Optional[torch.Tensor] = None):
        """position, normal [P,3] -> E [P,3]; one fused launch"""
        self.scene.set_texture(self.texture)
        P = position.shape[0]
        if shift is None:
            shift = draw_shift(P, position.device)
        mode = self.sample_type[0]
        if mode == "importance":
            raise ValueError("IrT sampling must be 'uniform' or 'cosine'")
        return ops.irradiance(self.scene, position, normal, shift, int(self.sample_l[0]), mode)

    def forward(self, position_texture: torch.Tensor, normal_texture: torch.Tensor, index_texture=None,
                shift: Optional[torch.Tensor] = None):
        """position_texture, normal_texture [H,W,3] (CUDA) -> irradiance texture [H,W,3]; texels whose
        index-texture entry is all zero (seams) are zeroed (tracer_o3d_irt.py:176-178)."""
        shp = position_texture.shape
        pos = position_texture.reshape(-1, 3)
        nrm = normal_texture.reshape(-1, 3)
        if pos.shape[0] % self.batch != 0:
            raise AssertionError("texel count must be a multiple of 512 (tracer_o3d_irt.py:161)")
        irr = self.irradiance(pos, nrm, shift).reshape(shp)
        if index_texture is not None:
            idx = torch.as_tensor(index_texture, device=irr.device)
            seam = (idx[..., 0] + idx[..., 1] + idx[..., 2]) == 0
            irr = irr.masked_fill(seam.unsqueeze(-1), 0.0)
        return irr
