"""Tensor-in / tensor-out operators of the hot path (CUDA float32, autograd-aware).

Each one is a thin torch.autograd.Function over one entry point of libtexir_b200.so:

    query_irf              models/mat_nvdiffrast.py:292-320, models/tracer_o3d_irt.py:240-269
    specular_reflectance   models/mat_nvdiffrast.py:260-279 (eps 1e-14); models/test_nvdiffrast.py:318-338 (1e-6)
    diffuse_reflectance    models/mat_nvdiffrast.py:252-258
    shade                  MaterialModel.render, models/mat_nvdiffrast.py:201-249 -- ONE fused kernel
                           (relighting branch models/test_nvdiffrast.py:275-295 when N_diff > 0)
    irradiance             the IrT loop body, models/tracer_o3d_irt.py:165-173 -- ONE fused kernel
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib
from .scene import Scene, _on, _ptr, _stream_ptr

TINY_NUMBER = 1e-6
TINY_TINY_NUMBER = 1e-14


def _f32c(t: torch.Tensor, dev=None) -> torch.Tensor:
    t = t.detach()
    if t.dtype is torch.float32 and (dev is None or t.device == dev) and t.is_contiguous():
        return t                      # the common case: three dispatcher calls saved per tensor, ~15 tensors per step
    if dev is not None and t.device != dev:
        t = t.to(dev)
    return t.to(torch.float32).contiguous()


def _need_cuda(name: str, t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.TexirError("%s: tensors must live on a CUDA device (there is no CPU path)" % name)


# --------------------------------------------------------------------------------------------------
# query_irf
# --------------------------------------------------------------------------------------------------
def query_irf(scene: Scene, points: torch.Tensor, directions: torch.Tensor, num_sample: Optional[int] = None
              ) -> torch.Tensor:
    """points [P,N,3] (the reference passes ``points.unsqueeze(1).expand_as(dirs)``; [P,3] also accepted),
    directions [P,N,1,3] or [P,N,3] -> radiance [P,N,3].  Not differentiable (the reference's goes
    through numpy)."""
    _need_cuda("query_irf", directions)
    dev = directions.device
    d = directions
    if d.dim() == 4:
        d = d.reshape(d.shape[0], d.shape[1], 3)
    P, N = int(d.shape[0]), int(d.shape[1])
    if num_sample is not None and int(num_sample) != N:
        raise ValueError("num_sample (%d) does not match directions (%d)" % (num_sample, N))
    if points.dim() == 3:
        if points.stride(1) == 0 or N == 1:
            org = points[:, 0, :]
        else:
            org = points[:, 0, :]
            if not torch.equal(points, org.unsqueeze(1).expand_as(points)):
                # per-sample origins: trace them as N=1 "points"
                return query_irf(scene, points.reshape(P * N, 3), d.reshape(P * N, 1, 3), 1).reshape(P, N, 3)
    else:
        org = points
    org = _f32c(org, dev)
    d = _f32c(d, dev)
    rad = torch.empty(P, N, 3, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.check(_lib.load().texir_query_irf(scene.handle, _ptr(org), _ptr(d), P, N, _ptr(rad), _stream_ptr(dev)),
                   "texir_query_irf")
    return rad


# --------------------------------------------------------------------------------------------------
# stand-alone BRDF reductions
# --------------------------------------------------------------------------------------------------
class _Specular(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lighting, h, n, v, l, roughness, eps):
        dev = lighting.device
        P, N = int(lighting.shape[0]), int(lighting.shape[1])
        L, hh, ll = _f32c(lighting), _f32c(h), _f32c(l)
        nn, vv = _f32c(n), _f32c(v)
        rr = _f32c(roughness).reshape(-1)
        out = torch.empty(P, 3, dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.check(_lib.load().texir_specular_reflectance(_ptr(L), _ptr(hh), _ptr(nn), _ptr(vv), _ptr(ll), _ptr(rr),
                                                              P, N, eps, _ptr(out), _stream_ptr(dev)),
                       "texir_specular_reflectance")
        ctx.save_for_backward(L, hh, nn, vv, ll, rr)
        ctx.eps = eps
        ctx.rshape = roughness.shape
        return out

    @staticmethod
    def backward(ctx, g_out):
        L, hh, nn, vv, ll, rr = ctx.saved_tensors
        dev = L.device
        P, N = int(L.shape[0]), int(L.shape[1])
        need = ctx.needs_input_grad
        gL = torch.empty_like(L) if need[0] else None
        gh = torch.empty_like(hh) if need[1] else None
        gl = torch.empty_like(ll) if need[4] else None
        gr = torch.empty(P, dtype=torch.float32, device=dev) if need[5] else None
        if need[2] or need[3]:
            raise NotImplementedError("specular_reflectance: gradients w.r.t. n / v are not provided "
                                      "(the reference detaches the G-buffer, mat_nvdiffrast.py:126,179)")
        g = _f32c(g_out)
        with _on(dev):
            _lib.check(_lib.load().texir_specular_reflectance_bwd(_ptr(L), _ptr(hh), _ptr(nn), _ptr(vv), _ptr(ll),
                                                                  _ptr(rr), P, N, ctx.eps, _ptr(g), _ptr(gL), _ptr(gh),
                                                                  _ptr(gl), _ptr(gr), _stream_ptr(dev)),
                       "texir_specular_reflectance_bwd")
        return gL, gh, None, None, gl, (gr.reshape(ctx.rshape) if gr is not None else None), None


def specular_reflectance(lighting, h, n, v, l, roughness, albedo=None, eps: float = TINY_TINY_NUMBER):
    """lighting, h, l [P,N,3]; n, v [P,3]; roughness [P,1] -> [P,3] (sum over samples, NOT divided by N).
    ``albedo`` is accepted and ignored like in the reference (mat_nvdiffrast.py:260)."""
    _need_cuda("specular_reflectance", lighting)
    return _Specular.apply(lighting, h, n, v, l, roughness, float(eps))


class _Diffuse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lighting, l, n, albedo, cosine):
        dev = lighting.device
        P, N = int(lighting.shape[0]), int(lighting.shape[1])
        L, ll, nn, aa = _f32c(lighting), _f32c(l), _f32c(n), _f32c(albedo)
        out = torch.empty(P, 3, dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.check(_lib.load().texir_diffuse_reflectance(_ptr(L), _ptr(ll), _ptr(nn), _ptr(aa), P, N, cosine,
                                                             _ptr(out), _stream_ptr(dev)), "texir_diffuse_reflectance")
        ctx.save_for_backward(L, ll, nn, aa)
        ctx.cosine = cosine
        return out

    @staticmethod
    def backward(ctx, g_out):
        L, ll, nn, aa = ctx.saved_tensors
        dev = L.device
        P, N = int(L.shape[0]), int(L.shape[1])
        need = ctx.needs_input_grad
        if need[2]:
            raise NotImplementedError("diffuse_reflectance: gradient w.r.t. n is not provided")
        gL = torch.empty_like(L) if need[0] else None
        gl = torch.empty_like(ll) if need[1] else None
        ga = torch.empty_like(aa) if need[3] else None
        g = _f32c(g_out)
        with _on(dev):
            _lib.check(_lib.load().texir_diffuse_reflectance_bwd(_ptr(L), _ptr(ll), _ptr(nn), _ptr(aa), P, N, ctx.cosine,
                                                                 _ptr(g), _ptr(gL), _ptr(gl), _ptr(ga),
                                                                 _stream_ptr(dev)), "texir_diffuse_reflectance_bwd")
        return gL, gl, None, ga, None


def diffuse_reflectance(lighting, l, n, albedo, sample_type: str = "uniform"):
    """lighting, l [P,N,3]; n, albedo [P,3] -> [P,3] (sum over samples; the caller divides by N)"""
    _need_cuda("diffuse_reflectance", lighting)
    return _Diffuse.apply(lighting, l, n, albedo, 1 if sample_type == "cosine" else 0)


# --------------------------------------------------------------------------------------------------
# fused render
# --------------------------------------------------------------------------------------------------
def _check_recompute_texture(ctx, casts_rays: bool) -> None:
    """a backward that re-casts the rays must find the radiance texture of its forward pass: set_texture repacks in place
    (stage -1 swaps self.texture and swaps it back, mat_nvdiffrast.py:143-150), so a texture replaced in between would
    silently give the gradients of another scene"""
    if casts_rays and ctx.scene.texture_key() != ctx.tex_key:
        raise _lib.TexirError("backward='recompute' re-traces the rays against the scene's radiance texture, which has been "
                              "replaced since the forward pass; use backward='stash' or restore the texture first")



class _Shade(torch.autograd.Function):
    """rgb = albedo * dfac + (1/N) sum_i L_i w_i(roughness).  backward='stash': the forward kernel also
    emits d rgb / d roughness by forward-mode differentiation (roughness is the only non-trivial
    input), so backward launches no rays; backward='recompute': nothing is stashed and the backward
    kernel re-generates the samples and re-casts every ray (the north-star formulation)."""

    @staticmethod
    def forward(ctx, scene, normal, albedo, roughness, points, cam, irr, shift_spec, shift_diff, N_spec, N_diff,
                mode_spec, mode_diff, eps, backward_mode):
        dev = normal.device
        P = int(normal.shape[0])
        nrm, alb, pos = _f32c(normal), _f32c(albedo), _f32c(points)
        rough = _f32c(roughness).reshape(-1)
        camc = _f32c(cam, dev).reshape(3)
        irrc = _f32c(irr) if irr is not None else None
        shs = _f32c(shift_spec, dev).reshape(P, 2)
        shd = _f32c(shift_diff, dev).reshape(P, 2) if N_diff > 0 else None
        need_r = roughness.requires_grad
        need_a = albedo.requires_grad
        stash = backward_mode == "stash"
        rgb = torch.empty(P, 3, dtype=torch.float32, device=dev)
        drgb = torch.empty(P, 3, dtype=torch.float32, device=dev) if (stash and need_r) else None
        ediff = torch.empty(P, 3, dtype=torch.float32, device=dev) if (stash and N_diff > 0 and need_a) else None
        with _on(dev):
            _lib.check(_lib.load().texir_shade_fwd(scene.handle, _ptr(nrm), _ptr(alb), _ptr(rough), _ptr(pos), _ptr(camc),
                                                   _ptr(irrc), _ptr(shs), _ptr(shd), P, N_spec, N_diff, mode_spec,
                                                   mode_diff, eps, _ptr(rgb), _ptr(drgb), _ptr(ediff),
                                                   _stream_ptr(dev)), "texir_shade_fwd")
        ctx.scene = scene
        ctx.tex_key = scene.texture_key()
        ctx.cfg = (P, N_spec, N_diff, mode_spec, mode_diff, eps)
        ctx.shapes = (albedo.shape, roughness.shape, None if irr is None else irr.shape)
        ctx.have = (drgb is not None, ediff is not None, irrc is not None, shd is not None)
        saved = [nrm, alb, rough, pos, camc, shs]
        for t in (irrc, shd, drgb, ediff):
            if t is not None:
                saved.append(t)
        ctx.save_for_backward(*saved)
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        saved = list(ctx.saved_tensors)
        nrm, alb, rough, pos, camc, shs = saved[:6]
        rest = saved[6:]
        have_drgb, have_ediff, have_irr, have_shd = ctx.have
        irrc = rest.pop(0) if have_irr else None
        shd = rest.pop(0) if have_shd else None
        drgb = rest.pop(0) if have_drgb else None
        ediff = rest.pop(0) if have_ediff else None
        P, N_spec, N_diff, mode_spec, mode_diff, eps = ctx.cfg
        dev = nrm.device
        need = ctx.needs_input_grad  # (scene, normal, albedo, roughness, points, cam, irr, ...)
        _check_recompute_texture(ctx, (need[3] and drgb is None) or (need[2] and N_diff > 0 and ediff is None))
        g = _f32c(g_rgb).reshape(P, 3)
        g_alb = torch.empty(P, 3, dtype=torch.float32, device=dev) if need[2] else None
        g_rough = torch.empty(P, dtype=torch.float32, device=dev) if need[3] else None
        g_irr = torch.empty(P, 3, dtype=torch.float32, device=dev) if (need[6] and have_irr) else None
        with _on(dev):
            _lib.check(_lib.load().texir_shade_bwd(
                ctx.scene.handle, _ptr(nrm), _ptr(alb), _ptr(rough), _ptr(pos), _ptr(camc), _ptr(irrc), _ptr(shs),
                _ptr(shd), P, N_spec, N_diff, mode_spec, mode_diff, eps, _ptr(g),
                _ptr(drgb), _ptr(ediff),
                _ptr(g_alb), _ptr(g_rough), _ptr(g_irr), _stream_ptr(dev)), "texir_shade_bwd")
        a_shape, r_shape, i_shape = ctx.shapes
        return (None, None, g_alb.reshape(a_shape) if g_alb is not None else None,
                g_rough.reshape(r_shape) if g_rough is not None else None, None, None,
                g_irr.reshape(i_shape) if g_irr is not None else None, None, None, None, None, None, None, None, None)


def shade(scene: Scene, normal, albedo, roughness, points, cam_position, irr, shift_spec, *, num_spec: int,
          shift_diff=None, num_diff: int = 0, mode_spec: str = "importance", mode_diff: str = "uniform",
          eps: float = TINY_TINY_NUMBER, backward: str = "stash") -> torch.Tensor:
    """Fused MaterialModel.render on flat per-point tensors: normal/albedo/points/irr [P,3], roughness [P,1],
    cam_position [3], shift_spec [P,1,2] -> rgb [P,3].  num_diff > 0 selects the relighting variant
    (Monte-Carlo diffuse with its own shift, models/test_nvdiffrast.py:275-281)."""
    _need_cuda("shade", normal)
    if backward not in ("stash", "recompute"):
        raise ValueError("backward must be 'stash' or 'recompute'")
    return _Shade.apply(scene, normal, albedo, roughness, points, cam_position, irr, shift_spec, shift_diff,
                        int(num_spec), int(num_diff), _lib.MODES[mode_spec], _lib.MODES[mode_diff], float(eps), backward)


class _ShadeTextured(torch.autograd.Function):
    """_Shade whose backward kernel accumulates straight into TEXTURE space: the per-pixel albedo / roughness values
    come in detached, the gradients go out as gradients of the material textures (texir_shade_bwd_tex: the render
    backward and the atomic scatter of dr.texture's backward in one launch, then the pyramid folds)."""

    @staticmethod
    def forward(ctx, scene, materials_a, materials_r, normal, albedo_px, rough_px, points, cam, irr, shift_spec,
                shift_diff, texc, texd, levels_a, levels_r, N_spec, N_diff, mode_spec, mode_diff, eps, backward_mode):
        dev = normal.device
        P = int(normal.shape[0])
        nrm, alb, pos = _f32c(normal), _f32c(albedo_px), _f32c(points)
        rough = _f32c(rough_px).reshape(-1)
        camc = _f32c(cam, dev).reshape(3)
        irrc = _f32c(irr) if irr is not None else None
        shs = _f32c(shift_spec, dev).reshape(P, 2)
        shd = _f32c(shift_diff, dev).reshape(P, 2) if N_diff > 0 else None
        need_a = materials_a is not None and materials_a.requires_grad
        need_r = materials_r is not None and materials_r.requires_grad
        stash = backward_mode == "stash"
        rgb = torch.empty(P, 3, dtype=torch.float32, device=dev)
        drgb = torch.empty(P, 3, dtype=torch.float32, device=dev) if (stash and need_r) else None
        ediff = torch.empty(P, 3, dtype=torch.float32, device=dev) if (stash and N_diff > 0 and need_a) else None
        with _on(dev):
            _lib.check(_lib.load().texir_shade_fwd(scene.handle, _ptr(nrm), _ptr(alb), _ptr(rough), _ptr(pos), _ptr(camc),
                                                   _ptr(irrc), _ptr(shs), _ptr(shd), P, N_spec, N_diff, mode_spec,
                                                   mode_diff, eps, _ptr(rgb), _ptr(drgb), _ptr(ediff),
                                                   _stream_ptr(dev)), "texir_shade_fwd")
        ctx.scene = scene
        ctx.tex_key = scene.texture_key()
        ctx.cfg = (P, N_spec, N_diff, mode_spec, mode_diff, eps, levels_a, levels_r)
        ctx.shape_a = tuple(materials_a.shape) if materials_a is not None else None
        ctx.shape_r = tuple(materials_r.shape) if materials_r is not None else None
        ctx.have = (drgb is not None, ediff is not None, irrc is not None, shd is not None, texd is not None)
        saved = [nrm, alb, rough, pos, camc, shs, _f32c(texc).reshape(P, 2)]
        for t in (irrc, shd, drgb, ediff, _f32c(texd).reshape(P, 4) if texd is not None else None):
            if t is not None:
                saved.append(t)
        ctx.save_for_backward(*saved)
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        saved = list(ctx.saved_tensors)
        nrm, alb, rough, pos, camc, shs, uv = saved[:7]
        rest = saved[7:]
        have_drgb, have_ediff, have_irr, have_shd, have_da = ctx.have
        irrc = rest.pop(0) if have_irr else None
        shd = rest.pop(0) if have_shd else None
        drgb = rest.pop(0) if have_drgb else None
        ediff = rest.pop(0) if have_ediff else None
        uv_da = rest.pop(0) if have_da else None
        P, N_spec, N_diff, mode_spec, mode_diff, eps, La, Lr = ctx.cfg
        dev = nrm.device
        need = ctx.needs_input_grad  # (scene, materials_a, materials_r, ...)
        _check_recompute_texture(ctx, (need[2] and drgb is None) or (need[1] and N_diff > 0 and ediff is None))
        lib = _lib.load()
        g = _f32c(g_rgb).reshape(P, 3)

        def target(shape, L, wanted):
            if not wanted:
                return None, None, 0, 0
            H, W, C = shape
            g_tex = torch.zeros(H, W, C, dtype=torch.float32, device=dev)
            g_mips = torch.zeros(max(int(lib.texir_texture_mip_elems(H, W, C, L)), 1), dtype=torch.float32, device=dev)
            return g_tex, g_mips, H, W

        ga, gma, Ha, Wa = target(ctx.shape_a, La, need[1])
        gr, gmr, Hr, Wr = target(ctx.shape_r, Lr, need[2])
        if ga is None and gr is None:
            return (None,) * 21
        with _on(dev):
            st = _stream_ptr(dev)
            _lib.check(lib.texir_shade_bwd_tex(
                ctx.scene.handle, _ptr(nrm), _ptr(alb), _ptr(rough), _ptr(pos), _ptr(camc), _ptr(irrc), _ptr(shs), _ptr(shd),
                P, N_spec, N_diff, mode_spec, mode_diff, eps, _ptr(g), _ptr(drgb), _ptr(ediff), _ptr(uv), _ptr(uv_da),
                _ptr(ga), _ptr(gma), Ha, Wa, La, _ptr(gr), _ptr(gmr), Hr, Wr, Lr, st), "texir_shade_bwd_tex")
            if ga is not None and La > 0:
                _lib.check(lib.texir_texture_mip_fold(_ptr(ga), _ptr(gma), Ha, Wa, 3, La, st), "texir_texture_mip_fold")
            if gr is not None and Lr > 0:
                _lib.check(lib.texir_texture_mip_fold(_ptr(gr), _ptr(gmr), Hr, Wr, 1, Lr, st), "texir_texture_mip_fold")
        return (None, ga, gr) + (None,) * 18


def shade_textured(scene: Scene, materials_a: Optional[torch.Tensor], materials_r: Optional[torch.Tensor], normal,
                   albedo_px, rough_px, points, cam_position, irr, shift_spec, texc, texd=None, *, num_spec: int,
                   albedo_mipmap: bool = True, rough_mipmap: bool = True, max_mip_level: Optional[int] = None,
                   shift_diff=None, num_diff: int = 0, mode_spec: str = "importance", mode_diff: str = "uniform",
                   eps: float = TINY_TINY_NUMBER, backward: str = "stash") -> torch.Tensor:
    """Fused render whose gradients land in the material TEXTURES: albedo_px [P,3] / rough_px [P,1] are the values
    fetched at texc [P,2] (texd [P,4] for the mip-mapped filter) from materials_a [Ha,Wa,3] / materials_r [Hr,Wr,1]
    (pass None for a texture that must not receive a gradient, e.g. materials_a in the reference's stage 1); the
    backward is ONE kernel that differentiates the render and scatters into both textures' gradients with atomics."""
    _need_cuda("shade_textured", normal)
    if backward not in ("stash", "recompute"):
        raise ValueError("backward must be 'stash' or 'recompute'")
    mip = texd is not None
    La = mip_levels(int(materials_a.shape[0]), int(materials_a.shape[1]), max_mip_level) if (materials_a is not None and mip and albedo_mipmap) else 0
    Lr = mip_levels(int(materials_r.shape[0]), int(materials_r.shape[1]), max_mip_level) if (materials_r is not None and mip and rough_mipmap) else 0
    da = texd if (La > 0 or Lr > 0) else None  # a texture with 0 levels is scattered bilinearly whatever uv_da says
    return _ShadeTextured.apply(scene, materials_a, materials_r, normal, albedo_px.detach(), rough_px.detach(), points,
                                cam_position, irr, shift_spec, shift_diff, texc, da, La, Lr, int(num_spec), int(num_diff),
                                _lib.MODES[mode_spec], _lib.MODES[mode_diff], float(eps), backward)


# --------------------------------------------------------------------------------------------------
# the whole Mat-stage step as ONE autograd node (fetch -> render -> texture-space backward)
# --------------------------------------------------------------------------------------------------
# measured on B200 (profiles/r02_gbuffer_ab.txt): 0.859 -> 0.802 ms per 16-spp step once the host no longer waits for
# the shift kernel (sample_util.deferred_rng_commit); while it did, the extra stream / event calls cost more than the
# overlap gave back.  TEXIR_MAT_PREFETCH=0 builds the pyramids on the caller's stream.
_PREFETCH = os.environ.get("TEXIR_MAT_PREFETCH", "1") != "0"


class MatWorkspace:
    """Buffers a Mat-stage step re-uses from one optimiser step to the next (one per model and device):

      * the mip pyramids of materials_a / materials_r / irrt: storage re-used from step to step; the trainable
        textures' pyramids are rebuilt every step (two launches each), the frozen IrT's only when the tensor changed;
      * the texture-space scatter targets of the backward kernel (level 0 + pyramid per material texture).  They are
        zeroed once; the fold kernel that consumes them leaves them zeroed again, so no step pays for a memset of the
        150 MiB they occupy at the reference's texture sizes (2048^2 x 3 and 4096^2 x 1, mat_nvdiffrast.py:68-69).
    """

    def __init__(self):
        self._pyr = {}       # name -> (key, mips, L)
        self._scatter = {}   # name -> (shape key, level-0 buffer, pyramid buffer)
        self.dirty = False   # a backward died between the scatter and the fold: re-zero before the next use
        self._ahead = {}     # name -> (key, event): pyramids built ahead of the step on the side stream (prefetch)
        self._side = {}      # device -> side stream

    def pyramid(self, name: str, tex: torch.Tensor, max_mip_level: Optional[int]):
        key = (tex.data_ptr(), tex._version, tuple(tex.shape), max_mip_level)
        hit = self._pyr.get(name)
        ahead = self._ahead.pop(name, None)
        if ahead is not None:
            torch.cuda.current_stream(tex.device).wait_event(ahead[1])   # built ahead on the side stream (prefetch)
            if hit is not None and ahead[0] == key and hit[0] == key:
                return hit[1], hit[2]
        # a trainable texture is rebuilt on every step (its optimiser updates it in between, possibly through .data,
        # which no version counter sees); a frozen one (the IrT, requires_grad=False) only when it was replaced
        if hit is not None and hit[0] == key and not tex.requires_grad:
            return hit[1], hit[2]
        H, W, C = (int(x) for x in tex.shape)
        L = mip_levels(H, W, max_mip_level)
        lib = _lib.load()
        n = mip_elems(H, W, C, L)
        mips = hit[1] if (hit is not None and hit[1].numel() == n and hit[1].device == tex.device) else \
            torch.empty(n, dtype=torch.float32, device=tex.device)
        t = _f32c(tex)
        with _on(tex.device):
            _lib.check(lib.texir_texture_mip_build_fast(_ptr(t), H, W, C, L, _ptr(mips), _stream_ptr(tex.device)),
                       "texir_texture_mip_build_fast")
        self._pyr[name] = (key, mips, L)
        return mips, L

    def prefetch(self, textures: dict, max_mip_level: Optional[int]) -> None:
        """build the pyramids of `textures` (name -> tensor) NOW, on a side stream, so that they run beside whatever the
        caller launches next on the current stream -- the G-buffer walk of the same step, which leaves most warp slots of
        the GPU empty.  The pyramid() call of the step then only waits for the event.  TEXIR_MAT_PREFETCH=0 turns it off."""
        if not textures or not _PREFETCH:
            return
        dev = next(iter(textures.values())).device
        main = torch.cuda.current_stream(dev)
        side = self._side.get(dev)
        if side is None:
            side = self._side[dev] = torch.cuda.Stream(device=dev)
        side.wait_stream(main)          # the optimiser's update of the textures is on the current stream
        with torch.cuda.stream(side):
            for name, tex in textures.items():
                self._ahead.pop(name, None)
                self.pyramid(name, tex, max_mip_level)
        ev = torch.cuda.Event()
        ev.record(side)
        for name, tex in textures.items():
            self._ahead[name] = ((tex.data_ptr(), tex._version, tuple(tex.shape), max_mip_level), ev)

    def scatter_target(self, name: str, shape, L: int, dev):
        H, W, C = shape
        key = (H, W, C, L, str(dev))
        hit = self._scatter.get(name)
        if hit is None or hit[0] != key or self.dirty:
            n = mip_elems(H, W, C, L)
            hit = (key, torch.zeros(H, W, C, dtype=torch.float32, device=dev), torch.zeros(n, dtype=torch.float32, device=dev))
            self._scatter[name] = hit
        return hit[1], hit[2]


class _MatStage(torch.autograd.Function):
    """(materials_a, materials_r) -> (rgb, albedo, roughness, roughness_womipmap, irr) per pixel: the four dr.texture
    fetches of models/mat_nvdiffrast.py:131-139, the fused render of :179-182, and ONE backward kernel that takes the
    gradients of all outputs and accumulates into the two textures' gradients (texir_shade_bwd_mat), then one fold
    launch per texture."""

    @staticmethod
    def forward(ctx, materials_a, materials_r, ws, scene, irrt, normal, points, cam, texc, texd, shift_spec, stage,
                N_spec, mode_spec, eps, backward_mode, max_mip_level, gather_group):
        dev = normal.device
        P = int(normal.shape[0])
        lib = _lib.load()
        nrm, pos = _f32c(normal), _f32c(points)
        camc = _f32c(cam, dev).reshape(3)
        uv, da = _f32c(texc).reshape(P, 2), _f32c(texd).reshape(P, 4)
        shs = _f32c(shift_spec, dev).reshape(P, 2)
        ta, tr, ti = _f32c(materials_a), _f32c(materials_r), _f32c(irrt)
        ma, La = ws.pyramid("a", materials_a, max_mip_level)
        mr, Lr = ws.pyramid("r", materials_r, max_mip_level)
        mi, Li = ws.pyramid("i", irrt, max_mip_level)
        f32 = dict(dtype=torch.float32, device=dev)
        albedo, irr = torch.empty(P, 3, **f32), torch.empty(P, 3, **f32)
        rough, rough_w = torch.empty(P, **f32), torch.empty(P, **f32)
        (Ha, Wa, _), (Hr, Wr, _), (Hi, Wi, _) = ta.shape, tr.shape, ti.shape
        need_a = bool(materials_a.requires_grad) and stage == 2
        need_r = bool(materials_r.requires_grad)
        stash = backward_mode == "stash"
        rgb = torch.empty(P, 3, **f32)
        drgb = torch.empty(P, 3, **f32) if (stash and need_r) else None
        rough_in = rough_w if stage == 1 else rough
        with _on(dev):
            st = _stream_ptr(dev)
            _lib.check(lib.texir_mat_fetch(_ptr(ta), _ptr(ma), Ha, Wa, La, _ptr(tr), _ptr(mr), Hr, Wr, Lr, _ptr(ti), _ptr(mi),
                                           Hi, Wi, Li, _ptr(uv), _ptr(da), P, _ptr(albedo), _ptr(rough), _ptr(rough_w),
                                           _ptr(irr), st), "texir_mat_fetch")
            _lib.check(lib.texir_shade_fwd(scene.handle, _ptr(nrm), _ptr(albedo), _ptr(rough_in), _ptr(pos), _ptr(camc),
                                           _ptr(irr), _ptr(shs), None, P, N_spec, 0, mode_spec, _lib.MODE_UNIFORM, eps,
                                           _ptr(rgb), _ptr(drgb), None, st), "texir_shade_fwd")
        ctx.ws, ctx.scene = ws, scene
        ctx.gather_group = gather_group
        ctx.tex_key = scene.texture_key()
        ctx.cfg = (P, N_spec, mode_spec, eps, stage, La, Lr, (int(Ha), int(Wa)), (int(Hr), int(Wr)), need_a)
        ctx.have_drgb = drgb is not None
        saved = [nrm, albedo, rough_in, pos, camc, irr, shs, uv, da] + ([drgb] if drgb is not None else [])
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(irr)
        ctx.set_materialize_grads(False)   # an output the loss does not touch arrives as None, not as a tensor of zeros
        return rgb, albedo, rough.unsqueeze(-1), rough_w.unsqueeze(-1), irr

    @staticmethod
    def backward(ctx, g_rgb, g_albedo, g_rough, g_rough_w, _g_irr):
        saved = list(ctx.saved_tensors)
        nrm, albedo, rough_in, pos, camc, irr, shs, uv, da = saved[:9]
        drgb = saved[9] if ctx.have_drgb else None
        P, N_spec, mode_spec, eps, stage, La, Lr, (Ha, Wa), (Hr, Wr), render_a = ctx.cfg
        dev = nrm.device
        need_a, need_r = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gather = ctx.gather_group is not None
        need_a = need_a and (render_a or g_albedo is not None or gather)   # stage 1: the render saw albedo.detach()
        if not need_a and not need_r:
            return (None,) * 18
        _check_recompute_texture(ctx, need_r and drgb is None)
        lib, ws = _lib.load(), ctx.ws
        g = _f32c(g_rgb).reshape(P, 3) if g_rgb is not None else torch.zeros(P, 3, dtype=torch.float32, device=dev)
        add_a = _f32c(g_albedo).reshape(P, 3) if (g_albedo is not None and need_a) else None
        add_r = _f32c(g_rough).reshape(P) if (g_rough is not None and need_r) else None
        add_rw = _f32c(g_rough_w).reshape(P) if (g_rough_w is not None and need_r) else None
        sa, sma = ws.scatter_target("a", (Ha, Wa, 3), La, dev) if need_a else (None, None)
        sr, smr = ws.scatter_target("r", (Hr, Wr, 1), Lr, dev) if need_r else (None, None)
        n_a, n_r = (Ha * Wa * 3 if need_a else 0), (Hr * Wr if need_r else 0)
        flat = torch.empty(n_a + n_r, dtype=torch.float32, device=dev)  # ONE buffer: a single all-reduce covers both
        out_a = flat[:n_a].view(Ha, Wa, 3) if need_a else None
        out_r = flat[n_a:].view(Hr, Wr, 1) if need_r else None
        ws.dirty = True
        if gather:
            # "gather" synchronisation: per-pixel gradients of THIS rank's pixels, then one all-gather of 11 floats per
            # pixel over all ranks and the same scatter on every rank -- instead of an all-reduce of the textures
            import torch.distributed as dist
            g_alb_px = torch.zeros(P, 3, dtype=torch.float32, device=dev)
            g_rough_px = torch.zeros(P, dtype=torch.float32, device=dev)
            with _on(dev):
                _lib.check(lib.texir_shade_bwd(ctx.scene.handle, _ptr(nrm), _ptr(albedo), _ptr(rough_in), _ptr(pos), _ptr(camc),
                                               _ptr(irr), _ptr(shs), None, P, N_spec, 0, mode_spec, _lib.MODE_UNIFORM, eps, _ptr(g),
                                               _ptr(drgb), None, _ptr(g_alb_px) if (render_a and need_a) else None,
                                               _ptr(g_rough_px) if need_r else None, None, _stream_ptr(dev)), "texir_shade_bwd")
            if add_a is not None:
                g_alb_px = g_alb_px + add_a
            zero = torch.zeros(P, dtype=torch.float32, device=dev)
            r_mip = (g_rough_px if stage == 2 else zero) + (add_r if add_r is not None else 0)
            r_bil = (g_rough_px if stage == 1 else zero) + (add_rw if add_rw is not None else 0)
            rows = torch.cat([uv, da, g_alb_px, r_mip.unsqueeze(1), r_bil.unsqueeze(1)], 1).contiguous()
            from .parallel import all_gather_rows
            world = dist.get_world_size(ctx.gather_group[1])
            allrows = all_gather_rows(rows, ctx.gather_group[1])
            with _on(dev):
                st = _stream_ptr(dev)
                _lib.check(lib.texir_mat_scatter_rows(_ptr(allrows), world * P, _ptr(sa), _ptr(sma), Ha, Wa, La if need_a else 0,
                                                      _ptr(sr), _ptr(smr), Hr, Wr, Lr if need_r else 0, st), "texir_mat_scatter_rows")
                if need_a:
                    _lib.check(lib.texir_texture_mip_fold_into(_ptr(sa), _ptr(sma), Ha, Wa, 3, La, _ptr(out_a), 1, st),
                               "texir_texture_mip_fold_into")
                if need_r:
                    _lib.check(lib.texir_texture_mip_fold_into(_ptr(sr), _ptr(smr), Hr, Wr, 1, Lr, _ptr(out_r), 1, st),
                               "texir_texture_mip_fold_into")
            ws.dirty = False
            return (out_a, out_r) + (None,) * 16
        with _on(dev):
            st = _stream_ptr(dev)
            _lib.check(lib.texir_shade_bwd_mat(
                ctx.scene.handle, _ptr(nrm), _ptr(albedo), _ptr(rough_in), _ptr(pos), _ptr(camc), _ptr(irr), _ptr(shs), None,
                P, N_spec, 0, mode_spec, _lib.MODE_UNIFORM, eps, _ptr(g), _ptr(drgb), None, _ptr(uv), _ptr(da),
                _ptr(add_a), _ptr(add_r), _ptr(add_rw), int(bool(render_a and need_a)), int(bool(need_r)),
                int(stage == 1), _ptr(sa), _ptr(sma), Ha, Wa, La if need_a else 0, _ptr(sr), _ptr(smr), Hr, Wr,
                Lr if need_r else 0, st), "texir_shade_bwd_mat")
            if need_a:
                _lib.check(lib.texir_texture_mip_fold_into(_ptr(sa), _ptr(sma), Ha, Wa, 3, La, _ptr(out_a), 1, st),
                           "texir_texture_mip_fold_into")
            if need_r:
                _lib.check(lib.texir_texture_mip_fold_into(_ptr(sr), _ptr(smr), Hr, Wr, 1, Lr, _ptr(out_r), 1, st),
                           "texir_texture_mip_fold_into")
        ws.dirty = False
        return (out_a, out_r) + (None,) * 16


def mat_stage(scene: Scene, ws: MatWorkspace, materials_a, materials_r, irrt, normal, points, cam_position, texc, texd,
              shift_spec, *, stage: int, num_spec: int, max_mip_level: Optional[int] = None,
              mode_spec: str = "importance", eps: float = TINY_TINY_NUMBER, backward: str = "stash", gather_group=None):
    """One Mat-stage step on flat per-pixel G-buffer tensors (normal, points [P,3]; texc [P,2]; texd [P,4]):
    -> (rgb [P,3], albedo [P,3], roughness [P,1], roughness_womipmap [P,1], irr [P,3]) exactly as
    MaterialModel.forward computes them for stage 1 (render(albedo.detach(), roughness_womipmap)) or stage 2
    (render(albedo, roughness)), models/mat_nvdiffrast.py:131-139,179-182.  Gradients of every output reach
    materials_a / materials_r through ONE backward kernel + one fold per texture.
    gather_group = (True, process group or None): multi-GPU "gather" synchronisation -- the pixels are this rank's share,
    the backward all-gathers the per-pixel gradients of all ranks (11 floats per pixel) and scatters all of them, so every
    rank ends up with the FULL texture gradients and no all-reduce of the textures is needed."""
    _need_cuda("mat_stage", normal)
    if stage not in (1, 2):
        raise ValueError("mat_stage covers the ray-traced stages 1 and 2")
    if backward not in ("stash", "recompute"):
        raise ValueError("backward must be 'stash' or 'recompute'")
    return _MatStage.apply(materials_a, materials_r, ws, scene, irrt, normal, points, cam_position, texc, texd, shift_spec,
                           int(stage), int(num_spec), _lib.MODES[mode_spec], float(eps), backward, max_mip_level, gather_group)


# --------------------------------------------------------------------------------------------------
# fused IrT batch
# --------------------------------------------------------------------------------------------------
def point_order(scene: Scene, position: torch.Tensor) -> torch.Tensor:
    """int32 [P] permutation that sorts the points along a Morton curve over the scene box"""
    _need_cuda("point_order", position)
    dev = position.device
    P = int(position.shape[0])
    pos = _f32c(position)
    order = torch.empty(P, dtype=torch.int32, device=dev)
    with _on(dev):
        _lib.check(_lib.load().texir_point_order(scene.handle, _ptr(pos), P, _ptr(order), _stream_ptr(dev)),
                   "texir_point_order")
    return order


def irradiance(scene: Scene, position: torch.Tensor, normal: torch.Tensor, shift: torch.Tensor, num_samples: int,
               mode: str = "uniform", out: Optional[torch.Tensor] = None, coherent: bool = True,
               order: Optional[torch.Tensor] = None) -> torch.Tensor:
    """E[p] = sum_i L_i clamp(n.l_i,0,1) 2 pi / N for P points in one launch (tracer_o3d_irt.py:168-171);
    position, normal [P,3], shift [P,1,2] or [P,2] -> [P,3].  coherent=True lets warps take the points in Morton
    order (texir_point_order); the values do not depend on it.  order (int32, device): compute only the listed
    points (rows of `out` that are not listed are left untouched)."""
    _need_cuda("irradiance", position)
    dev = position.device
    P = int(position.shape[0])
    pos, nrm = _f32c(position), _f32c(normal)
    sh = _f32c(shift, dev).reshape(P, 2)
    E = out if out is not None else torch.empty(P, 3, dtype=torch.float32, device=dev)
    if order is None:
        order = point_order(scene, pos) if (coherent and P > 4096) else None
        count = P
    else:
        order = order.to(device=dev, dtype=torch.int32).contiguous()
        count = int(order.numel())
    with _on(dev):
        _lib.check(_lib.load().texir_irradiance_fwd(scene.handle, _ptr(pos), _ptr(nrm), _ptr(sh), _ptr(order), count,
                                                    int(num_samples), _lib.MODES[mode], _ptr(E), _stream_ptr(dev)),
                   "texir_irradiance_fwd")
    return E


def irradiance_peers(scene: Scene, position: torch.Tensor, normal: torch.Tensor, shift: torch.Tensor,
                     order: torch.Tensor, num_samples: int, mode: str, peer_ptrs, num_peers: int) -> None:
    """this rank's share (`order`: int32 point indices) of a multi-GPU IrT; every result row is stored into ALL the
    [P,3] buffers of peer_ptrs (ctypes array of device pointers, parallel.PeerBuffer.ptr_array) by the kernel itself"""
    _need_cuda("irradiance_peers", position)
    dev = position.device
    P = int(position.shape[0])
    pos, nrm = _f32c(position), _f32c(normal)
    sh = _f32c(shift, dev).reshape(P, 2)
    order = order.to(device=dev, dtype=torch.int32).contiguous()
    with _on(dev):
        _lib.check(_lib.load().texir_irradiance_fwd_peers(scene.handle, _ptr(pos), _ptr(nrm), _ptr(sh), _ptr(order),
                                                          int(order.numel()), int(num_samples), _lib.MODES[mode], peer_ptrs,
                                                          int(num_peers), _stream_ptr(dev)), "texir_irradiance_fwd_peers")


def radiance_samples(scene: Scene, position: torch.Tensor, normal: torch.Tensor, shift: torch.Tensor,
                     num_samples: int, mode: str = "uniform", return_dirs: bool = False):
    """Target generator of the NIrF / irradiance-field stages (models/tracer_o3d.py:73-116): the N
    shifted-Hammersley directions of every point are traced and the hit radiance returned un-reduced.
    position, normal [P,3]; shift [P,1,2]; -> rad [P,N,3] (and the directions [P,N,3])."""
    _need_cuda("radiance_samples", position)
    dev = position.device
    P, N = int(position.shape[0]), int(num_samples)
    pos, nrm = _f32c(position), _f32c(normal)
    sh = _f32c(shift, dev).reshape(P, 2)
    rad = torch.empty(P, N, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(P, N, 3, dtype=torch.float32, device=dev) if return_dirs else None
    with _on(dev):
        _lib.check(_lib.load().texir_radiance_samples(scene.handle, _ptr(pos), _ptr(nrm), _ptr(sh), P, N,
                                                      _lib.MODES[mode], _ptr(rad), _ptr(dirs), _stream_ptr(dev)),
                   "texir_radiance_samples")
    return (rad, dirs) if return_dirs else rad


def gbuffer(scene: Scene, rays: torch.Tensor, offset: float = 0.0,
            ray_differentials: Optional[torch.Tensor] = None) -> dict:
    """G-buffer from primary rays [...,6] (replaces dr.rasterize + dr.interpolate, mat_nvdiffrast.py:119-128):
    dict(position, normal [...,3], texc [...,2] when the scene has UVs, primitive_ids, primitive_uvs, mask).
    position = hit + offset * normal (the reference shifts by 1e-2, :179); uncovered pixels get position =
    normal = (1,0,0).  With ray_differentials [...,6] = (d dir/dX, d dir/dY) per pixel also texd [...,4] =
    (du/dX, du/dY, dv/dX, dv/dY), the uv_da that the mip-mapped texture_fetch takes."""
    _need_cuda("gbuffer", rays)
    dev = rays.device
    r = _f32c(rays).reshape(-1, 6)
    shape = tuple(rays.shape[:-1])
    R = int(r.shape[0])
    pos = torch.empty(R, 3, dtype=torch.float32, device=dev)
    nrm = torch.empty(R, 3, dtype=torch.float32, device=dev)
    uv = torch.empty(R, 2, dtype=torch.float32, device=dev) if scene.has_uvs else None
    prim = torch.empty(R, dtype=torch.int32, device=dev)
    bary = torch.empty(R, 2, dtype=torch.float32, device=dev)
    mask = torch.empty(R, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.check(_lib.load().texir_gbuffer(scene.handle, _ptr(r), R, _ptr(pos), _ptr(nrm), _ptr(uv), _ptr(prim),
                                             _ptr(bary), _ptr(mask), _stream_ptr(dev)), "texir_gbuffer")
    if offset:
        pos = torch.where(mask.unsqueeze(-1) > 0, pos + offset * nrm, pos)
    out = {"position": pos.reshape(shape + (3,)), "normal": nrm.reshape(shape + (3,)),
           "primitive_ids": prim.reshape(shape), "primitive_uvs": bary.reshape(shape + (2,)),
           "mask": mask.reshape(shape + (1,))}
    if uv is not None:
        out["texc"] = uv.reshape(shape + (2,))
    if ray_differentials is not None:
        dd = _f32c(ray_differentials, dev).reshape(-1, 6)
        if dd.shape[0] != R:
            raise ValueError("ray_differentials must have one (d dir/dX, d dir/dY) row per ray")
        texd = torch.empty(R, 4, dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.check(_lib.load().texir_gbuffer_uv_da(scene.handle, _ptr(r), _ptr(dd), _ptr(prim), R, _ptr(texd),
                                                       _stream_ptr(dev)), "texir_gbuffer_uv_da")
        out["texd"] = texd.reshape(shape + (4,))
    return out


def gbuffer_mvp(scene: Scene, mvp: torch.Tensor, cam_position: torch.Tensor, res: int, return_rays: bool = False,
                shard=None, flip_v: bool = False) -> dict:
    """The G-buffer of the views the reference rasterises, straight from its camera matrices and in one launch
    (mat_nvdiffrast.py:116-128): mvp [V,4,4] (clip = [v,1] @ mvp[i]), cam_position [3] -> dict(position, normal
    [V,res,res,3], texc [..,2], texd [..,4], primitive_ids, primitive_uvs, mask [..,1]) (+ rays [..,6]).
    shard = (rank, world): only this rank's pixels (128-pixel blocks rank, rank + world, ... of every view, see
    parallel.local_pixels), flat [V*res*res/world, .].  flip_v: texc = (u, 1 - v) and dv/dX, dv/dY negated, the
    convention of the reference's self.uvs (pyredner.load_obj flips v by default, mat_nvdiffrast.py:194)."""
    _need_cuda("gbuffer_mvp", mvp)
    dev = mvp.device
    m = _f32c(mvp).reshape(-1, 4, 4)
    cam = _f32c(cam_position, dev).reshape(3)
    V, res = int(m.shape[0]), int(res)
    rank, world = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
    if world > 1 and (res * res) % (128 * world) != 0:
        raise ValueError("a sharded G-buffer needs res*res to be a multiple of 128 * world")
    R = V * res * res // world
    f32 = dict(dtype=torch.float32, device=dev)
    pos, nrm, bary = torch.empty(R, 3, **f32), torch.empty(R, 3, **f32), torch.empty(R, 2, **f32)
    uv = torch.empty(R, 2, **f32) if scene.has_uvs else None
    texd = torch.empty(R, 4, **f32) if scene.has_uvs else None
    prim = torch.empty(R, dtype=torch.int32, device=dev)
    mask = torch.empty(R, **f32)
    rays = torch.empty(R, 6, **f32) if return_rays else None
    with _on(dev):
        _lib.check(_lib.load().texir_gbuffer_mvp_shard(scene.handle, _ptr(m), _ptr(cam), V, res, rank, world, int(flip_v),
                                                       _ptr(pos), _ptr(nrm), _ptr(uv), _ptr(texd), _ptr(prim), _ptr(bary),
                                                       _ptr(mask), _ptr(rays), _stream_ptr(dev)), "texir_gbuffer_mvp_shard")
    shape = (V, res, res) if world == 1 else (R,)
    out = {"position": pos.reshape(shape + (3,)), "normal": nrm.reshape(shape + (3,)),
           "primitive_ids": prim.reshape(shape), "primitive_uvs": bary.reshape(shape + (2,)),
           "mask": mask.reshape(shape + (1,))}
    if uv is not None:
        out["texc"] = uv.reshape(shape + (2,))
        out["texd"] = texd.reshape(shape + (4,))
    if rays is not None:
        out["rays"] = rays.reshape(shape + (6,))
    return out


# --------------------------------------------------------------------------------------------------
# material-texture fetch / gradient scatter (dr.texture, models/mat_nvdiffrast.py:131-139)
# --------------------------------------------------------------------------------------------------
class _TextureFetch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv):
        dev = tex.device
        H, W, C = (int(x) for x in tex.shape)
        P = int(uv.shape[0])
        t, u = _f32c(tex), _f32c(uv)
        out = torch.empty(P, C, dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.check(_lib.load().texir_texture_fetch(_ptr(t), H, W, C, _ptr(u), P, _ptr(out), _stream_ptr(dev)),
                       "texir_texture_fetch")
        ctx.save_for_backward(u)
        ctx.shape = (H, W, C)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (u,) = ctx.saved_tensors
        dev = u.device
        H, W, C = ctx.shape
        g = _f32c(g_out)
        g_tex = torch.zeros(H, W, C, dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.check(_lib.load().texir_texture_scatter(_ptr(g), H, W, C, _ptr(u), int(u.shape[0]), _ptr(g_tex),
                                                         _stream_ptr(dev)), "texir_texture_scatter")
        return g_tex, None


_mip_levels_cache = {}


def mip_levels(H: int, W: int, max_mip_level: Optional[int] = None) -> int:
    """mip levels above level 0 that dr.texture builds for a [H,W] texture under a max_mip_level cap"""
    key = (int(H), int(W), max_mip_level)
    L = _mip_levels_cache.get(key)
    if L is None:
        L = int(_lib.load().texir_texture_mip_levels(int(H), int(W), -1 if max_mip_level is None else int(max_mip_level)))
        _mip_levels_cache[key] = L
    if L < 0:
        raise ValueError("a %dx%d texture has an odd side that would have to be halved; lower max_mip_level" % (H, W))
    return L


_mip_elems_cache = {}


def mip_elems(H: int, W: int, C: int, L: int) -> int:
    """floats in the packed pyramid (levels 1..L) of a [H,W,C] texture"""
    key = (int(H), int(W), int(C), int(L))
    n = _mip_elems_cache.get(key)
    if n is None:
        n = _mip_elems_cache[key] = max(int(_lib.load().texir_texture_mip_elems(*key)), 1)
    return n


def build_mips(tex: torch.Tensor, max_mip_level: Optional[int] = None):
    """-> (packed levels 1..L as one float32 tensor, L); level l is [max(H>>l,1), max(W>>l,1), C]"""
    _need_cuda("build_mips", tex)
    H, W, C = (int(x) for x in tex.shape)
    L = mip_levels(H, W, max_mip_level)
    lib = _lib.load()
    t = _f32c(tex)
    mips = torch.empty(max(int(lib.texir_texture_mip_elems(H, W, C, L)), 1), dtype=torch.float32, device=tex.device)
    with _on(tex.device):
        _lib.check(lib.texir_texture_mip_build(_ptr(t), H, W, C, L, _ptr(mips), _stream_ptr(tex.device)),
                   "texir_texture_mip_build")
    return mips, L


class _TextureFetchMip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv, uv_da, max_mip_level):
        dev = tex.device
        H, W, C = (int(x) for x in tex.shape)
        P = int(uv.shape[0])
        t, u, d = _f32c(tex), _f32c(uv), _f32c(uv_da)
        mips, L = build_mips(t, max_mip_level)
        out = torch.empty(P, C, dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.check(_lib.load().texir_texture_fetch_mip(_ptr(t), _ptr(mips), H, W, C, L, _ptr(u), _ptr(d), P,
                                                           _ptr(out), None, _stream_ptr(dev)),
                       "texir_texture_fetch_mip")
        ctx.save_for_backward(u, d)
        ctx.shape = (H, W, C, L, mips.numel())
        return out

    @staticmethod
    def backward(ctx, g_out):
        u, d = ctx.saved_tensors
        dev = u.device
        H, W, C, L, n_mips = ctx.shape
        g = _f32c(g_out)
        g_tex = torch.zeros(H, W, C, dtype=torch.float32, device=dev)
        g_mips = torch.zeros(n_mips, dtype=torch.float32, device=dev)
        lib = _lib.load()
        with _on(dev):
            _lib.check(lib.texir_texture_scatter_mip(_ptr(g), H, W, C, L, _ptr(u), _ptr(d), int(u.shape[0]),
                                                     _ptr(g_tex), _ptr(g_mips), _stream_ptr(dev)),
                       "texir_texture_scatter_mip")
            _lib.check(lib.texir_texture_mip_fold(_ptr(g_tex), _ptr(g_mips), H, W, C, L, _stream_ptr(dev)),
                       "texir_texture_mip_fold")
        return g_tex, None, None, None


def texture_fetch(tex: torch.Tensor, uv: torch.Tensor, uv_da: Optional[torch.Tensor] = None,
                  filter_mode: str = "auto", max_mip_level: Optional[int] = None) -> torch.Tensor:
    """dr.texture(tex[None], uv, uv_da, filter_mode=..., max_mip_level=...) for a [H,W,C] texture, uv [P,2] and
    uv_da [P,4] = (du/dX, du/dY, dv/dX, dv/dY): 'linear' (bilinear) or 'linear-mipmap-linear' (trilinear over the
    box-filtered pyramid); 'auto' picks the latter when uv_da is given, as nvdiffrast does.  The backward scatters
    into the texture (and its pyramid, folded back to level 0) with atomics; uv is not differentiated (the
    reference's uvs are constants, mat_nvdiffrast.py:56)."""
    _need_cuda("texture_fetch", tex)
    if filter_mode == "auto":
        filter_mode = "linear" if uv_da is None else "linear-mipmap-linear"
    if filter_mode == "linear":
        return _TextureFetch.apply(tex, uv)
    if filter_mode != "linear-mipmap-linear":
        raise ValueError("filter_mode must be 'linear' or 'linear-mipmap-linear', got %r" % (filter_mode,))
    if uv_da is None:
        raise ValueError("'linear-mipmap-linear' needs uv_da (the reference always passes texd)")
    return _TextureFetchMip.apply(tex, uv, uv_da, max_mip_level)
