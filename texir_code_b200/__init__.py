"""texir_code_b200 -- B200 (sm_100a) implementation of TexIR's irradiance-texture / per-pixel BRDF hot path.

Public surface (mirrors the reference's names for this path, see DESIGN.md / INTEGRATION.md):

    sample_util.generate_dir (alias sample_hemisphere), RadicalInverse, Hammersley, generate_fixed_samples
    Scene, RaycastingScene                      (replaces o3d.t.geometry.RaycastingScene + triangle_uvs + texture)
    ops.query_irf / specular_reflectance / diffuse_reflectance / shade / irradiance / texture_fetch
    models.HotPathMixin / MaterialModel / TracerO3d
    postprocess.pad_texture / padding_texture / imwrite_hdr / imread_hdr / save_irr_texture   (IrT post-processing)
    parallel.shard_bounds / gather_points / allreduce_gradients / sharded_irradiance
    graphed.GraphedMatStep                      (one Mat-stage optimisation step captured as a CUDA graph)

Everything computes in libtexir_b200.so (hand-written CUDA behind the C ABI of include/texir_b200.h);
nothing here falls back to the CPU.
"""
from . import _lib
from ._lib import TexirError, launch_count


def __getattr__(name):  # lazy: importing the package must not require torch.cuda or the built library
    import importlib
    table = {
        "Scene": ("scene", "Scene"), "RaycastingScene": ("scene", "RaycastingScene"),
        "generate_dir": ("sample_util", "generate_dir"), "sample_hemisphere": ("sample_util", "generate_dir"),
        "query_irf": ("ops", "query_irf"), "trace": ("ops", "query_irf"),
        "specular_reflectance": ("ops", "specular_reflectance"),
        "diffuse_reflectance": ("ops", "diffuse_reflectance"),
        "shade": ("ops", "shade"), "irradiance": ("ops", "irradiance"), "texture_fetch": ("ops", "texture_fetch"),
        "radiance_samples": ("ops", "radiance_samples"), "gbuffer": ("ops", "gbuffer"),
        "point_order": ("ops", "point_order"), "gbuffer_mvp": ("ops", "gbuffer_mvp"),
        "HotPathMixin": ("models", "HotPathMixin"), "MaterialModel": ("models", "MaterialModel"),
        "TracerO3d": ("models", "TracerO3d"),
        "pad_texture": ("postprocess", "pad_texture"), "padding_texture": ("postprocess", "padding_texture"),
        "imwrite_hdr": ("postprocess", "imwrite_hdr"), "imread_hdr": ("postprocess", "imread_hdr"),
        "save_irr_texture": ("postprocess", "save_irr_texture"),
        "GraphedMatStep": ("graphed", "GraphedMatStep"),
    }
    if name in ("scene", "sample_util", "ops", "models", "parallel", "synth", "build", "postprocess", "graphed", "conf_models"):
        return importlib.import_module("." + name, __name__)
    if name in table:
        mod, attr = table[name]
        return getattr(importlib.import_module("." + mod, __name__), attr)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


__all__ = ["TexirError", "launch_count", "Scene", "RaycastingScene", "generate_dir", "sample_hemisphere", "query_irf",
           "trace", "specular_reflectance", "diffuse_reflectance", "shade", "irradiance", "texture_fetch", "radiance_samples", "gbuffer", "gbuffer_mvp", "point_order",
           "HotPathMixin", "MaterialModel", "TracerO3d", "pad_texture", "padding_texture", "imwrite_hdr", "imread_hdr",
           "save_irr_texture", "GraphedMatStep"]
