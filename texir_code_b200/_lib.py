"""ctypes binding of libtexir_b200.so (C ABI: include/texir_b200.h).

There is no fallback: if the shared library is missing or a call fails, the op raises.  The library is
built in-tree by texir_code_b200/build.py (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# TEXIR_LIB: tuning aid -- load another build of the same library (e.g. compiled with different launch bounds)
SO_PATH = os.environ.get("TEXIR_LIB") or os.path.join(HERE, "libtexir_b200.so")

MODE_UNIFORM, MODE_COSINE, MODE_IMPORTANCE = 0, 1, 2
MODES = {"uniform": MODE_UNIFORM, "cosine": MODE_COSINE, "importance": MODE_IMPORTANCE}
TEX_CHW, TEX_HWC = 0, 1
BVH2, BVH8 = 2, 8
BUILDER_AUTO, BUILDER_HOST, BUILDER_DEVICE = 0, 1, 2
BUILDERS = {"auto": 0, "host": 1, "device": 2}


class TexirError(RuntimeError):
    pass


class SceneInfo(Structure):
    _fields_ = [
        ("num_tris", c_int64), ("num_nodes2", c_int64), ("num_nodes8", c_int64),
        ("depth2", c_int32), ("depth8", c_int32), ("tex_h", c_int32), ("tex_w", c_int32),
        ("bvh_kind", c_int32), ("device", c_int32),
        ("sah2", c_double), ("sah8", c_double), ("build_seconds", c_double), ("device_bytes", c_int64),
        ("builder", c_int32), ("reserved", c_int32),
    ]


_P = c_void_p
# name -> (restype, argtypes); every symbol include/texir_b200.h declares
SIGNATURES = {
    "texir_last_error": (c_char_p, []),
    "texir_version": (c_int, []),
    "texir_launch_count": (c_int64, []),
    "texir_scene_create": (c_int, [_P, c_int64, _P, c_int64, _P, c_int, c_int, POINTER(c_void_p)]),
    "texir_scene_create_ex": (c_int, [_P, c_int64, _P, c_int64, _P, c_int, c_int, c_int, POINTER(c_void_p)]),
    "texir_scene_destroy": (c_int, [_P]),
    "texir_scene_get_info": (c_int, [_P, POINTER(SceneInfo)]),
    "texir_scene_set_bvh": (c_int, [_P, c_int]),
    "texir_scene_set_texture": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "texir_scene_set_texture_rgbe": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "texir_cast_rays": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P]),
    "texir_cast_rays_stats": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P]),
    "texir_generate_dir": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    "texir_generate_dir_samples": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    "texir_generate_dir_bwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, _P]),
    "texir_debug_sincos": (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    "texir_query_irf": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "texir_irradiance_fwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    "texir_irradiance_fwd_peers": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, c_int, _P]),
    "texir_point_order": (c_int, [_P, _P, c_int64, _P, _P]),
    "texir_radiance_samples": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "texir_gbuffer": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "texir_gbuffer_mvp": (c_int, [_P, _P, _P, c_int, c_int] + [_P] * 9),
    "texir_gbuffer_mvp_shard": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int] + [_P] * 9),
    "texir_gbuffer_uv_da": (c_int, [_P, _P, _P, _P, c_int64, _P, _P]),
    "texir_shade_fwd": (c_int, [_P] + [_P] * 8 + [c_int64, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "texir_shade_bwd": (c_int, [_P] + [_P] * 8 + [c_int64, c_int, c_int, c_int, c_int, c_float] + [_P] * 7),
    "texir_shade_bwd_tex": (c_int, [_P] + [_P] * 8 + [c_int64, c_int, c_int, c_int, c_int, c_float] + [_P] * 5
                            + [_P, _P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, _P]),
    "texir_shade_bwd_mat": (c_int, [_P] + [_P] * 8 + [c_int64, c_int, c_int, c_int, c_int, c_float] + [_P] * 8
                            + [c_int, c_int, c_int] + [_P, _P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, _P]),
    "texir_specular_reflectance": (c_int, [_P] * 6 + [c_int64, c_int, c_float, _P, _P]),
    "texir_specular_reflectance_bwd": (c_int, [_P] * 6 + [c_int64, c_int, c_float] + [_P] * 6),
    "texir_diffuse_reflectance": (c_int, [_P] * 4 + [c_int64, c_int, c_int, _P, _P]),
    "texir_diffuse_reflectance_bwd": (c_int, [_P] * 4 + [c_int64, c_int, c_int] + [_P] * 5),
    "texir_texture_fetch": (c_int, [_P, c_int, c_int, c_int, _P, c_int64, _P, _P]),
    "texir_texture_scatter": (c_int, [_P, c_int, c_int, c_int, _P, c_int64, _P, _P]),
    "texir_texture_mip_levels": (c_int, [c_int, c_int, c_int]),
    "texir_texture_mip_elems": (c_int64, [c_int, c_int, c_int, c_int]),
    "texir_texture_mip_build": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "texir_texture_fetch_mip": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P, _P, _P]),
    "texir_texture_scatter_mip": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P, _P, _P]),
    "texir_texture_mip_fold": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "texir_texture_mip_build_fast": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "texir_texture_mip_fold_into": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "texir_mat_fetch": (c_int, [_P, _P, c_int, c_int, c_int] * 3 + [_P, _P, c_int64, _P, _P, _P, _P, _P]),
    "texir_mat_scatter_rows": (c_int, [_P, c_int64, _P, _P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, _P]),
    "texir_mt19937_uniform": (c_int, [_P, c_int, c_int, c_int64, _P, _P, _P]),
    "texir_mt19937_advance_host": (c_int, [_P, _P, _P, c_int64]),
    "texir_peer_alloc": (c_int, [c_int, c_int64, POINTER(c_void_p), _P]),
    "texir_peer_free": (c_int, [c_int, _P]),
    "texir_peer_open": (c_int, [c_int, _P, POINTER(c_void_p)]),
    "texir_peer_close": (c_int, [c_int, _P]),
    "texir_peer_barrier": (c_int, [_P, c_int, c_int, ctypes.c_uint32, c_int, _P]),
    "texir_pad_texture": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "texir_rgbe_encode": (c_int, [_P, c_int64, c_int, _P, _P]),
    "texir_rgbe_decode": (c_int, [_P, c_int64, c_int, _P, _P]),
    "texir_hdr_write_host": (c_int, [c_char_p, _P, c_int, c_int]),
    "texir_hdr_read_info": (c_int, [c_char_p, POINTER(c_int), POINTER(c_int)]),
    "texir_hdr_read_host": (c_int, [c_char_p, _P, c_int, c_int]),
    "texir_png_unfilter_host": (c_int, [_P, c_int, c_int, c_int, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen libtexir_b200.so and type every entry point; raises TexirError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH) and "TEXIR_LIB" not in os.environ:
            try:  # a fresh checkout (the .so is git-ignored): compile it once with nvcc
                from . import build as _build
                _build.build_library()
            except Exception as exc:
                raise TexirError(
                    "libtexir_b200.so is not built (%s missing) and building it failed (%s): run "
                    "`python -m texir_code_b200.build` or __graft_entry__.build(); there is no CPU fallback"
                    % (SO_PATH, exc))
        if not os.path.exists(SO_PATH):
            raise TexirError("libtexir_b200.so not found at %s; there is no CPU fallback" % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().texir_last_error()
        raise TexirError("%s failed (code %d): %s" % (what or "texir call", rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(load().texir_launch_count())
