"""Builds tests/cpu_emul/libtexir_emul.so (test infrastructure: the kernels' host/device headers
compiled for the CPU).  -ffp-contract=off is mandatory: the headers' explicitly rounded operations
are plain C expressions in a host build."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "texir_code_b200", "csrc")
SO = os.path.join(HERE, "libtexir_emul.so")
SOURCES = [os.path.join(HERE, "emul.cpp"), os.path.join(CSRC, "bvh_build.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, h) for h in ("texir_math.h", "texir_bvh.h", "texir_walk.h", "texir_point.h",
                                                   "texir_post.h", "texir_tex.h", "bvh_build.h")]


def build(force: bool = False) -> str:
    stale = not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS)
    if force or stale:
        cmd = ["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
               "-o", SO] + SOURCES
        subprocess.check_call(cmd)
    return SO


def load():
    lib = ctypes.CDLL(build())
    vp, i64, i32, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
    lib.emul_scene_create.restype = vp
    lib.emul_scene_create.argtypes = [vp, i64, vp, i64, vp, i32]
    lib.emul_scene_destroy.argtypes = [vp]
    lib.emul_scene_stats.argtypes = [vp, vp]
    lib.emul_set_texture.argtypes = [vp, vp, i32, i32]
    lib.emul_cast_rays.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp, vp]
    lib.emul_generate_dir.argtypes = [vp, vp, vp, i64, i32, i32, vp]
    lib.emul_query_irf.argtypes = [vp, i32, vp, vp, i64, i32, vp]
    lib.emul_irradiance.argtypes = [vp, i32, vp, vp, vp, i64, i32, i32, vp]
    lib.emul_shade.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32, vp, vp, vp]
    lib.emul_cast_rays_fast.argtypes = [vp, i32, vp, i64, vp, vp, vp]
    lib.emul_pad_texture.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.emul_rgbe_encode.argtypes = [vp, i64, i32, vp]
    lib.emul_rgbe_decode.argtypes = [vp, i64, i32, vp]
    lib.emul_trace_segments.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.emul_simulate_warp.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.emul_simulate_stream.argtypes = [vp, vp, i32, i32, vp]
    lib.emul_cast_rays_front.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    lib.emul_cast_rays_free.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.emul_pop_stats.argtypes = [vp, vp, i64, vp]
    lib.emul_cast_rays_tnear.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.emul_mip_levels.argtypes = [i32, i32, i32]
    lib.emul_mip_elems.restype = i64
    lib.emul_mip_elems.argtypes = [i32, i32, i32, i32]
    lib.emul_mip_build.argtypes = [vp, i32, i32, i32, i32, vp]
    lib.emul_texture_fetch_mip.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, i64, vp, vp]
    return lib


if __name__ == "__main__":
    print(build(force=True))
