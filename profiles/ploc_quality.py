"""Tree quality of PLOC (the device builder's algorithm) against the host binned-SAH builder, on the benchmark scene:
SAH cost, nodes visited / triangles tested per ray (IrT-like rays: surface points, uniform hemisphere directions), depth.
    python profiles/ploc_quality.py [log2 triangles] [radius ...]"""
import ctypes, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from texir_code_b200 import synth

lt = int(sys.argv[1]) if len(sys.argv) > 1 else 20
radii = [int(x) for x in sys.argv[2:]] or [8, 16, 32]
so = "/tmp/libploc_proto.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread",
                       "-I", os.path.join(ROOT, "texir_code_b200", "csrc"), os.path.join(ROOT, "profiles", "ploc_prototype.cpp"), "-o", so])
lib = ctypes.CDLL(so)
VP = ctypes.c_void_p
lib.ploc_quality.argtypes = [VP, ctypes.c_int64, VP, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, VP,
                             ctypes.c_int64, VP, VP, VP]
sc = synth.make_room_scene(1 << lt, seed=666, tex_res=4096)
v = np.ascontiguousarray(sc.vertices, np.float32)
t = np.ascontiguousarray(sc.triangles, np.int32)
P, N = 2048, 16
pos, nrm = synth.sample_surface_points(sc, P, seed=5, seam_fraction=0.0)
rng = np.random.default_rng(1)
d = rng.normal(size=(P, N, 3))
d /= np.linalg.norm(d, axis=-1, keepdims=True)
d *= np.sign((d * nrm[:, None, :]).sum(-1, keepdims=True))        # into the hemisphere of the normal
rays = np.ascontiguousarray(np.concatenate([np.broadcast_to(pos[:, None, :], d.shape), d], -1).reshape(-1, 6), np.float32)
R = rays.shape[0]
res = {}
for mode, radius, leaf in [(0, 0, 3)] + [(1, r, 3) for r in radii]:
    out = (ctypes.c_double * 5)()
    prim = np.empty(R, np.uint32)
    th = np.empty(R, np.float32)
    t0 = time.time()
    lib.ploc_quality(v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], mode, radius, leaf, 8, rays.ctypes.data, R,
                     ctypes.addressof(out), prim.ctypes.data, th.ctypes.data)
    res[(mode, radius)] = (prim, th)
    name = "host binned SAH" if mode == 0 else "PLOC r=%d" % radius
    print("%-16s sah %8.3f  nodes/ray %6.2f  tris/ray %5.2f  depth %3d  inner nodes %8d   (%.1f s)" % (
        name, out[0], out[1], out[2], int(out[3]), int(out[4]), time.time() - t0), flush=True)
ref = res[(0, 0)]
for k, (prim, th) in res.items():
    print(k, "hits identical to the SAH tree:", bool(np.array_equal(prim, ref[0]) and np.array_equal(th, ref[1])))
