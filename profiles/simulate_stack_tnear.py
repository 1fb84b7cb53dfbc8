"""CPU model: how many node steps / triangle tests a traversal stack that ALSO stores each postponed child's entry
distance would save (a pop whose stored tnear already exceeds the closest hit is dropped at once instead of costing a
node step that finds neither child, or a leaf's triangle tests).   python profiles/simulate_stack_tnear.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emul"))
import build as emul_build  # noqa: E402
from texir_code_b200 import synth  # noqa: E402

lib = emul_build.load()
T = 1 << 20
sc = synth.make_room_scene(T, seed=666)
v = np.ascontiguousarray(sc.vertices, np.float32)
t = np.ascontiguousarray(sc.triangles, np.int32)
h = lib.emul_scene_create(v.ctypes.data, len(v), t.ctypes.data, len(t), None, 8)


def report(label, rays):
    rays = np.ascontiguousarray(rays, np.float32)
    out = np.zeros(6, np.int64)
    lib.emul_pop_stats(h, rays.ctypes.data, len(rays), out.ctypes.data)
    R = len(rays)
    print("%-34s node steps/ray %.1f | inner pops %.2f, dead on arrival %.2f (%.1f %% of the node steps) | leaf pops %.2f, dead %.2f | "
          "triangle tests %.2f" % (label, out[0] / R, out[1] / R, out[2] / R, 100.0 * out[2] / out[0], out[3] / R, out[4] / R, out[5] / R))


P, N = 48, 2048
pos, nrm = synth.sample_surface_points(sc, P, seed=3, seam_fraction=0.0)
shift = np.random.default_rng(1).random((P, 2)).astype(np.float32)
dirs = np.empty((P, N, 3), np.float32)
lib.emul_generate_dir(nrm.ctypes.data, None, shift.ctypes.data, P, N, 0, dirs.ctypes.data)
report("IrT rays (uniform hemisphere)", np.concatenate([np.repeat(pos[:, None, :], N, 1), dirs], -1).reshape(-1, 6))
report("primary rays (cube map 64^2)", synth.cube_map_rays(64))
