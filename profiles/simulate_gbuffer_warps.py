"""CPU model of the G-buffer walk (6 x 128^2 primary rays from the room centre, T = 2^20): warp-level node / triangle
iterations per warp under while-while for the pixel-to-lane mappings that were tried -- mean (throughput) and MAX (the
kernel has 0.65 waves: every warp is resident from the start, so its duration is its slowest warp's critical path).
python profiles/simulate_gbuffer_warps.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emul"))
import build as emul_build  # noqa: E402
from texir_code_b200 import synth  # noqa: E402

T, RES, MAXSEG = 1 << 20, 128, 256
lib = emul_build.load()
sc = synth.make_room_scene(T, seed=666)
v = np.ascontiguousarray(sc.vertices, np.float32)
t = np.ascontiguousarray(sc.triangles, np.int32)
h = lib.emul_scene_create(v.ctypes.data, len(v), t.ctypes.data, len(t), None, 8)
rays = np.ascontiguousarray(synth.cube_map_rays(RES), np.float32)
R = len(rays)
n_seg = np.empty(R, np.int32)
seg = np.zeros((R, MAXSEG), np.int32)
ltri = np.zeros((R, MAXSEG), np.int32)
lib.emul_trace_segments(h, rays.ctypes.data, R, MAXSEG, n_seg.ctypes.data, seg.ctypes.data, ltri.ctypes.data)
assert n_seg.max() <= MAXSEG, n_seg.max()
nodes, tris = seg.sum(-1), ltri.sum(-1)
print("%d primary rays: nodes/ray mean %.1f  p50 %d  p99 %d  max %d; leaf visits/ray %.2f (max %d); triangles/ray %.2f"
      % (R, nodes.mean(), np.median(nodes), np.percentile(nodes, 99), nodes.max(), (n_seg - 1).mean(), n_seg.max() - 1, tris.mean()))
pix = np.arange(R).reshape(6, RES, RES)


def groups(kind, lanes):
    if kind == "row":
        return pix.reshape(-1, lanes)
    th, tw = {32: (4, 8), 16: (4, 4), 8: (2, 4), 4: (2, 2)}[lanes]
    return pix.reshape(6, RES // th, th, RES // tw, tw).transpose(0, 1, 3, 2, 4).reshape(-1, lanes)


print("%-28s %8s | node it. mean / p99 / max | tri it. mean / max | est. cycles of the slowest warp" % ("mapping", "warps"))
for kind, lanes in (("row", 32), ("tile", 32), ("row", 16), ("tile", 16), ("row", 8), ("tile", 8), ("tile", 4)):
    g = groups(kind, lanes)
    ni = seg[g].max(1).sum(-1)          # every round costs its slowest lane
    ti = ltri[g].max(1).sum(-1)
    est = ni * 350 + ti * 600           # ~ L2-latency node step, triangle test: dependent-issue bound at 5 warps/scheduler
    print("%-4s %2d lanes per warp       %8d | %8.1f / %6.0f / %5d | %6.1f / %4d | %8d"
          % (kind, lanes, len(g), ni.mean(), np.percentile(ni, 99), ni.max(), ti.mean(), ti.max(), est.max()))
