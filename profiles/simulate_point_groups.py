"""CPU model: what would a warp cost if its 32 lanes held the SAME direction cell of 32 NEIGHBOURING points (Morton
order) instead of 32 directions of one point's patch?  Hammersley's 2048 samples fill a 64 x 32 grid of boxes of the unit
square exactly once (i = u * 32 + rev5(v)); a point's Cranley-Patterson shift moves its boxes by (64 a, 32 b), so lane l
takes the box (U - round(64 a_l), V - round(32 b_l)) of its own point: every (point, sample) pair is still visited exactly
once, and the 32 rays of a warp agree in direction to about a box.   python profiles/simulate_point_groups.py [groups]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emul"))
import build as emul_build  # noqa: E402
from texir_code_b200 import synth  # noqa: E402

T, PALL, N, MAXSEG = 1 << 20, 1 << 20, 2048, 64
GROUPS = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = emul_build.load()
sc = synth.make_room_scene(T, seed=666)
v = np.ascontiguousarray(sc.vertices, np.float32)
t = np.ascontiguousarray(sc.triangles, np.int32)
h = lib.emul_scene_create(v.ctypes.data, len(v), t.ctypes.data, len(t), None, 8)
pos_all, nrm_all = synth.sample_surface_points(sc, PALL, seed=3, seam_fraction=0.0)


def morton(p):
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / (hi - lo) * 1023.0), 0, 1023).astype(np.uint64)
    def spread(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


order = np.argsort(morton(pos_all.astype(np.float64)), kind="stable")
rng = np.random.default_rng(5)
starts = rng.integers(0, PALL // 32, GROUPS) * 32
sel = np.concatenate([order[s:s + 32] for s in starts])
pos, nrm = np.ascontiguousarray(pos_all[sel]), np.ascontiguousarray(nrm_all[sel])
P = len(sel)
shift = rng.random((P, 2)).astype(np.float32)
dirs = np.empty((P, N, 3), np.float32)
lib.emul_generate_dir(nrm.ctypes.data, None, shift.ctypes.data, P, N, 0, dirs.ctypes.data)
rays = np.concatenate([np.repeat(pos[:, None, :], N, 1), dirs], -1).reshape(-1, 6).astype(np.float32)
n_seg = np.empty(P * N, np.int32)
seg = np.zeros((P * N, MAXSEG), np.int32)
ltri = np.zeros((P * N, MAXSEG), np.int32)
lib.emul_trace_segments(h, rays.ctypes.data, P * N, MAXSEG, n_seg.ctypes.data, seg.ctypes.data, ltri.ctypes.data)
assert n_seg.max() <= MAXSEG
seg, ltri = seg.reshape(P, N, MAXSEG), ltri.reshape(P, N, MAXSEG)
nodes = seg.sum(-1)
print("%d groups of 32 Morton-neighbouring points (spread within a group: %.1f cm median), nodes/ray %.1f"
      % (GROUPS, 100 * np.median([np.linalg.norm(pos[g * 32:(g + 1) * 32].max(0) - pos[g * 32:(g + 1) * 32].min(0)) for g in range(GROUPS)]),
         nodes.mean()))
same_n = np.mean([(np.abs(nrm[g * 32:(g + 1) * 32] @ nrm[g * 32]) > 0.999).mean() for g in range(GROUPS)])
print("lanes of a group whose normal agrees with lane 0's to 2.5 degrees: %.0f %%" % (100 * same_n))

# (A) today's grouping: one point per warp, 32 samples of a patch (a = 3)
a = 3
M = N >> a
idxA = np.array([[(j << a) + c for j in range(k, k + 32)] for c in range(1 << a) for k in range(0, M, 32)])   # [64, 32]
sA = seg[:, idxA, :]                                     # [P, 64, 32, MAXSEG]
nodeA = sA.max(2).sum(-1).mean()
triA = ltri[:, idxA, :].max(2).sum(-1).mean()
# (B) 32 points per warp, one direction box per round
rev5 = np.array([int(format(i, "05b")[::-1], 2) for i in range(32)])
ka, kb = np.rint(shift[:, 0] * 64).astype(int), np.rint(shift[:, 1] * 32).astype(int)
U, V = np.meshgrid(np.arange(64), np.arange(32), indexing="ij")
U, V = U.reshape(-1), V.reshape(-1)                      # 2048 rounds
nodeB, triB = [], []
for g in range(GROUPS):
    pts = np.arange(g * 32, (g + 1) * 32)
    u = (U[:, None] - ka[None, pts]) % 64                # [2048, 32]
    vv = (V[:, None] - kb[None, pts]) % 32
    i = u * 32 + rev5[vv]
    s = seg[pts[None, :], i, :]                          # [2048, 32, MAXSEG]
    l = ltri[pts[None, :], i, :]
    nodeB.append(s.max(1).sum(-1).mean())
    triB.append(l.max(1).sum(-1).mean())
    assert np.array_equal(np.sort(i, axis=0), np.tile(np.arange(N)[:, None], (1, 32)))   # every sample once per point
ideal = nodes.mean()
print("%-58s %9s %9s %9s" % ("grouping", "node it.", "lane use", "tri it."))
print("%-58s %9.1f %8.1f%% %9.1f" % ("one point per warp, 32 samples of a patch (kernels)", nodeA, 100 * ideal / nodeA, triA))
print("%-58s %9.1f %8.1f%% %9.1f" % ("32 neighbouring points per warp, one direction box", np.mean(nodeB), 100 * ideal / np.mean(nodeB), np.mean(triB)))
print("per group (B):", " ".join("%.0f" % x for x in nodeB))
