"""CPU model of the lock-step (while-while) BVH walk of one warp, on the benchmark scene: how many warp-level node
iterations and triangle iterations 32 rays cost when they are grouped the way the kernels group them, against the sum of
what the rays need individually.  Uses the device headers compiled for the CPU (tests/cpu_emul), so the node counts are the
kernel's own.   python profiles/simulate_warp_divergence.py [tris] [points]
Answers: where the 13-of-32 lane utilisation measured by ncu comes from, and whether another grouping would do better."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emul"))
import build as emul_build  # noqa: E402
from texir_code_b200 import synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
P = int(sys.argv[2]) if len(sys.argv) > 2 else 24
N, MAXSEG = 2048, 64
lib = emul_build.load()
sc = synth.make_room_scene(T, seed=666)
v = np.ascontiguousarray(sc.vertices, np.float32)
t = np.ascontiguousarray(sc.triangles, np.int32)
h = lib.emul_scene_create(v.ctypes.data, len(v), t.ctypes.data, len(t), None, 8)
pos, nrm = synth.sample_surface_points(sc, P, seed=3, seam_fraction=0.0)
shift = np.random.default_rng(1).random((P, 2)).astype(np.float32)
dirs = np.empty((P, N, 3), np.float32)
lib.emul_generate_dir(nrm.ctypes.data, None, shift.ctypes.data, P, N, 0, dirs.ctypes.data)
rays = np.concatenate([np.repeat(pos[:, None, :], N, 1), dirs], -1).reshape(-1, 6).astype(np.float32)
n_seg = np.empty(P * N, np.int32)
seg = np.zeros((P * N, MAXSEG), np.int32)
ltri = np.zeros((P * N, MAXSEG), np.int32)
lib.emul_trace_segments(h, rays.ctypes.data, P * N, MAXSEG, n_seg.ctypes.data, seg.ctypes.data, ltri.ctypes.data)
assert n_seg.max() <= MAXSEG
seg, ltri, n_seg = seg.reshape(P, N, MAXSEG), ltri.reshape(P, N, MAXSEG), n_seg.reshape(P, N)
per_ray_nodes, per_ray_tris = seg.sum(-1), ltri.sum(-1)
print("T = %d, %d points x %d samples: nodes/ray %.1f (std %.1f), leaf visits/ray %.2f, triangles/ray %.2f" % (
    T, P, N, per_ray_nodes.mean(), per_ray_nodes.std(), (n_seg - 1).mean(), per_ray_tris.mean()))


def warp_cost(idx):
    """idx [G,32] sample indices per warp iteration -> (warp node iterations, warp triangle iterations) per point"""
    s = seg[:, idx, :]                       # [P,G,32,MAXSEG]
    l = ltri[:, idx, :]
    return s.max(2).sum(-1).sum(-1), l.max(2).sum(-1).sum(-1), s.sum(-1).max(2).sum(-1)


print("%-34s %10s %10s %12s %10s" % ("grouping", "node it.", "lane use", "no-wait it.", "tri it."))
ideal_nodes = per_ray_nodes.sum(-1) / 32.0
for a in range(0, 7):
    M = N >> a
    idx = np.array([[(j << a) + c for j in range(k, k + 32)] for c in range(1 << a) for k in range(0, M, 32)])
    wn, wt, wmax = warp_cost(idx)
    print("patch bits a = %d (%4.2f x %5.3f)      %10.1f %9.1f%% %12.1f %10.1f" % (
        a, 32.0 * (1 << a) / N, 1.0 / (1 << a), wn.mean() / (N / 32), 100 * (ideal_nodes / wn).mean(), wmax.mean() / (N / 32),
        wt.mean() / (N / 32)))
rng = np.random.default_rng(0)
idx = rng.permutation(N).reshape(-1, 32)
wn, wt, wmax = warp_cost(idx)
print("random grouping                    %10.1f %9.1f%% %12.1f %10.1f" % (wn.mean() / (N / 32), 100 * (ideal_nodes / wn).mean(),
                                                                         wmax.mean() / (N / 32), wt.mean() / (N / 32)))
order = np.argsort(per_ray_nodes, axis=1)    # oracle grouping: rays sorted by their own length (upper bound on any grouping)
wn = np.stack([seg[p, order[p].reshape(-1, 32), :].max(1).sum(-1).sum(-1) for p in range(P)])
print("sorted by ray length (oracle)      %10.1f %9.1f%%" % (wn.mean() / (N / 32), 100 * (ideal_nodes / wn).mean()))
print("columns: node it. = warp-level node iterations per 32 rays under while-while (every round costs its slowest lane);")
print("         no-wait it. = max over the 32 lanes of their own total (what a loop without leaf reconvergence would cost);")
print("         lane use = sum of per-ray nodes / (32 x node it.)")

# ---- loop structures, simulated in lock step on the kernels' own grouping (a = 3) ----
a = 3
M = N >> a
idx = np.array([[(j << a) + c for j in range(k, k + 32)] for c in range(1 << a) for k in range(0, M, 32)])
rays_pn = rays.reshape(P, N, 6)
print()
print("%-46s %9s %9s %11s %10s" % ("loop structure (32 lanes, lock step)", "node it.", "tri it.", "nodes/ray", "tris/ray"))
for label, Q, stop in (("while-while (kernels)", 0, 0), ("1 postponed leaf, stop when all found", 1, 1),
                       ("1 postponed leaf, run until blocked", 1, 0), ("2 postponed leaves, stop when all found", 2, 1),
                       ("2 postponed leaves, run until blocked", 2, 0), ("4 postponed leaves, run until blocked", 4, 0)):
    tot = np.zeros(4, np.int64)
    out = np.zeros(4, np.int64)
    groups = 0
    for p in range(min(P, 8)):
        for g in idx[::4]:
            r = np.ascontiguousarray(rays_pn[p, g])
            lib.emul_simulate_warp(h, r.ctypes.data, 32, Q, stop, out.ctypes.data)
            tot += out
            groups += 1
    print("%-46s %9.1f %9.1f %11.1f %10.2f" % (label, tot[0] / groups, tot[1] / groups, tot[2] / groups / 32, tot[3] / groups / 32))

# ---- ray pools: all 2048 rays of a point handed to one warp in the kernels' order ----
order = idx.reshape(-1)
print()
print("%-46s %9s %9s %9s %12s   (per 32 rays)" % ("ray pool of one point (2048 rays, one warp)", "node it.", "tri it.", "rounds", "rounds w/ end"))
for label, mode in (("refill finished lanes at round boundaries", 0), ("two rays per lane, step whichever can", 1)):
    tot = np.zeros(4, np.int64)
    out = np.zeros(4, np.int64)
    npts = min(P, 6)
    for p in range(npts):
        r = np.ascontiguousarray(rays_pn[p, order])
        lib.emul_simulate_stream(h, r.ctypes.data, N, mode, out.ctypes.data)
        tot += out
    k = npts * N / 32
    print("%-46s %9.1f %9.1f %9.1f %12.1f" % (label, tot[0] / k, tot[1] / k, tot[2] / k, tot[3] / k))
