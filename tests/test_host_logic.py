"""CPU: host-side logic of the package that needs no GPU -- Hammersley helpers, the synthetic scene
generator, point sharding, and the multi-process (gloo, world_size 2) gather / gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden


def test_hammersley_helpers_match_reference_goldens():
    from texir_code_b200 import sample_util as su
    g = load_golden("hammersley")
    for i, want in zip(g["ri_in"], g["ri_out"]):
        assert su.RadicalInverse(int(i)) == want
    assert np.array_equal(np.array([su.Hammersley(i, 16) for i in range(16)]), g["ham16"])
    assert np.array_equal(su.hammersley_table(16), g["ham16"].astype(np.float32))
    t = su.hammersley_table(2048)
    assert t.shape == (2048, 2) and len(np.unique(t[:, 1])) == 2048 and t.min() >= 0 and t.max() < 1


def test_synthetic_scene_is_a_closed_room_with_exact_counts():
    from texir_code_b200 import synth
    for T in (3000, 20000):
        sc = synth.make_room_scene(T, clutter=6, tex_res=256)
        assert sc.triangles.shape == (T, 3) and sc.triangle_uvs.shape == (3 * T, 2)
        assert sc.triangle_uvs.dtype == np.float64 and sc.vertices.dtype == np.float32
        assert sc.triangles.min() >= 0 and sc.triangles.max() < len(sc.vertices)
        assert 0 <= sc.triangle_uvs.min() and sc.triangle_uvs.max() <= 1
        assert np.all(sc.tri_areas > 0)
    a = synth.make_room_scene(3000, clutter=6, tex_res=256)
    b = synth.make_room_scene(3000, clutter=6, tex_res=256)
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.triangles, b.triangles)
    tex = synth.make_radiance_texture(64)
    assert tex.shape == (1, 3, 64, 64) and tex.dtype == np.float32 and tex.min() > 0
    pos, nrm = synth.sample_surface_points(a, 2048, seam_fraction=0.05)
    seam = np.all(nrm == 0, axis=1)
    assert 0 < seam.sum() < 300 and np.all(pos[seam] == 0)
    assert np.allclose(np.linalg.norm(nrm[~seam], axis=1), 1, atol=1e-5)
    rays = synth.cube_map_rays(8)
    assert rays.shape == (6 * 64, 6) and np.allclose(np.linalg.norm(rays[:, 3:], axis=1), 1, atol=1e-6)


def test_shard_bounds_cover_and_align():
    from texir_code_b200.parallel import shard_bounds, shard_sizes
    for P, W, al in ((1048576, 8, 512), (98304, 4, 1), (1000, 3, 512), (7, 8, 1), (0, 2, 1), (513, 2, 512)):
        spans = [shard_bounds(P, r, W, al) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == P
        for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
            assert e0 == b1 and b0 <= e0
        assert all(b % al == 0 or b == P for b, _ in spans)
        assert sum(shard_sizes(P, W, al)) == P
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from texir_code_b200.parallel import allreduce_gradients, gather_points, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = 1536 + 200  # ragged: rank 0 gets 1024, rank 1 gets 712
        full = torch.arange(P * 3, dtype=torch.float32).reshape(P, 3)
        b, e = shard_bounds(P, rank, world, 512)
        out = gather_points(full[b:e].clone(), P, 512)
        ok_gather = torch.equal(out, full)
        # texture-gradient all-reduce: two "textures", one flat buffer
        a = torch.nn.Parameter(torch.zeros(8, 8, 3))
        r = torch.nn.Parameter(torch.zeros(16, 16, 1))
        a.grad = torch.full_like(a, float(rank + 1))
        r.grad = torch.full_like(r, 10.0 * (rank + 1))
        nbytes = allreduce_gradients([a, r])
        ok_red = bool((a.grad == 3).all() and (r.grad == 30).all() and nbytes == (8 * 8 * 3 + 16 * 16) * 4)
        q.put((rank, ok_gather, ok_red))
    finally:
        dist.destroy_process_group()


def test_gather_and_gradient_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True, True), (1, True, True)]


def test_single_process_parallel_helpers_are_noops():
    from texir_code_b200.parallel import allreduce_gradients, gather_points
    x = torch.ones(5, 3)
    assert gather_points(x, 5) is x
    p = torch.nn.Parameter(torch.zeros(2))
    p.grad = torch.ones(2)
    assert allreduce_gradients([p]) == 0


def _reference_cube_mvps(c2w: np.ndarray):
    """datasets/dataset.py:419-463: six world-to-camera matrices from one extrinsic, times projection(), transposed"""
    import torch
    ext = torch.from_numpy(c2w).float()
    faces = []
    f0 = ext.clone(); R, F = f0[0:3, 0].clone(), f0[0:3, 2].clone(); f0[0:3, 2] = -R; f0[0:3, 0] = F; faces.append(torch.inverse(f0))
    faces.append(torch.inverse(ext))
    f2 = ext.clone(); R, F = f2[0:3, 0].clone(), f2[0:3, 2].clone(); f2[0:3, 2] = R; f2[0:3, 0] = -F; faces.append(torch.inverse(f2))
    f3 = ext.clone(); R, F = f3[0:3, 0].clone(), f3[0:3, 2].clone(); f3[0:3, 2] = -F; f3[0:3, 0] = -R; faces.append(torch.inverse(f3))
    f4 = ext.clone(); F, R = f4[0:3, 2].clone(), f4[0:3, 0].clone(); U = torch.linalg.cross(R, F); f4[0:3, 2] = U; f4[0:3, 1] = F; faces.append(torch.inverse(f4))
    f5 = ext.clone(); F, R = f5[0:3, 2].clone(), f5[0:3, 0].clone(); U = torch.linalg.cross(R, F); f5[0:3, 2] = -U; f5[0:3, 1] = -F; faces.append(torch.inverse(f5))
    w2c = torch.stack(faces)
    n, f = 1e-4, 100.0
    proj = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, (f + n) / (f - n), -(2 * f * n) / (f - n)], [0, 0, 1, 0]])
    mvp = torch.einsum("ijk,ikl->ijl", proj.expand(6, 4, 4), w2c)
    return torch.transpose(mvp, 1, 2).float(), ext[0:3, -1].clone()


def test_rays_from_mvp_invert_the_reference_cameras():
    """every ray of rays_from_mvp projects back to its own pixel centre through the reference's mvp, the six faces
    cover the sphere, and the differentials match central differences between neighbouring pixels"""
    import torch
    from texir_code_b200.models import rays_from_mvp
    rng = np.random.default_rng(4)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, :3] = q
    c2w[:3, 3] = [0.4, -0.3, 1.1]
    mvp, cam = _reference_cube_mvps(c2w)
    res = 16
    rays, dd = rays_from_mvp(mvp, cam, res)
    assert rays.shape == (6, res, res, 6) and dd.shape == (6, res, res, 6)
    assert torch.allclose(rays[..., :3], cam.expand(6, res, res, 3))
    assert torch.allclose(torch.linalg.norm(rays[..., 3:], dim=-1), torch.ones(6, res, res), atol=1e-6)
    pts = torch.cat([rays[..., :3].double() + 3.0 * rays[..., 3:].double(), torch.ones(6, res, res, 1, dtype=torch.float64)], -1)
    clip = torch.einsum("vhwk,vkl->vhwl", pts, mvp.double())
    ndc = clip[..., :2] / clip[..., 3:4]
    lin = (torch.arange(res, dtype=torch.float64) + 0.5) / res * 2 - 1
    assert torch.allclose(ndc[..., 0], lin.view(1, 1, res).expand(6, res, res), atol=1e-4)   # X = column
    assert torch.allclose(ndc[..., 1], lin.view(1, res, 1).expand(6, res, res), atol=1e-4)   # Y = row, row 0 at -1
    assert bool((clip[..., 3] > 0).all())                                                    # in front of the camera
    d = rays[..., 3:].double()
    assert torch.linalg.norm(d.reshape(-1, 3).mean(0)) < 1e-6                                # faces tile the sphere
    fdx = (d[:, :, 2:] - d[:, :, :-2]) / 2
    fdy = (d[:, 2:] - d[:, :-2]) / 2
    scale = float(dd.abs().max())
    assert float((fdx - dd[:, :, 1:-1, :3]).abs().max()) < 0.02 * scale
    assert float((fdy - dd[:, 1:-1, :, 3:]).abs().max()) < 0.02 * scale


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours) needs no GPU: stdout must carry exactly one
    JSON line with the contract keys, whatever libraries print meanwhile"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--tris", "20000", "--tex", "256"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ray-samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


# ---------------------------------------------------------------------------------------------------
# round 2: block-cyclic sharding, the flat gradient buffer, the IrT producer's index arithmetic
# ---------------------------------------------------------------------------------------------------
def test_block_cyclic_shares_partition_the_items():
    from texir_code_b200.parallel import block_cyclic_index, local_pixels
    for num, world, block in ((1 << 14, 4, 512), (5000, 3, 512), (700, 8, 128), (6, 2, 4)):
        parts = [block_cyclic_index(num, r, world, block) for r in range(world)]
        allidx = torch.cat(parts)
        assert torch.equal(torch.sort(allidx).values, torch.arange(num))           # every item exactly once
        for r, p in enumerate(parts):
            assert bool(((p // block) % world == r).all()) and bool((p[1:] > p[:-1]).all())
        if num % (block * world) == 0:
            assert len({len(p) for p in parts}) == 1                               # equal shares
    img = torch.arange(6 * 16 * 16 * 2, dtype=torch.float32).reshape(6, 16, 16, 2)  # 256 pixels per view = 2 blocks of 128
    a, b = local_pixels(img, 0, 2, 6), local_pixels(img, 1, 2, 6)
    assert a.shape == (6 * 128, 2) and torch.equal(a[:128], img[0].reshape(-1, 2)[:128]) and torch.equal(b[:128], img[0].reshape(-1, 2)[128:])
    assert torch.equal(local_pixels(img, 0, 1, 6), img.reshape(-1, 2))
    with pytest.raises(ValueError):
        local_pixels(img, 0, 4, 6)                                                 # 256 is not a multiple of 128 * 4


def _worker_flat(rank, world, port, q):
    import torch.distributed as dist
    from texir_code_b200.parallel import _flat_view_of, allreduce_gradients
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class TwoGrads(torch.autograd.Function):      # what ops._MatStage.backward does: both gradients are views of ONE buffer
            @staticmethod
            def forward(ctx, a, r):
                return a.sum() + r.sum()

            @staticmethod
            def backward(ctx, g):
                flat = torch.full((8 * 8 * 3 + 16 * 16,), float(rank + 1))
                TwoGrads.flat = flat
                return flat[:192].view(8, 8, 3), flat[192:].view(16, 16, 1)

        a = torch.nn.Parameter(torch.zeros(8, 8, 3))
        r = torch.nn.Parameter(torch.zeros(16, 16, 1))
        TwoGrads.apply(a, r).backward()
        flat = _flat_view_of([a.grad, r.grad])
        shares = flat is not None and flat.data_ptr() == TwoGrads.flat.data_ptr() and flat.numel() == 448
        before = a.grad.data_ptr()
        nbytes = allreduce_gradients([a, r])
        ok = bool((a.grad == 3).all() and (r.grad == 3).all()) and nbytes == 448 * 4 and a.grad.data_ptr() == before \
            and bool((TwoGrads.flat == 3).all())                                   # reduced in place, no copy
        other = _flat_view_of([a.grad, torch.ones(3)]) is None and _flat_view_of([a.grad]) is not None
        q.put((rank, bool(shares), ok, other))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_buffer_is_reduced_in_place_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_flat, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True, True, True), (1, True, True, True)]


def test_seam_mask_sums_in_the_arrays_own_dtype():
    """tracer_o3d_irt.py:138,177 adds the three uint16 channels of the index texture in uint16 (numpy): 30000 + 30000 +
    5536 wraps to 0 and counts as a seam there, so it must here"""
    from texir_code_b200.models import _seam_mask
    idx = np.zeros((2, 3, 3), np.uint16)
    idx[0, 1] = (30000, 30000, 5536)      # wraps to 0
    idx[1, 2] = (1, 0, 0)
    idx[1, 0] = (50000, 20000, 3)
    ref = (idx[:, :, 0] + idx[:, :, 1] + idx[:, :, 2]) == 0
    assert np.array_equal(_seam_mask(idx, "cpu").numpy(), ref) and ref[0, 1] and not ref[1, 2]
    assert np.array_equal(_seam_mask(idx.astype(np.int64), "cpu").numpy(), idx.astype(np.int64).sum(-1) == 0)


def test_position_normal_texture_lookup_matches_the_reference_numpy():
    """calcute_position_normal_texture (tracer_o3d_irt.py:115-142): index -> panorama pixel with the reference's
    arithmetic ((idx / 50000 * size).astype(int), clip), per view, seams zeroed -- restated here in numpy"""
    rng = np.random.default_rng(2)
    H = W = 24
    views = [rng.normal(size=(16, 32, 6)).astype(np.float32) for _ in range(3)]
    idx = np.zeros((H, W, 3), np.uint16)
    idx[..., 0] = rng.integers(0, 50001, (H, W))
    idx[..., 1] = rng.integers(0, 50001, (H, W))
    idx[..., 2] = rng.integers(0, 3, (H, W))
    idx[3, 4] = 0
    idx[10, 0] = (50000, 50000, 2)
    pos_ref = np.zeros((H, W, 3), np.float32)
    nrm_ref = np.zeros((H, W, 3), np.float32)
    for v in np.unique(idx[:, :, 2]):
        rows, cols = np.where(idx[:, :, 2] == v)
        hdr = views[v]
        h, w, _ = hdr.shape
        c = np.clip((idx[rows, cols, 1] / 50000 * w).astype(int), 0, w - 1)
        r = np.clip((idx[rows, cols, 0] / 50000 * h).astype(int), 0, h - 1)
        pos_ref[rows, cols] = hdr[r, c, 0:3]
        nrm_ref[rows, cols] = hdr[r, c, 3:6]
    seam = (idx[:, :, 0] + idx[:, :, 1] + idx[:, :, 2]) == 0
    pos_ref[seam] = 0
    nrm_ref[seam] = 0

    from texir_code_b200 import models

    class Stub(models.TracerO3d):          # the lookup needs no scene: give it a CPU "device"
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.scene = type("S", (), {"device": torch.device("cpu")})()
            self.position_texture = self.normal_texture = self.index_texture = None

    t = Stub()
    pos, nrm = t.calcute_position_normal_texture([torch.from_numpy(v) for v in views], idx)
    assert np.array_equal(pos.numpy(), pos_ref) and np.array_equal(nrm.numpy(), nrm_ref)
    assert t.position_texture is pos and t.index_texture is idx
    with pytest.raises(ValueError):
        Stub().forward()                                                          # nothing to trace yet


def test_draw_shift_is_the_reference_stream_on_any_path():
    from texir_code_b200.sample_util import draw_shift
    torch.manual_seed(11)
    a = torch.rand(7, 1, 2)
    b = torch.rand(5, 1, 2)
    torch.manual_seed(11)
    assert torch.equal(draw_shift(7, "cpu"), a) and torch.equal(draw_shift(5, "cpu"), b)


def test_mt19937_advance_host_is_where_torch_rand_leaves_the_generator():
    """texir_mt19937_advance_host (the host half of the graphed shift draw): state, left and next after n draws must be
    torch's own after torch.rand(n) -- from a fresh seed, mid-block, on block boundaries, for 0 and 196 608 draws"""
    import ctypes
    import torch
    from texir_code_b200 import _lib
    lib = _lib.load()
    for seed, pre, n in ((1, 0, 5000), (7, 333, 196608), (9, 623, 1), (3, 624, 625), (4, 100, 3), (5, 0, 0), (6, 1247, 624)):
        torch.manual_seed(seed)
        if pre:
            torch.rand(pre)
        st = torch.get_rng_state().numpy().copy()
        state = st[24:24 + 8 * 624].view(np.uint64).astype(np.uint32)
        left = ctypes.c_int(int(st[8:12].view(np.int32)[0]))
        nxt = ctypes.c_int(int(st[16:24].view(np.uint64)[0]))
        if n:
            torch.rand(n)
        assert lib.texir_mt19937_advance_host(state.ctypes.data, ctypes.byref(left), ctypes.byref(nxt), n) == 0
        st2 = torch.get_rng_state().numpy()
        assert np.array_equal(state, st2[24:24 + 8 * 624].view(np.uint64).astype(np.uint32))
        assert left.value == int(st2[8:12].view(np.int32)[0]) and nxt.value == int(st2[16:24].view(np.uint64)[0])
    bad = ctypes.c_int(9999)
    assert lib.texir_mt19937_advance_host(state.ctypes.data, ctypes.byref(bad), ctypes.byref(nxt), 5) != 0
