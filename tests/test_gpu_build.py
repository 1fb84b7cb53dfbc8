"""GPU (-m gpu): the device BVH builder (csrc/bvh_build_gpu.cu) against the host builder (csrc/bvh_build.cpp) and
the brute-force oracle.  Both builders run the same binned-SAH algorithm with the same arithmetic, so they must
produce THE SAME TREE: equal node count, depth and SAH cost, and -- ray by ray -- equal hits AND equal traversal
counters (nodes visited, triangles tested).  Degenerate inputs take the builder's fallback paths (identical
centroids -> median split by position, boxes recomputed by the children); there only the hits must agree."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tx(cuda_device):
    import texir_code_b200 as tx
    tx._lib.load()
    return tx


def _rays(sc, n, seed):
    from texir_code_b200 import synth
    P = n // 8
    pos, nrm = synth.sample_surface_points(sc, P, seed=seed, seam_fraction=0.0)
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(P, 8, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d *= np.sign((d * nrm[:, None, :]).sum(-1, keepdims=True))
    return np.ascontiguousarray(np.concatenate([np.broadcast_to(pos[:, None, :], d.shape), d], -1).reshape(-1, 6), np.float32)


@pytest.mark.parametrize("num_tris,clutter", [(4000, 6), (20000, 12), (200000, 32), (1 << 20, 64)])
def test_device_builder_builds_the_host_builders_tree(tx, cuda_device, num_tris, clutter):
    from texir_code_b200 import synth
    sc = synth.make_room_scene(num_tris, seed=666, clutter=clutter, tex_res=256)
    t0 = time.time()
    host = tx.Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=cuda_device, builder="host")
    t1 = time.time()
    dev = tx.Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=cuda_device, builder="device")
    t2 = time.time()
    hi, di = host.info(), dev.info()
    print("T=%d: host %.3f s, device %.3f s (create, wall); nodes %d depth %d" % (num_tris, t1 - t0, t2 - t1, di["num_nodes2"], di["depth2"]))
    assert hi["builder"] == 1 and di["builder"] == 2
    assert di["num_nodes2"] == hi["num_nodes2"] and di["depth2"] == hi["depth2"]
    assert abs(di["sah2"] - hi["sah2"]) <= 1e-9 * hi["sah2"]
    rays = torch.from_numpy(_rays(sc, 40000, 5)).to(cuda_device)
    a, b = host.cast_rays(rays, stats=True), dev.cast_rays(rays, stats=True)
    for k in ("t_hit", "primitive_ids", "nodes_visited", "tris_tested"):
        assert torch.equal(a[k], b[k]), k
    a, b = host.cast_rays(rays), dev.cast_rays(rays)            # the quantised 32-byte nodes the kernels walk
    for k in ("t_hit", "primitive_ids", "primitive_uvs"):
        assert torch.equal(a[k], b[k]), k
    assert float(torch.isfinite(a["t_hit"]).float().mean()) > 0.99
    dev.set_bvh(8)                                               # the 8-wide layout is built lazily, on the host
    c = dev.cast_rays(rays)
    assert torch.equal(c["t_hit"], a["t_hit"]) and torch.equal(c["primitive_ids"], a["primitive_ids"])
    host.close()
    dev.close()


def test_device_builder_hits_are_the_brute_force_hits(tx, oracle, cuda_device, scene_golden):
    name, g = scene_golden
    dev = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device, builder="device")
    assert dev.info()["builder"] == 2
    rng = np.random.default_rng(3)
    o = rng.uniform(-1.0, 1.0, (20000, 3)) * np.array([3.5, 1.2, 2.5])
    d = rng.normal(size=(20000, 3))
    rays = np.concatenate([o, d / np.linalg.norm(d, axis=1, keepdims=True)], 1).astype(np.float32)
    got = dev.cast_rays(torch.from_numpy(rays).to(cuda_device))
    t, i, uv = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays)
    assert np.array_equal(got["primitive_ids"].cpu().numpy().view(np.uint32), i)
    assert np.array_equal(got["t_hit"].cpu().numpy(), t)
    assert np.array_equal(got["primitive_uvs"].cpu().numpy(), uv)


def test_device_builder_degenerate_inputs(tx, cuda_device):
    rng = np.random.default_rng(0)
    # 5000 copies of the same three triangles (identical centroids: the SAH finds no split -> median by position)
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [2, 0, 0], [3, 0, 0], [2, 1, 0.5]], np.float32)
    verts = np.tile(base, (5000, 1))
    tris = np.arange(len(verts), dtype=np.int32).reshape(-1, 3)
    # a flat sheet (zero extent in z) of 8192 triangles
    n = 64
    gx, gy = np.meshgrid(np.linspace(0, 1, n + 1), np.linspace(0, 1, n + 1))
    fv = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], 1).astype(np.float32)
    j, i = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    v00 = (j * (n + 1) + i).ravel()
    ft = np.concatenate([np.stack([v00, v00 + 1, v00 + n + 2], 1), np.stack([v00, v00 + n + 2, v00 + n + 1], 1)]).astype(np.int32)
    # a small scene forced onto the device builder
    sv = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    stri = rng.integers(0, 300, (100, 3)).astype(np.int32)
    for verts_, tris_ in ((verts, tris), (fv, ft), (sv, stri)):
        host = tx.Scene(verts_, tris_, None, device=cuda_device, builder="host")
        dev = tx.Scene(verts_, tris_, None, device=cuda_device, builder="device")
        o = rng.uniform(-0.5, 2.5, (8000, 3)).astype(np.float32)
        d = rng.normal(size=(8000, 3)).astype(np.float32)
        rays = torch.from_numpy(np.concatenate([o, d], 1)).to(cuda_device)
        a, b = host.cast_rays(rays), dev.cast_rays(rays)
        for k in ("t_hit", "primitive_ids", "primitive_uvs"):
            assert torch.equal(a[k], b[k]), k
        assert int((a["primitive_ids"] >= 0).sum()) > 50
    bad = np.array([[0, 1, 99999]], np.int32).repeat(5000, 0)
    with pytest.raises(tx.TexirError):
        tx.Scene(sv, bad, None, device=cuda_device, builder="device")
    nanv = sv.copy()
    nanv[7, 1] = np.nan
    with pytest.raises(tx.TexirError):
        tx.Scene(nanv, np.tile(stri, (50, 1)), None, device=cuda_device, builder="device")
