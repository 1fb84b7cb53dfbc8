"""GPU (-m gpu): the per-step forms of the Mat stage (csrc/texir_matstage.cu, ops.mat_stage) and the pixel / texel
sharding of the multi-GPU path, through the package's binding of the C ABI.

The few-launch kernels are checked BIT FOR BIT against the level-by-level kernels they replace (same operands in the
same order per texel); ops.mat_stage against the op-by-op composition texture_fetch -> shade -> autograd, including
the direct gradients the reference's loss puts on preds['albedo'] / ['roughness'] / ['roughness_womipmap']
(models/loss.py:94,103,113); tolerance 1e-5 rel-L2 on gradients (float atomics arrive in arbitrary order)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.fixture(scope="module")
def tx(cuda_device):
    import texir_code_b200 as tx
    tx._lib.load()
    return tx


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@pytest.mark.parametrize("shape,max_level", [((64, 64, 3), None), ((256, 256, 3), 13), ((128, 96, 1), 5), ((32, 32, 1), None),
                                             ((64, 32, 3), 3), ((2048, 2048, 3), 13), ((4096, 4096, 1), 13), ((12, 8, 3), 2),
                                             ((1, 16, 3), None), ((8, 4, 1), None), ((6, 6, 3), 1), ((16, 16, 2), None),
                                             ((64, 16, 1), None), ((16, 64, 1), 4), ((32, 48, 1), 4), ((1, 32, 1), None)])
def test_fast_pyramid_build_and_fold_equal_the_levelwise_kernels(tx, cuda_device, shape, max_level):
    H, W, C = shape
    lib = tx._lib.load()
    L = tx.ops.mip_levels(H, W, max_level)
    n = max(int(lib.texir_texture_mip_elems(H, W, C, L)), 1)
    g = torch.Generator(device="cpu").manual_seed(H * 7 + W + C)
    tex = torch.rand(H, W, C, generator=g).to(cuda_device)
    a = torch.full((n,), float("nan"), device=cuda_device)
    b = torch.full((n,), float("nan"), device=cuda_device)
    st = _stream(cuda_device)
    tx._lib.check(lib.texir_texture_mip_build(_ptr(tex), H, W, C, L, _ptr(a), st))
    n0 = tx.launch_count()
    tx._lib.check(lib.texir_texture_mip_build_fast(_ptr(tex), H, W, C, L, _ptr(b), st))
    if H % 32 == 0 and W % 32 == 0:
        assert tx.launch_count() - n0 <= 2
    assert torch.equal(a, b)
    # fold: level by level (in place) against the gather kernel, separate output + clear, and in place
    g0 = torch.randn(H, W, C, generator=g).to(cuda_device)
    gm = torch.randn(n, generator=g).to(cuda_device)
    ref0, refm = g0.clone(), gm.clone()
    tx._lib.check(lib.texir_texture_mip_fold(_ptr(ref0), _ptr(refm), H, W, C, L, st))
    s0, sm = g0.clone(), gm.clone()
    out = torch.full((H, W, C), float("nan"), device=cuda_device)
    n0 = tx.launch_count()
    tx._lib.check(lib.texir_texture_mip_fold_into(_ptr(s0), _ptr(sm), H, W, C, L, _ptr(out), 1, st))
    if W % 4 == 0 and C in (1, 3) and L > 0:
        assert tx.launch_count() - n0 == 1                         # ONE streaming gather pass
    assert torch.equal(out, ref0)
    assert not bool(s0.any()) and (L == 0 or not bool(sm[: int(lib.texir_texture_mip_elems(H, W, C, L))].any()))
    s0, sm = g0.clone(), gm.clone()
    tx._lib.check(lib.texir_texture_mip_fold_into(_ptr(s0), _ptr(sm), H, W, C, L, _ptr(s0), 0, st))
    assert torch.equal(s0, ref0)                                             # in place
    if W % 4 == 0 and C in (1, 3) and L > 0:
        assert torch.equal(sm, gm)                                           # the gather leaves the pyramid as it was


def test_mat_fetch_equals_the_four_fetches(tx, cuda_device):
    rng = np.random.default_rng(5)
    P = 5000
    ta = T(rng.uniform(0, 1, (64, 64, 3)).astype(np.float32), cuda_device)
    tr = T(rng.uniform(0, 1, (128, 128, 1)).astype(np.float32), cuda_device)
    ti = T(rng.uniform(0, 3, (32, 32, 3)).astype(np.float32), cuda_device)
    uv = T(rng.uniform(-0.2, 1.2, (P, 2)).astype(np.float32), cuda_device)
    da = T((rng.normal(size=(P, 4)) * np.exp(rng.uniform(-7, -1, (P, 1)))).astype(np.float32), cuda_device)
    ws = tx.ops.MatWorkspace()
    lib = tx._lib.load()
    (ma, La), (mr, Lr), (mi, Li) = ws.pyramid("a", ta, 13), ws.pyramid("r", tr, 13), ws.pyramid("i", ti, 13)
    n0 = tx.launch_count()
    assert ws.pyramid("a", ta, 13)[0] is ma and tx.launch_count() == n0      # frozen texture: cached while unchanged
    ta.mul_(1.0)
    ma2, _ = ws.pyramid("a", ta, 13)                                          # in-place update -> rebuilt, two launches
    assert tx.launch_count() - n0 == 2
    ta.requires_grad_(True)
    ws.pyramid("a", ta, 13)                                                   # trainable texture: rebuilt every time
    assert tx.launch_count() - n0 == 4
    ta.requires_grad_(False)
    albedo, irr = torch.empty(P, 3, device=cuda_device), torch.empty(P, 3, device=cuda_device)
    rough, rough_w = torch.empty(P, device=cuda_device), torch.empty(P, device=cuda_device)
    n0 = tx.launch_count()
    tx._lib.check(lib.texir_mat_fetch(_ptr(ta), _ptr(ma2), 64, 64, La, _ptr(tr), _ptr(mr), 128, 128, Lr, _ptr(ti), _ptr(mi), 32, 32,
                                      Li, _ptr(uv), _ptr(da), P, _ptr(albedo), _ptr(rough), _ptr(rough_w), _ptr(irr),
                                      _stream(cuda_device)))
    assert tx.launch_count() - n0 == 1
    assert torch.equal(albedo, tx.ops.texture_fetch(ta, uv, da, max_mip_level=13))
    assert torch.equal(rough, tx.ops.texture_fetch(tr, uv, da, max_mip_level=13).reshape(-1))
    assert torch.equal(rough_w, tx.ops.texture_fetch(tr, uv, None, "linear").reshape(-1))
    assert torch.equal(irr, tx.ops.texture_fetch(ti, uv, da, max_mip_level=13))


@pytest.mark.parametrize("mode", ["stash", "recompute"])
@pytest.mark.parametrize("stage", [1, 2])
def test_mat_stage_is_the_composition_of_its_parts(tx, cuda_device, scene_golden, mode, stage):
    """ops.mat_stage (fetches + render + ONE backward kernel + one fold per texture) against
    texture_fetch x4 -> shade -> autograd, with a loss that also touches the fetched materials directly"""
    name, g = scene_golden
    sc = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    sc.set_texture(torch.from_numpy(g["texture"]))
    P = len(g["pos"])
    rng = np.random.default_rng(12)
    ta = T(rng.uniform(0.1, 0.9, (32, 32, 3)).astype(np.float32), cuda_device).requires_grad_(True)
    tr = T(rng.uniform(0.05, 0.7, (64, 64, 1)).astype(np.float32), cuda_device).requires_grad_(True)
    ti = T(rng.uniform(0.5, 3.0, (32, 32, 3)).astype(np.float32), cuda_device)
    texc = T(rng.uniform(0, 1, (P, 2)).astype(np.float32), cuda_device)
    texd = T((rng.normal(size=(P, 4)) * 0.03).astype(np.float32), cuda_device)
    nrm, pos, cam = (T(g[k], cuda_device) for k in ("nrm", "pos", "cam"))
    shift = T(g["render_e14_16_shift"], cuda_device)
    g_rgb = T(g["g_rgb"], cuda_device)
    w_a = T(rng.normal(size=(P, 3)).astype(np.float32), cuda_device)
    w_r = T(rng.normal(size=(P, 1)).astype(np.float32), cuda_device)
    w_rw = T(rng.normal(size=(P, 1)).astype(np.float32), cuda_device)

    def loss(rgb, a, r, rw):
        l = (rgb * g_rgb).sum() + (r * w_r).sum() + (rw * w_rw).sum()
        return l + (a * w_a).sum() if stage == 2 else l

    ta.grad = tr.grad = None
    a = tx.ops.texture_fetch(ta, texc, texd, max_mip_level=13)
    r = tx.ops.texture_fetch(tr, texc, texd, max_mip_level=13)
    rw = tx.ops.texture_fetch(tr, texc, None, "linear")
    irr = tx.ops.texture_fetch(ti, texc, texd, max_mip_level=13)
    rgb = tx.shade(sc, nrm, a.detach() if stage == 1 else a, rw if stage == 1 else r, pos, cam, irr, shift, num_spec=16,
                   backward=mode)
    loss(rgb, a, r, rw).backward()
    ref = (rgb.detach(), None if ta.grad is None else ta.grad.clone(), tr.grad.clone())

    ws = tx.ops.MatWorkspace()
    for rep in range(2):                      # twice: the second step runs on the scatter buffers the first one left
        ta.grad = tr.grad = None
        n0 = tx.launch_count()
        out = tx.ops.mat_stage(sc, ws, ta, tr, ti, nrm, pos, cam, texc, texd, shift, stage=stage, num_spec=16,
                               max_mip_level=13, backward=mode)
        n1 = tx.launch_count()
        nb = lambda side: 1 if tx.ops.mip_levels(side, side, 13) <= 5 else 2
        assert n1 - n0 == 2 + nb(32) + nb(64) + (0 if rep else nb(32))   # fetch + render + the trainable textures' pyramids (+ the IrT's once)
        loss(out[0], out[1], out[2], out[3]).backward()
        assert tx.launch_count() - n1 == (3 if stage == 2 else 2)     # ONE backward kernel + one fold per texture
        assert torch.equal(out[0], ref[0]) and torch.equal(out[1], a.detach()) and torch.equal(out[2], r.detach())
        assert torch.equal(out[3], rw.detach()) and torch.equal(out[4], irr)
        if stage == 2:
            assert rel_l2(ta.grad.cpu().numpy(), ref[1].cpu().numpy()) < 1e-5
        else:
            assert ta.grad is None and ref[1] is None
        assert rel_l2(tr.grad.cpu().numpy(), ref[2].cpu().numpy()) < 1e-5
    if stage == 2:                                        # both gradients are views of ONE buffer (a single all-reduce)
        flat = tx.parallel._flat_view_of([ta.grad, tr.grad])
        assert flat is not None and flat.numel() == ta.numel() + tr.numel()


def test_sharded_gbuffer_and_v_flip(tx, cuda_device, scene_golden):
    from test_host_logic import _reference_cube_mvps
    name, g = scene_golden
    sc = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.2, -0.1, 0.3]
    mvp, cam = _reference_cube_mvps(c2w)
    res = 32                                                       # 1024 pixels per view = 8 blocks of 128
    full = tx.gbuffer_mvp(sc, mvp.to(cuda_device), cam.to(cuda_device), res, return_rays=True)
    flip = tx.gbuffer_mvp(sc, mvp.to(cuda_device), cam.to(cuda_device), res, flip_v=True)
    hit = full["mask"][..., 0] > 0
    assert torch.equal(flip["texc"][..., 0], full["texc"][..., 0])
    assert torch.equal(flip["texc"][..., 1][hit], 1.0 - full["texc"][..., 1][hit])
    assert bool((flip["texc"][..., 1][~hit] == 0).all())
    assert torch.equal(flip["texd"][..., :2], full["texd"][..., :2]) and torch.equal(flip["texd"][..., 2:], -full["texd"][..., 2:])
    for world in (2, 4):
        for rank in range(world):
            part = tx.gbuffer_mvp(sc, mvp.to(cuda_device), cam.to(cuda_device), res, return_rays=True, shard=(rank, world))
            for k in ("position", "normal", "texc", "texd", "mask", "rays", "primitive_uvs"):
                want = tx.parallel.local_pixels(full[k], rank, world, 6)
                assert torch.equal(part[k], want), (k, rank, world)
            assert torch.equal(part["primitive_ids"], tx.parallel.local_pixels(full["primitive_ids"].unsqueeze(-1), rank, world, 6)[:, 0])
    with pytest.raises(ValueError):
        tx.gbuffer_mvp(sc, mvp.to(cuda_device), cam.to(cuda_device), 8, shard=(0, 2))


def test_forward_addresses_the_material_textures_like_the_reference(tx, cuda_device):
    """The reference fetches irrt / materials with pyredner's flipped v (mat_nvdiffrast.py:194,131-139): with an IrT
    that stores (u, v') in its texels, stage 0 must return (u, 1 - v) of the hit point's raw OBJ coordinates"""
    from test_host_logic import _reference_cube_mvps
    from texir_code_b200 import synth
    sc = synth.make_room_scene(3000, seed=3, clutter=3, tex_res=64, uv_kind="planar")
    W = 256
    x = (np.arange(W, dtype=np.float32) + 0.5) / W
    irrt = np.stack(np.broadcast_arrays(x[None, :], x[:, None], np.float32(0.25)), -1).astype(np.float32)  # texel (row, col) = (u, v')
    tex = torch.from_numpy(synth.make_radiance_texture(64, kind="smooth"))
    mm = tx.MaterialModel(sc.vertices, sc.triangles, sc.triangle_uvs, tex, device=cuda_device, materials_a_res=32,
                          materials_r_res=32, cube_res=32, irrt=irrt, max_mip_level=0)
    mm._warned_normals = True
    c2w = np.eye(4, dtype=np.float32)
    mvp, cam = _reference_cube_mvps(c2w)
    out = mm(mvp.to(cuda_device), 0, cam, 0)
    irr = out["rgb"] * np.pi / 0.5                                  # stage 0: rgb = irr * albedo / pi, albedo = 0.5
    raw = tx.gbuffer_mvp(mm.scene, mvp.to(cuda_device), cam.to(cuda_device), 32)
    hit = raw["mask"][..., 0] > 0
    u, v = raw["texc"][..., 0][hit], raw["texc"][..., 1][hit]
    inside = (u > 0.02) & (u < 0.98) & (v > 0.02) & (v < 0.98)      # away from the wrap seam of the bilinear taps
    eu = (irr[..., 0][hit] - u)[inside].abs()
    ev = (irr[..., 1][hit] - (1.0 - v))[inside].abs()
    assert float(eu.max()) < 1e-3 and float(ev.max()) < 1e-3
    # and NOT the un-flipped address (which the round-1 model used): that would be off by |1 - 2v|
    assert float((irr[..., 1][hit] - v)[inside].abs().max()) > 0.1


def test_model_step_leaves_the_cpu_generator_where_the_reference_does(tx, cuda_device):
    """cube_res 128 -> the 98 304-pixel shift comes from the device-side mt19937 with a deferred hand-back of the state:
    after forward() torch's CPU generator must stand where torch.rand(98304, 1, 2) leaves it (utils/sample_util.py:102),
    and the step must have drawn exactly those numbers (same colours as a step fed by the host draw)"""
    import os
    from texir_code_b200 import models, sample_util, synth
    from test_host_logic import _reference_cube_mvps
    sc = synth.make_room_scene(4000, seed=3, tex_res=256, clutter=6)
    mvp, cam = _reference_cube_mvps(np.eye(4, dtype=np.float32))
    mvp, cam = mvp.to(cuda_device), cam.to(cuda_device)
    mm = models.MaterialModel(sc.vertices, sc.triangles, sc.triangle_uvs, torch.from_numpy(synth.make_radiance_texture(256)), sample_l=[32, 16],
                              materials_a_res=128, materials_r_res=256, cube_res=128, irrt=torch.rand(64, 64, 3), device=cuda_device)
    mm._warned_normals = True
    torch.manual_seed(41)
    torch.rand(98304, 1, 2)
    want_state = torch.get_rng_state()
    torch.manual_seed(41)
    a = mm(mvp, 0, cam, 2)["rgb"]
    assert sample_util._mt_pending is None and sample_util._mt_ok is True
    assert torch.equal(torch.get_rng_state(), want_state)
    os.environ["TEXIR_SHIFT_DEVICE"] = "0"
    try:
        torch.manual_seed(41)
        b = mm(mvp, 0, cam, 2)["rgb"]
    finally:
        del os.environ["TEXIR_SHIFT_DEVICE"]
    assert torch.equal(torch.get_rng_state(), want_state) and torch.equal(a, b)


def test_sharded_mat_step_sums_to_the_one_gpu_step(tx, cuda_device, scene_golden):
    """two shards run one after the other on one GPU: their colours are the one-GPU colours of their pixels and their
    texture gradients add up to the one-GPU gradient (what the single all-reduce computes)"""
    from test_host_logic import _reference_cube_mvps
    name, g = scene_golden
    res = 32
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.2, -0.1, 0.3]
    mvp, cam = _reference_cube_mvps(c2w)
    rng = np.random.default_rng(3)
    irrt = rng.uniform(0.5, 3.0, (32, 32, 3)).astype(np.float32)
    scene = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    a0 = T(rng.uniform(0.1, 0.9, (64, 64, 3)).astype(np.float32), cuda_device)
    r0 = T(rng.uniform(0.05, 0.7, (64, 64, 1)).astype(np.float32), cuda_device)
    g_full = T(rng.normal(size=(6, res, res, 3)).astype(np.float32), cuda_device)

    def run(shard):
        mm = tx.MaterialModel(None, None, g["triangle_uvs"], torch.from_numpy(g["texture"]), sample_l=[16, 16], scene=scene,
                              materials_a_res=64, materials_r_res=64, cube_res=res, irrt=irrt, max_mip_level=13, shard=shard)
        mm._warned_normals = True
        with torch.no_grad():
            mm.materials_a.copy_(a0)
            mm.materials_r.copy_(r0)
        torch.manual_seed(666)
        out = mm(mvp.to(cuda_device), 0, cam, 2)
        gl = g_full if shard is None else tx.parallel.local_pixels(g_full, shard[0], shard[1], 6)
        (out["rgb"] * gl.reshape(out["rgb"].shape)).sum().backward()
        return out["rgb"].detach(), mm.materials_a.grad.clone(), mm.materials_r.grad.clone()

    full = run(None)
    parts = [run((r, 2)) for r in range(2)]
    for r in range(2):
        assert torch.equal(parts[r][0], tx.parallel.local_pixels(full[0], r, 2, 6))
    assert rel_l2((parts[0][1] + parts[1][1]).cpu().numpy(), full[1].cpu().numpy()) < 1e-5
    assert rel_l2((parts[0][2] + parts[1][2]).cpu().numpy(), full[2].cpu().numpy()) < 1e-5


_PEER_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
rank, world, port = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
dist.init_process_group("gloo", rank=rank, world_size=world)      # both ranks share cuda:0; NCCL would refuse that
from texir_code_b200 import models, parallel, synth
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sc = synth.make_room_scene(4000, seed=666, clutter=6, tex_res=128, uv_kind="planar")
tex = torch.from_numpy(synth.make_radiance_texture(128, kind="smooth"))
P, N = 4096, 64
pos, nrm = synth.sample_surface_points(sc, P, seed=3)
torch.manual_seed(666)
shift = torch.rand(P, 1, 2).to(dev)
tr = models.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, tex, sample_l=[N, 16], device=dev)
p, n = torch.from_numpy(pos).to(dev), torch.from_numpy(nrm).to(dev)
want = tr.irradiance(p, n, shift)
res = {}
for coll in ("peer", "nccl"):
    job = parallel.ShardedIrradiance(tr, P, collective=coll)
    for it in range(3):
        got = job(p, n, shift).clone()
    torch.cuda.synchronize()
    res[coll] = bool(torch.equal(got, want))
    if coll == "peer":
        res["timed_out"] = job.peers.timed_out()
    job.close()
print("RESULT", rank, res["peer"], res["nccl"], res["timed_out"], flush=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_ranks_sharded_irt_with_peer_stores(tx, cuda_device, tmp_path):
    """ShardedIrradiance on two processes (both on cuda:0, gloo for the rendezvous): the kernel of each rank stores its
    texels into BOTH ranks' buffers through CUDA IPC mappings, a flag barrier orders them; every rank ends up with the
    one-GPU result bit for bit.  The library formulation (all-reduce of a zeroed buffer) gives the same."""
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        line = [l for l in o.splitlines() if l.startswith("RESULT")][-1].split()
        assert line[1:] == [str(r), "True", "True", "False"], o[-3000:]


_MAT_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
rank, world, port = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
dist.init_process_group("gloo", rank=rank, world_size=world)
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from test_host_logic import _reference_cube_mvps
from texir_code_b200 import models, parallel, synth, Scene
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sc = synth.make_room_scene(4000, seed=666, clutter=6, tex_res=128, uv_kind="planar")
tex = torch.from_numpy(synth.make_radiance_texture(128, kind="smooth"))
scene = Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev)
c2w = np.eye(4, dtype=np.float32); c2w[:3, 3] = [0.2, -0.1, 0.3]
mvp, cam = _reference_cube_mvps(c2w)
res = 32
rng = np.random.default_rng(3)
irrt = rng.uniform(0.5, 3.0, (32, 32, 3)).astype(np.float32)
a0 = torch.from_numpy(rng.uniform(0.1, 0.9, (64, 64, 3)).astype(np.float32)).to(dev)
r0 = torch.from_numpy(rng.uniform(0.05, 0.7, (64, 64, 1)).astype(np.float32)).to(dev)
g_full = torch.from_numpy(rng.normal(size=(6, res, res, 3)).astype(np.float32)).to(dev)
w_r = torch.from_numpy(rng.normal(size=(6, res, res, 1)).astype(np.float32)).to(dev)
def run(shard, sync):
    mm = models.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[16, 16], scene=scene, materials_a_res=64,
                              materials_r_res=64, cube_res=res, irrt=irrt, max_mip_level=13, shard=shard, grad_sync=sync)
    mm._warned_normals = True
    with torch.no_grad():
        mm.materials_a.copy_(a0); mm.materials_r.copy_(r0)
    torch.manual_seed(666)
    out = mm(mvp.to(dev), 0, cam, 2)
    loc = (lambda t: t) if shard is None else (lambda t: parallel.local_pixels(t, shard[0], shard[1], 6))
    loss = (out["rgb"] * loc(g_full).reshape(out["rgb"].shape)).sum() + (out["roughness"] * loc(w_r).reshape(out["roughness"].shape)).sum()
    loss.backward()
    if shard is not None and sync == "allreduce":
        parallel.allreduce_gradients([mm.materials_a, mm.materials_r])
    return mm.materials_a.grad.clone(), mm.materials_r.grad.clone()
full = run(None, "allreduce")
ar = run((rank, world), "allreduce")
ga = run((rank, world), "gather")
rel = lambda x, y: float(torch.linalg.norm(x - y) / torch.linalg.norm(y))
print("RESULT", rank, rel(ar[0], full[0]) < 1e-5, rel(ar[1], full[1]) < 1e-5, rel(ga[0], full[0]) < 1e-5, rel(ga[1], full[1]) < 1e-5, flush=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_ranks_sharded_mat_step_allreduce_and_gather(tx, cuda_device, tmp_path):
    """the sharded Mat-stage step on two processes (both on cuda:0, gloo): with either synchronisation -- ONE all-reduce of
    the flat texture-gradient buffer, or one all-gather of the per-pixel gradient rows followed by the same scatter on every
    rank -- every rank ends up with the one-GPU texture gradients (render terms and the loss's direct roughness term)"""
    script = tmp_path / "mat_worker.py"
    script.write_text(_MAT_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        line = [l for l in o.splitlines() if l.startswith("RESULT")][-1].split()
        assert line[1:] == [str(r), "True", "True", "True", "True"], o[-3000:]


def test_graphed_step_is_the_eager_step(tx, cuda_device):
    """graphed.GraphedMatStep: zero_grad + forward + loss + backward captured once, replayed with new inputs.  Against the
    same steps run eagerly from the same seed: the loss bit for bit (the forward is deterministic), the gradients to the
    rounding of their atomics, and torch's CPU generator in the identical state after every step (the shift stream
    travels through the graph: sample_util.GraphRng)"""
    from test_host_logic import _reference_cube_mvps
    from texir_code_b200 import graphed, models, sample_util, synth
    sc = synth.make_room_scene(4000, seed=3, tex_res=256, clutter=6)
    tex = torch.from_numpy(synth.make_radiance_texture(256))
    rng = np.random.default_rng(0)
    views = []
    for k in range(4):
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, 3] = rng.uniform(-0.5, 0.5, 3)
        mvp, cam = _reference_cube_mvps(c2w)
        views.append((mvp.to(cuda_device), cam.to(cuda_device), torch.rand(6, 64, 64, 3, device=cuda_device)))

    def make():
        torch.manual_seed(0)
        mm = models.MaterialModel(sc.vertices, sc.triangles, sc.triangle_uvs, tex, sample_l=[32, 16], materials_a_res=128,
                                  materials_r_res=256, cube_res=64, irrt=torch.rand(64, 64, 3) + 0.5, device=cuda_device)
        mm._warned_normals = True
        return mm

    loss_fn = lambda out, gt: ((out["rgb"] - gt) ** 2).mean() + 0.1 * out["roughness"].mean()
    for stage in (2, 1, 0):
        eager = make()
        torch.manual_seed(77)
        ref = []
        for mvp, cam, gt in views:
            eager.zero_grad(set_to_none=True)
            loss = loss_fn(eager(mvp, 0, cam, stage), gt)
            loss.backward()
            ref.append((loss.detach().clone(), None if eager.materials_a.grad is None else eager.materials_a.grad.clone(),
                        eager.materials_r.grad.clone(), torch.get_rng_state()))
        mm = make()
        # a living autograd graph from an eager call on the default stream pins the parameters' AccumulateGrad nodes to
        # that stream; the capture must not depend on them (it enters the textures as fresh leaves)
        keep = mm(views[0][0], 0, views[0][1], stage)
        step = graphed.GraphedMatStep(mm, stage, loss_fn, example=views[0])
        assert keep["rgb"].grad_fn is not None
        n0 = tx.launch_count()
        torch.manual_seed(77)
        for (mvp, cam, gt), (loss_w, ga_w, gr_w, state_w) in zip(views, ref):
            loss = step(mvp, cam, gt)
            assert torch.equal(loss, loss_w)
            sample_util.finish_pending()
            assert torch.equal(torch.get_rng_state(), state_w)
            if ga_w is None:
                assert mm.materials_a.grad is None or float(mm.materials_a.grad.abs().max()) == 0.0
            else:
                assert torch.allclose(mm.materials_a.grad, ga_w, rtol=1e-4, atol=1e-9)
            assert torch.allclose(mm.materials_r.grad, gr_w, rtol=1e-4, atol=1e-9)
            assert float(gr_w.abs().max()) > 0
        assert tx.launch_count() == n0          # replays launch nothing from Python: the graph holds the kernels
