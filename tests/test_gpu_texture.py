"""GPU (-m gpu): mip-mapped material-texture fetch and its texture-space gradient, and the UV
derivatives of the G-buffer that drive it, through the package's binding of the C ABI.  Reference: the oracle's torch
restatement of dr.texture (gradients from torch autograd) and its float64 ray-differential formula.
Tolerances: forward 1e-5 rel-L2 (fp32 taps and blends), gradients 1e-4 (float atomics in arbitrary order)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.fixture(scope="module")
def tx(cuda_device):
    import texir_code_b200 as tx
    tx._lib.load()
    return tx


def _case(H, W, C, P, seed):
    rng = np.random.default_rng(seed)
    tex = rng.uniform(0, 1, (H, W, C)).astype(np.float32)
    uv = rng.uniform(-0.3, 1.3, (P, 2)).astype(np.float32)
    scale = np.exp(rng.uniform(np.log(0.05), np.log(40.0), (P, 1))) / max(H, W)
    uv_da = (rng.normal(size=(P, 4)) * scale).astype(np.float32)
    uv_da[:8] = 0.0
    g_out = rng.normal(size=(P, C)).astype(np.float32)
    return tex, uv, uv_da, g_out


@pytest.mark.parametrize("shape,max_level", [((64, 64, 3), None), ((32, 128, 1), None), ((256, 256, 3), 13),
                                             ((64, 64, 3), 2), ((12, 8, 2), 2), ((1, 16, 3), None)])
def test_mip_fetch_fwd_bwd_vs_oracle(tx, oracle, cuda_device, shape, max_level):
    tex, uv, uv_da, g_out = _case(*shape, 20000, 3)
    t = T(tex, cuda_device).requires_grad_(True)
    out = tx.ops.texture_fetch(t, T(uv, cuda_device), T(uv_da, cuda_device), "linear-mipmap-linear", max_level)
    (out * T(g_out, cuda_device)).sum().backward()
    t_ref = torch.from_numpy(tex).requires_grad_(True)
    ref = oracle.texture_mipmap_linear(t_ref, torch.from_numpy(uv), torch.from_numpy(uv_da), max_level)
    (ref * torch.from_numpy(g_out)).sum().backward()
    assert rel_l2(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-5
    assert rel_l2(t.grad.cpu().numpy(), t_ref.grad.numpy()) < 1e-4
    # 'auto' = mip-mapped when derivatives are given, bilinear otherwise (dr.texture's default)
    auto = tx.ops.texture_fetch(t, T(uv, cuda_device), T(uv_da, cuda_device), max_mip_level=max_level)
    assert torch.equal(auto, out)
    lin = tx.ops.texture_fetch(t, T(uv, cuda_device))
    assert rel_l2(lin.detach().cpu().numpy(), oracle.texture_linear(torch.from_numpy(tex), torch.from_numpy(uv)).numpy()) < 1e-5


def test_mip_pyramid_and_fold_are_transposes(tx, oracle, cuda_device):
    """<build(x), y> == <x, fold(y)>: the gradient fold is the exact transpose of the box-filter pyramid"""
    rng = np.random.default_rng(5)
    H, W, C = 64, 32, 3
    x = rng.normal(size=(H, W, C)).astype(np.float32)
    mips, L = tx.ops.build_mips(T(x, cuda_device))
    levels = oracle.mip_pyramid(torch.from_numpy(x))
    assert L == len(levels) - 1 == 6
    ref = torch.cat([l.reshape(-1) for l in levels[1:]]).numpy()
    assert rel_l2(mips.cpu().numpy(), ref) < 1e-6
    y = rng.normal(size=mips.shape).astype(np.float32)
    g_tex = torch.zeros(H, W, C, device=cuda_device)
    g_mips = T(y, cuda_device).clone()
    lib = tx._lib.load()
    tx._lib.check(lib.texir_texture_mip_fold(g_tex.data_ptr(), g_mips.data_ptr(), H, W, C, L, None))
    torch.cuda.synchronize()
    lhs = float((mips.double().cpu() * torch.from_numpy(y).double()).sum())
    rhs = float((torch.from_numpy(x).double() * g_tex.double().cpu()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_mip_fetch_at_material_texture_sizes(tx, cuda_device):
    """materials_a 2048^2 x 3 and materials_r 4096^2 x 1 with max_mip_level = get_mip_level(8192) = 13
    (mat_nvdiffrast.py:65-69): level counts, limits, and gradient mass conservation at full size"""
    P = 6 * 128 * 128
    rng = np.random.default_rng(9)
    uv = T(rng.uniform(0, 1, (P, 2)).astype(np.float32), cuda_device)
    for res, C, L in ((2048, 3, 11), (4096, 1, 12)):
        assert tx.ops.mip_levels(res, res, 13) == L
        tex = torch.rand(res, res, C, device=cuda_device).requires_grad_(True)
        zero = torch.zeros(P, 4, device=cuda_device)
        a = tx.ops.texture_fetch(tex, uv, zero, "linear-mipmap-linear", 13)
        assert torch.allclose(a, tx.ops.texture_fetch(tex, uv), rtol=0, atol=1e-6)      # level 0 == bilinear
        huge = torch.full((P, 4), 4.0, device=cuda_device)
        b = tx.ops.texture_fetch(tex, uv, huge, "linear-mipmap-linear", 13)
        assert torch.allclose(b, tex.detach().reshape(-1, C).mean(0).expand(P, C), rtol=1e-4, atol=1e-5)
        da = T((rng.normal(size=(P, 4)) * np.exp(rng.uniform(-9, -3, (P, 1)))).astype(np.float32), cuda_device)
        out = tx.ops.texture_fetch(tex, uv, da, "linear-mipmap-linear", 13)
        g = torch.rand(P, C, device=cuda_device)
        (out * g).sum().backward()
        # every tap's weights sum to 1 on every level and the fold preserves mass
        assert abs(float(tex.grad.double().sum()) - float(g.double().sum())) < 1e-3 * float(g.double().sum())


def test_gbuffer_uv_derivatives(tx, oracle, cuda_device, scene_golden):
    from texir_code_b200 import synth
    name, g = scene_golden
    sc = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    res = 32
    rays = synth.cube_map_rays(res, cam=(0.2, -0.1, 0.3))
    dd = synth.cube_map_ray_differentials(res)
    gb = tx.gbuffer(sc, T(rays, cuda_device).reshape(6, res, res, 6), ray_differentials=T(dd, cuda_device))
    assert gb["texd"].shape == (6, res, res, 4)
    prim = gb["primitive_ids"].cpu().numpy().reshape(-1).view(np.uint32)
    ref = oracle.gbuffer_uv_da(g["vertices"], g["triangles"], g["triangle_uvs"], rays, dd, prim)
    got = gb["texd"].cpu().numpy().reshape(-1, 4)
    hit = prim != 0xFFFFFFFF
    assert np.all(got[~hit] == 0)
    assert rel_l2(got[hit], ref[hit]) < 1e-4
    # independent check: central differences of the G-buffer's own texc along the ray differentials, on the pixels
    # whose displaced rays stay on the same triangle
    h = 0.05
    for k in range(2):
        step = dd[:, 3 * k:3 * k + 3] * h
        plus, minus = rays.copy(), rays.copy()
        plus[:, 3:] += step
        minus[:, 3:] -= step
        gp = tx.gbuffer(sc, T(plus, cuda_device))
        gm = tx.gbuffer(sc, T(minus, cuda_device))
        same = hit & (gp["primitive_ids"].cpu().numpy().view(np.uint32) == prim) & \
            (gm["primitive_ids"].cpu().numpy().view(np.uint32) == prim)
        assert same.mean() > 0.3
        fd = (gp["texc"].double() - gm["texc"].double()).cpu().numpy() / (2 * h)
        an = got[:, [k, 2 + k]]
        assert rel_l2(an[same], fd[same]) < 5e-3


def test_mip_fetch_argument_errors(tx, cuda_device):
    tex = torch.rand(6, 4, 3, device=cuda_device)
    uv = torch.rand(10, 2, device=cuda_device)
    with pytest.raises(ValueError):
        tx.ops.texture_fetch(tex, uv, torch.zeros(10, 4, device=cuda_device), "linear-mipmap-linear")   # 6 -> 3 -> odd
    with pytest.raises(ValueError):
        tx.ops.texture_fetch(tex, uv, None, "linear-mipmap-linear")
    with pytest.raises(ValueError):
        tx.ops.texture_fetch(tex, uv, None, "nearest")
    out = tx.ops.texture_fetch(tex, uv, torch.zeros(10, 4, device=cuda_device), "linear-mipmap-linear", 1)
    assert out.shape == (10, 3)


def test_material_model_forward_reference_signature(tx, oracle, cuda_device, scene_golden):
    """MaterialModel.forward(mvp, id, cam_position, stage) as in mat_nvdiffrast.py:105-191: primary rays from the
    reference's own cube-face mvps -> G-buffer -> mip-mapped fetches -> fused render.  Checked: the G-buffer ids
    against the brute-force oracle, the result keys / shapes, which textures receive gradients per stage, the
    stage -1 texture swap, and stage 2 against the same chain assembled op by op on the oracle (CPU, torch autograd)."""
    from test_host_logic import _reference_cube_mvps
    name, g = scene_golden
    res = 8
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.2, -0.1, 0.3]
    mvp, cam = _reference_cube_mvps(c2w)
    rng = np.random.default_rng(3)
    irrt = rng.uniform(0.5, 3.0, (32, 32, 3)).astype(np.float32)
    mm = tx.MaterialModel(g["vertices"], g["triangles"], g["triangle_uvs"], torch.from_numpy(g["texture"]),
                          sample_l=[16, 16], device=cuda_device, materials_a_res=32, materials_r_res=64, cube_res=res,
                          irrt=irrt, max_mip_level=13, hdr_exposure=5.0)
    with torch.no_grad():
        mm.materials_a.copy_(T(rng.uniform(0.1, 0.9, (32, 32, 3)).astype(np.float32), cuda_device))
        mm.materials_r.copy_(T(rng.uniform(0.05, 0.7, (64, 64, 1)).astype(np.float32), cuda_device))
    gb = mm.gbuffers(mvp, cam)
    rays, dd = tx.models.rays_from_mvp(mvp, cam, res)
    t0, i0, u0 = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays.reshape(-1, 6).numpy())
    assert np.array_equal(gb["primitive_ids"].cpu().numpy().reshape(-1).view(np.uint32), i0)
    keys = {"rgb", "albedo", "normal", "position", "empty_mask", "roughness_womipmap", "roughness"}
    grads = {}
    for stage in (-1, 0, 1, 2):
        mm.zero_grad()
        tex_before = mm.texture
        torch.manual_seed(666)
        out = mm(mvp.to(cuda_device), 0, cam, stage)
        assert mm.texture is tex_before                                     # stage -1 swaps and restores
        assert set(out) == keys
        assert out["rgb"].shape == (6, res, res, 3) and out["empty_mask"].shape == (6, res, res, 1)
        assert out["roughness"].shape == (6, res, res, 1) and bool(torch.isfinite(out["rgb"]).all())
        if stage >= 0:
            out["rgb"].sum().backward()
        grads[stage] = (mm.materials_a.grad is not None and bool((mm.materials_a.grad != 0).any()),
                        mm.materials_r.grad is not None and bool((mm.materials_r.grad != 0).any()))
    assert grads[0] == (True, False)       # lambertian initialisation: albedo only
    assert grads[1] == (False, True)       # albedo detached, un-mip-mapped roughness
    assert grads[2] == (True, True)
    # stage 2 on the oracle: same G-buffer, torch restatement of dr.texture, reference render restatement -- executed
    # like the reference executes it: everything on the CUDA device except query_irf (host), mat_nvdiffrast.py:294
    P = 6 * res * res
    dv = cuda_device
    texc, texd = gb["texc"].reshape(P, 2).cpu(), gb["texd"].reshape(P, 4).cpu()
    ta = mm.materials_a.detach().cpu().clone().requires_grad_(True)
    tr = mm.materials_r.detach().cpu().clone().requires_grad_(True)
    a = oracle.texture_mipmap_linear(ta, texc, texd, 13).to(dv)          # the fetch restatement runs on the host ...
    r = oracle.texture_mipmap_linear(tr, texc, texd, 13).to(dv)
    irr = oracle.texture_mipmap_linear(torch.from_numpy(irrt), texc, texd, 13).to(dv)
    nrm, pos = gb["normal"].reshape(P, 3), gb["position"].reshape(P, 3)  # ... the render where the reference runs it
    torch.manual_seed(666)
    shift = torch.rand(P, 1, 2).to(dv)
    ref = oracle.render(oracle.OracleScene(g["vertices"], g["triangles"]), g["triangle_uvs"], torch.from_numpy(g["texture"]),
                        [16, 16], ["uniform", "importance"], nrm, a, r, pos + 1e-2 * nrm, cam.to(dv), irr, shift_spec=shift)
    g_rgb = torch.from_numpy(rng.normal(size=(P, 3)).astype(np.float32)).to(dv)
    (ref["rgb"] * g_rgb).sum().backward()
    mm.zero_grad()
    torch.manual_seed(666)
    out = mm(mvp.to(cuda_device), 0, cam, 2)
    (out["rgb"].reshape(P, 3) * g_rgb).sum().backward()
    assert rel_l2(out["rgb"].detach().cpu().numpy().reshape(P, 3), ref["rgb"].detach().cpu().numpy()) < 1e-4
    assert rel_l2(mm.materials_a.grad.cpu().numpy(), ta.grad.cpu().numpy()) < 1e-3
    assert rel_l2(mm.materials_r.grad.cpu().numpy(), tr.grad.cpu().numpy()) < 1e-3


def test_gbuffer_mvp_one_launch_equals_rays_plus_gbuffer(tx, cuda_device, scene_golden):
    """ops.gbuffer_mvp (camera matrices -> rays -> BVH -> G-buffer + texd in one kernel) against the two-step path
    rays_from_mvp (torch, float64) -> ops.gbuffer(..., ray_differentials)"""
    from test_host_logic import _reference_cube_mvps
    name, g = scene_golden
    sc = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    rng = np.random.default_rng(8)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, :3] = q
    c2w[:3, 3] = [0.3, 0.1, -0.4]
    mvp, cam = _reference_cube_mvps(c2w)
    res = 24
    n0 = tx.launch_count()
    a = tx.gbuffer_mvp(sc, mvp.to(cuda_device), cam.to(cuda_device), res, return_rays=True)
    assert tx.launch_count() - n0 == 1
    rays, dd = tx.models.rays_from_mvp(mvp, cam, res)
    assert float((a["rays"].cpu() - rays).abs().max()) < 2e-7
    b = tx.gbuffer(sc, a["rays"], ray_differentials=dd.to(cuda_device))      # same rays: everything but texd is identical
    for k in ("position", "normal", "texc", "primitive_ids", "primitive_uvs", "mask"):
        assert torch.equal(a[k], b[k]), k
    hit = a["mask"].reshape(-1) > 0
    assert float(hit.float().mean()) > 0.9
    assert rel_l2(a["texd"].reshape(-1, 4)[hit].cpu().numpy(), b["texd"].reshape(-1, 4)[hit].cpu().numpy()) < 1e-5
    assert bool((a["texd"].reshape(-1, 4)[~hit] == 0).all())
    bad = torch.zeros(1, 4, 4, device=cuda_device)                            # singular mvp: no hits, no NaNs
    z = tx.gbuffer_mvp(sc, bad, cam.to(cuda_device), 4)
    assert bool((z["mask"] == 0).all()) and bool(torch.isfinite(z["position"]).all())


@pytest.mark.parametrize("mode", ["stash", "recompute"])
def test_shade_textured_backward_lands_in_texture_space(tx, cuda_device, scene_golden, mode):
    """ops.shade_textured (ONE backward kernel: render backward + atomic scatter into the material textures' gradients,
    then the pyramid folds) against the two-step path texture_fetch -> shade -> autograd, for the mip-mapped filter
    (stage 2), the bilinear roughness-only routing (stage 1), and a mixed call"""
    name, g = scene_golden
    sc = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    sc.set_texture(torch.from_numpy(g["texture"]))
    P = len(g["pos"])
    rng = np.random.default_rng(12)
    ta = T(rng.uniform(0.1, 0.9, (32, 32, 3)).astype(np.float32), cuda_device).requires_grad_(True)
    tr = T(rng.uniform(0.05, 0.7, (64, 64, 1)).astype(np.float32), cuda_device).requires_grad_(True)
    texc = T(rng.uniform(0, 1, (P, 2)).astype(np.float32), cuda_device)
    texd = T((rng.normal(size=(P, 4)) * 0.03).astype(np.float32), cuda_device)
    nrm, pos, cam, irr = (T(g[k], cuda_device) for k in ("nrm", "pos", "cam", "irr"))
    shift = T(g["render_e14_16_shift"], cuda_device)
    g_rgb = T(g["g_rgb"], cuda_device)

    def two_step(da_a, da_r, use_a):
        ta.grad = tr.grad = None
        a = tx.ops.texture_fetch(ta, texc, da_a, max_mip_level=13)
        r = tx.ops.texture_fetch(tr, texc, da_r, max_mip_level=13)
        rgb = tx.shade(sc, nrm, a if use_a else a.detach(), r, pos, cam, irr, shift, num_spec=16, backward=mode)
        rgb.backward(g_rgb)
        return rgb.detach(), None if ta.grad is None else ta.grad.clone(), tr.grad.clone()

    def fused(da, use_a, albedo_mipmap=True, rough_mipmap=True):
        ta.grad = tr.grad = None
        with torch.no_grad():
            a = tx.ops.texture_fetch(ta, texc, da if albedo_mipmap else None, max_mip_level=13)
            r = tx.ops.texture_fetch(tr, texc, da if rough_mipmap else None, max_mip_level=13)
        n0 = tx.launch_count()
        rgb = tx.ops.shade_textured(sc, ta if use_a else None, tr, nrm, a, r, pos, cam, irr, shift, texc, da, num_spec=16,
                                    albedo_mipmap=albedo_mipmap, rough_mipmap=rough_mipmap, max_mip_level=13, backward=mode)
        n1 = tx.launch_count()
        rgb.backward(g_rgb)
        return rgb.detach(), None if ta.grad is None else ta.grad.clone(), tr.grad.clone(), tx.launch_count() - n1, n1 - n0

    # stage 2: both textures mip-mapped
    ref = two_step(texd, texd, True)
    got = fused(texd, True)
    assert torch.equal(got[0], ref[0])
    assert rel_l2(got[1].cpu().numpy(), ref[1].cpu().numpy()) < 1e-5 and rel_l2(got[2].cpu().numpy(), ref[2].cpu().numpy()) < 1e-5
    assert got[4] == 1                                            # forward: the render kernel alone
    assert got[3] == 1 + tx.ops.mip_levels(32, 32, 13) + tx.ops.mip_levels(64, 64, 13)   # backward: one kernel + the folds
    # stage 1: albedo detached, roughness fetched bilinearly
    ref = two_step(None, None, False)
    got = fused(None, False)
    assert got[1] is None and ref[1] is None
    assert rel_l2(got[2].cpu().numpy(), ref[2].cpu().numpy()) < 1e-5 and got[3] == 1
    # mixed: mip-mapped albedo next to a bilinear roughness, one call
    ref = two_step(texd, None, True)
    got = fused(texd, True, rough_mipmap=False)
    assert rel_l2(got[1].cpu().numpy(), ref[1].cpu().numpy()) < 1e-5 and rel_l2(got[2].cpu().numpy(), ref[2].cpu().numpy()) < 1e-5
