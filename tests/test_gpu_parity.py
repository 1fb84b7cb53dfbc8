"""GPU (-m gpu): the CUDA path, called through the package's ctypes binding of the C ABI, against
(1) the golden vectors produced by the reference's own Python, (2) the CPU oracle on the same seeded inputs,
(3) size-independent properties at the full 1M-triangle size.

Tolerances (BASELINE.json north_star): forward rgb / irradiance rel-L2 <= 1e-4, gradients rel-L2 <= 1e-3;
ray/triangle hits (ids, t, barycentrics) and texture lookups are bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 1e-3


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.fixture(scope="module")
def tx(cuda_device):
    import texir_code_b200 as tx
    tx._lib.load()
    return tx


@pytest.fixture(scope="module")
def gscene(tx, cuda_device, scene_golden):
    name, g = scene_golden
    sc = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    sc.set_texture(torch.from_numpy(g["texture"]))
    return name, g, sc


def _rays(R, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform([-3.5, -1.2, -2.5], [3.5, 1.2, 2.5], size=(R, 3)).astype(np.float32)
    d = rng.normal(size=(R, 3)).astype(np.float32)
    d[:4] = 0.0
    d[4:8] = [[1, 0, 0], [0, -1, 0], [0, 0, 1], [-1, 0, 0]]
    return np.concatenate([o, d], 1)


# ---------------------------------------------------------------------------------------------------
# cast_rays
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [2, 8])
def test_cast_rays_bit_exact_vs_oracle(tx, oracle, cuda_device, kind):
    g = load_golden("scene_atlas")
    rays = _rays(20000, 5)
    t0, i0, u0 = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays)
    sc = tx.Scene(g["vertices"], g["triangles"], device=cuda_device, bvh=kind)
    n0 = tx.launch_count()
    res = sc.cast_rays(T(rays, cuda_device))
    assert tx.launch_count() > n0  # the native library ran
    assert np.array_equal(res["primitive_ids"].cpu().numpy().view(np.uint32), i0)
    assert np.array_equal(res["t_hit"].cpu().numpy(), t0)
    assert np.array_equal(res["primitive_uvs"].cpu().numpy(), u0)
    st = sc.cast_rays(T(rays, cuda_device), stats=True)
    assert np.array_equal(st["t_hit"].cpu().numpy(), t0)
    assert 1 < st["nodes_visited"].float().mean().item() < 80


def test_raycasting_scene_is_an_open3d_drop_in(tx, oracle, cuda_device):
    g = load_golden("scene_atlas")
    rays = _rays(3000, 9).reshape(30, 100, 6)
    rs = tx.RaycastingScene(device=cuda_device)
    assert rs.add_triangles(g["vertices"], g["triangles"]) == 0
    out = rs.cast_rays(rays)  # numpy in, objects with .numpy() out
    t0, i0, u0 = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays)
    assert out["t_hit"].numpy().shape == (30, 100) and out["primitive_ids"].numpy().dtype == np.uint32
    assert np.array_equal(out["primitive_ids"].numpy(), i0) and np.array_equal(out["t_hit"].numpy(), t0)
    assert np.array_equal(out["primitive_uvs"].numpy(), u0)
    ids = out["primitive_ids"].numpy()
    ids[ids == 0xFFFFFFFF] = 0  # the reference writes into the returned array (mat_nvdiffrast.py:305)
    nrm = out["primitive_normals"].numpy()
    hit = np.isfinite(t0)
    assert np.allclose(np.linalg.norm(nrm[hit], axis=-1), 1, atol=1e-5) and np.all(nrm[~hit] == 0)
    assert np.all(out["geometry_ids"].numpy()[hit] == 0)
    # the REAL reference query_irf logic on top of it == our fused query_irf
    rad_ref = oracle.query_irf(rs, g["triangle_uvs"], torch.from_numpy(g["texture"]), torch.from_numpy(g["pos"]).unsqueeze(1).expand(96, 16, 3),
                               torch.from_numpy(g["render_e14_16_l"]).unsqueeze(-2), 16)
    assert np.array_equal(rad_ref.numpy(), g["render_e14_16_lighting"])


# ---------------------------------------------------------------------------------------------------
# generate_dir
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [16, 64])
@pytest.mark.parametrize("mode", ["uniform", "cosine", "importance"])
def test_generate_dir_vs_reference(tx, cuda_device, N, mode):
    g = load_golden("generate_dir")
    nrm = T(g["normals"], cuda_device)
    rough = T(g["roughness"], cuda_device).requires_grad_(True)
    torch.manual_seed(666)  # the op draws torch.rand(b,1,2) on the CPU generator like the reference
    L = tx.generate_dir(nrm, N, None, mode, rough if mode == "importance" else None)
    want = g["L_%s_%d" % (mode, N)]
    assert L.shape == want.shape and L.is_cuda
    assert np.abs(L.detach().cpu().numpy() - want).max() <= (2e-6 if mode == "importance" else 4e-7)
    assert rel_l2(L.detach().cpu().numpy(), want) < 1e-6
    assert torch.all(L[3] == 0)  # zero normal
    if mode == "importance":
        (L * T(g["gL_%d" % N], cuda_device)).sum().backward()
        assert rel_l2(rough.grad.cpu().numpy(), g["grough_%d" % N]) < GRAD_TOL


def test_generate_dir_edge_shapes(tx, cuda_device):
    nrm = torch.tensor([[0.0, 1.0, 0.0]], device=cuda_device)
    for N in (1, 3, 5, 2048):
        L = tx.generate_dir(nrm, N, None, "cosine", shift=torch.zeros(1, 1, 2))
        assert L.shape == (1, N, 3) and torch.isfinite(L).all()
        assert (L[..., 1] > -1e-6).all()  # upper hemisphere
    assert tx.generate_dir(torch.zeros(0, 3, device=cuda_device), 8).shape == (0, 8, 3)
    Li = tx.generate_dir(nrm.expand(7, 3).contiguous(), 32, None, "uniform", pre_mode="independent")
    assert Li.shape == (7, 32, 3) and (Li[..., 1] > -1e-6).all()
    with pytest.raises(ValueError):
        tx.generate_dir(nrm, 4, None, "importance")
    with pytest.raises(ValueError):
        tx.generate_dir(nrm, 4, None, "bogus")


# ---------------------------------------------------------------------------------------------------
# query_irf
# ---------------------------------------------------------------------------------------------------
def test_query_irf_quad_known_answers(tx, cuda_device):
    q = load_golden("query_irf_quad")
    for kind in (2, 8):
        sc = tx.Scene(q["vertices"], q["triangles"], q["triangle_uvs"], device=cuda_device, bvh=kind)
        sc.set_texture(torch.from_numpy(q["texture"]))
        rad = tx.query_irf(sc, T(q["org"], cuda_device).unsqueeze(1), T(q["dir"], cuda_device).reshape(6, 1, 1, 3), 1)
        assert np.array_equal(rad.cpu().numpy(), q["rad"])


@pytest.mark.parametrize("kind", [2, 8])
def test_query_irf_bit_exact_vs_reference(tx, gscene, cuda_device, kind):
    name, g, sc = gscene
    sc.set_bvh(kind)
    try:
        for k in ("render_e14_16_", "render_e6_64_"):
            l = T(g[k + "l"], cuda_device)
            pts = T(g["pos"], cuda_device).unsqueeze(1).expand_as(l)
            rad = tx.query_irf(sc, pts, l.unsqueeze(-2), l.shape[1])
            assert np.array_equal(rad.cpu().numpy(), g[k + "lighting"])
    finally:
        sc.set_bvh(2)


def test_texture_swap_is_seen(tx, gscene, cuda_device):
    """stage -1 swaps self.texture for a thresholded copy and restores it (mat_nvdiffrast.py:143-150)"""
    name, g, sc = gscene
    l = T(g["render_e14_16_l"], cuda_device)
    pts = T(g["pos"], cuda_device)
    tex = torch.from_numpy(g["texture"]).clone()
    sc.set_texture(tex)
    a = tx.query_irf(sc, pts, l)
    sc.set_texture(tex * 2)
    b = tx.query_irf(sc, pts, l)
    tex.mul_(0.5)  # in-place edit of the same tensor: the version counter changes
    sc.set_texture(tex)
    c = tx.query_irf(sc, pts, l)
    sc.set_texture(torch.from_numpy(g["texture"]).to(cuda_device))  # device-resident source, CHW
    d = tx.query_irf(sc, pts, l)
    sc.set_texture(torch.from_numpy(g["texture"])[0].permute(1, 2, 0).contiguous())  # HWC
    e = tx.query_irf(sc, pts, l)
    assert torch.equal(b, a * 2) and torch.equal(c, a * 0.5) and torch.equal(d, a) and torch.equal(e, a)
    sc.set_texture(torch.from_numpy(g["texture"]))


# ---------------------------------------------------------------------------------------------------
# stand-alone BRDF reductions
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,N,eps", [("e14", 16, 1e-14), ("e6", 64, 1e-6)])
def test_specular_reflectance_fwd_bwd(tx, scene_golden, cuda_device, tag, N, eps):
    name, g = scene_golden
    k = "render_%s_%d_" % (tag, N)
    L = T(g[k + "lighting"], cuda_device).requires_grad_(True)
    h = T(g[k + "h"], cuda_device).requires_grad_(True)
    l = T(g[k + "l"], cuda_device).requires_grad_(True)
    r = T(g["rough"], cuda_device).requires_grad_(True)
    out = tx.specular_reflectance(L, h, T(g["nrm"], cuda_device), T(g[k + "view"], cuda_device), l, r, eps=eps)
    assert rel_l2(out.detach().cpu().numpy(), g[k + "spec"]) < 1e-5
    (out * T(g["g_rgb"], cuda_device)).sum().backward()
    assert rel_l2(L.grad.cpu().numpy(), g[k + "spec_gL"]) < 1e-5
    assert rel_l2(h.grad.cpu().numpy(), g[k + "spec_gh"]) < GRAD_TOL
    assert rel_l2(l.grad.cpu().numpy(), g[k + "spec_gl"]) < GRAD_TOL
    assert rel_l2(r.grad.cpu().numpy().reshape(-1), g[k + "spec_grough"]) < GRAD_TOL


@pytest.mark.parametrize("dmode", ["uniform", "cosine"])
def test_diffuse_reflectance_fwd_bwd(tx, oracle, scene_golden, cuda_device, dmode):
    name, g = scene_golden
    Lh, lh = g["diffuse_%s_lighting" % dmode], g["diffuse_%s_l" % dmode]
    L = T(Lh, cuda_device).requires_grad_(True)
    l = T(lh, cuda_device).requires_grad_(True)
    a = T(g["albedo"], cuda_device).requires_grad_(True)
    out = tx.diffuse_reflectance(L, l, T(g["nrm"], cuda_device), a, dmode)
    assert rel_l2(out.detach().cpu().numpy(), g["diffuse_%s_out" % dmode]) < 1e-5
    (out * T(g["g_rgb"], cuda_device)).sum().backward()
    Lc, lc, ac = (torch.from_numpy(x).requires_grad_(True) for x in (Lh, lh, g["albedo"]))
    ref = oracle.diffuse_reflectance(Lc, lc, torch.from_numpy(g["nrm"]), ac, dmode)
    (ref * torch.from_numpy(g["g_rgb"])).sum().backward()
    assert rel_l2(L.grad.cpu().numpy(), Lc.grad.numpy()) < 1e-5
    assert rel_l2(a.grad.cpu().numpy(), ac.grad.numpy()) < 1e-5
    if dmode == "uniform":
        assert rel_l2(l.grad.cpu().numpy(), lc.grad.numpy()) < GRAD_TOL
    else:
        assert torch.all(l.grad == 0)


# ---------------------------------------------------------------------------------------------------
# fused render forward + backward
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("backward", ["stash", "recompute"])
@pytest.mark.parametrize("tag,N,eps", [("e14", 16, 1e-14), ("e6", 64, 1e-6)])
def test_shade_fwd_bwd_vs_reference(tx, gscene, cuda_device, tag, N, eps, backward):
    name, g, sc = gscene
    k = "render_%s_%d_" % (tag, N)
    a = T(g["albedo"], cuda_device).requires_grad_(True)
    r = T(g["rough"], cuda_device).requires_grad_(True)
    i = T(g["irr"], cuda_device).requires_grad_(True)
    n0 = tx.launch_count()
    rgb = tx.shade(sc, T(g["nrm"], cuda_device), a, r, T(g["pos"], cuda_device), T(g["cam"], cuda_device), i,
                   T(g[k + "shift"], cuda_device), num_spec=N, eps=eps, backward=backward)
    (rgb * T(g["g_rgb"], cuda_device)).sum().backward()
    assert tx.launch_count() - n0 == 2  # ONE forward kernel + ONE backward kernel
    assert rel_l2(rgb.detach().cpu().numpy(), g[k + "rgb"]) < FWD_TOL
    assert rel_l2(a.grad.cpu().numpy(), g[k + "g_albedo"]) < GRAD_TOL
    assert rel_l2(r.grad.cpu().numpy().reshape(-1), g[k + "g_rough"]) < GRAD_TOL
    assert rel_l2(i.grad.cpu().numpy(), g[k + "g_irr"]) < GRAD_TOL


@pytest.mark.parametrize("dmode", ["uniform", "cosine"])
def test_relighting_render_vs_reference(tx, gscene, oracle, cuda_device, dmode):
    name, g, sc = gscene
    a = T(g["albedo"], cuda_device).requires_grad_(True)
    rgb = tx.shade(sc, T(g["nrm"], cuda_device), a, T(g["rough"], cuda_device), T(g["pos"], cuda_device),
                   T(g["cam"], cuda_device), None, T(g["relight_%s_shift_spec" % dmode], cuda_device), num_spec=16,
                   shift_diff=T(g["relight_%s_shift_diff" % dmode], cuda_device), num_diff=32, mode_diff=dmode, eps=1e-6)
    assert rel_l2(rgb.detach().cpu().numpy(), g["relight_%s_rgb" % dmode]) < FWD_TOL
    (rgb * T(g["g_rgb"], cuda_device)).sum().backward()
    ac = torch.from_numpy(g["albedo"]).requires_grad_(True)
    ref = oracle.render(oracle.OracleScene(g["vertices"], g["triangles"]), g["triangle_uvs"], torch.from_numpy(g["texture"]),
                        [32, 16], [dmode, "importance"], torch.from_numpy(g["nrm"]), ac, torch.from_numpy(g["rough"]),
                        torch.from_numpy(g["pos"]), torch.from_numpy(g["cam"]), torch.from_numpy(g["irr"]),
                        shift_spec=torch.from_numpy(g["relight_%s_shift_spec" % dmode]),
                        shift_diff=torch.from_numpy(g["relight_%s_shift_diff" % dmode]), relighting=True, eps=1e-6)
    (ref["rgb"] * torch.from_numpy(g["g_rgb"])).sum().backward()
    assert rel_l2(a.grad.cpu().numpy(), ac.grad.numpy()) < GRAD_TOL


@pytest.mark.parametrize("fused", [True, False])
def test_material_model_render_like_the_reference_call(tx, scene_golden, cuda_device, fused):
    """MaterialModel.render with the reference's calling convention ([6,h,w,c] tensors, RNG seeded by the
    caller): fused kernel and op-by-op composition both reproduce the reference output and gradients."""
    name, g = scene_golden
    mm = tx.MaterialModel(g["vertices"], g["triangles"], g["triangle_uvs"], torch.from_numpy(g["texture"]),
                          sample_l=[16, 16], device=cuda_device, fused=fused)
    shp = (6, 4, 4)
    a = T(g["albedo"], cuda_device).reshape(*shp, 3).requires_grad_(True)
    r = T(g["rough"], cuda_device).reshape(*shp, 1).requires_grad_(True)
    torch.manual_seed(666)
    res = mm.render(T(g["nrm"], cuda_device).reshape(*shp, 3), a, r, T(g["pos"], cuda_device).reshape(*shp, 3),
                    T(g["cam"], cuda_device), T(g["irr"], cuda_device).reshape(*shp, 3))
    assert set(res) == {"rgb", "albedo", "normal", "position"} and res["rgb"].shape == (*shp, 3)
    assert not res["normal"].requires_grad and not res["position"].requires_grad
    (res["rgb"] * T(g["g_rgb"], cuda_device).reshape(*shp, 3)).sum().backward()
    assert rel_l2(res["rgb"].detach().cpu().numpy().reshape(-1, 3), g["render_e14_16_rgb"]) < FWD_TOL
    assert rel_l2(a.grad.cpu().numpy().reshape(-1, 3), g["render_e14_16_g_albedo"]) < GRAD_TOL
    assert rel_l2(r.grad.cpu().numpy().reshape(-1), g["render_e14_16_g_rough"]) < GRAD_TOL
    want_pos = g["pos"] + np.float32(2e-2) * g["nrm"]
    assert np.allclose(res["position"].cpu().numpy().reshape(-1, 3), want_pos, atol=1e-6)
    # the trainer mutates sample_l (train_material_syn.py:735): it must be read at call time
    mm.sample_l[1] = 64
    mm.brdf_eps = 1e-6
    torch.manual_seed(666)
    res = mm.render(T(g["nrm"], cuda_device), T(g["albedo"], cuda_device), T(g["rough"], cuda_device),
                    T(g["pos"], cuda_device), T(g["cam"], cuda_device), T(g["irr"], cuda_device))
    assert rel_l2(res["rgb"].cpu().numpy(), g["render_e6_64_rgb"]) < FWD_TOL


def test_shade_albedo_only_backward_casts_no_rays(tx, gscene, cuda_device):
    """stage 0/1 style: only one of the materials needs a gradient"""
    name, g, sc = gscene
    a = T(g["albedo"], cuda_device).requires_grad_(True)
    rgb = tx.shade(sc, T(g["nrm"], cuda_device), a, T(g["rough"], cuda_device), T(g["pos"], cuda_device),
                   T(g["cam"], cuda_device), T(g["irr"], cuda_device), T(g["render_e14_16_shift"], cuda_device),
                   num_spec=16)
    (rgb * T(g["g_rgb"], cuda_device)).sum().backward()
    assert rel_l2(a.grad.cpu().numpy(), g["render_e14_16_g_albedo"]) < GRAD_TOL


# ---------------------------------------------------------------------------------------------------
# IrT
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [2, 8])
def test_irradiance_vs_reference(tx, gscene, cuda_device, kind):
    name, g, sc = gscene
    sc.set_bvh(kind)
    try:
        n0 = tx.launch_count()
        E = tx.irradiance(sc, T(g["irt_pos"], cuda_device), T(g["irt_nrm"], cuda_device), T(g["irt_shift"], cuda_device),
                          int(g["irt_N"]), "uniform")
        assert tx.launch_count() - n0 == 1
    finally:
        sc.set_bvh(2)
    assert rel_l2(E.cpu().numpy(), g["irt_E"]) < FWD_TOL
    seam = np.all(g["irt_nrm"] == 0, axis=1)
    assert torch.all(E[T(seam, cuda_device)] == 0)


def test_tracer_forward_like_the_reference(tx, scene_golden, cuda_device):
    name, g = scene_golden
    tr = tx.TracerO3d(g["vertices"], g["triangles"], g["triangle_uvs"], torch.from_numpy(g["texture"]),
                      sample_l=[int(g["irt_N"]), 16], device=cuda_device)
    pos = T(g["irt_pos"], cuda_device).reshape(32, 32, 3)
    nrm = T(g["irt_nrm"], cuda_device).reshape(32, 32, 3)
    idx = np.ones((32, 32, 3), np.float32)
    idx[3, 5] = 0  # a seam texel
    torch.manual_seed(666)  # reference stream: torch.rand(512,1,2) per batch == one torch.rand(P,1,2)
    irr = tr(pos, nrm, idx)
    want = g["irt_E"].reshape(32, 32, 3).copy()
    want[3, 5] = 0
    assert irr.shape == (32, 32, 3) and rel_l2(irr.cpu().numpy(), want) < FWD_TOL
    assert torch.all(irr[3, 5] == 0)
    with pytest.raises(AssertionError):
        tr(pos[:, :5], nrm[:, :5])  # texel count not a multiple of 512 (tracer_o3d_irt.py:161)


def test_irradiance_edge_cases(tx, gscene, cuda_device):
    name, g, sc = gscene
    pos, nrm, sh = T(g["irt_pos"], cuda_device), T(g["irt_nrm"], cuda_device), T(g["irt_shift"], cuda_device)
    assert tx.irradiance(sc, pos[:0], nrm[:0], sh[:0], 8).shape == (0, 3)
    for N in (1, 3, 7, 31, 33, 100, 2048):  # group sizes below / above a warp, non powers of two
        E = tx.irradiance(sc, pos[:37], nrm[:37], sh[:37], N)
        assert E.shape == (37, 3) and torch.isfinite(E).all() and (E >= 0).all()
    # the fused kernel == generate_dir + query_irf + reduction composed from the stand-alone ops
    N = 48
    l = tx.generate_dir(nrm[:64], N, None, "uniform", shift=sh[:64])
    rad = tx.query_irf(sc, pos[:64], l)
    ndl = torch.clamp((nrm[:64].unsqueeze(1) * l).sum(-1, keepdim=True), 0, 1)
    ref = (rad * ndl).sum(1) * 2 * np.pi / N
    E = tx.irradiance(sc, pos[:64], nrm[:64], sh[:64], N)
    assert rel_l2(E.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    with pytest.raises(tx.TexirError):
        tx.irradiance(sc, pos[:4], nrm[:4], sh[:4], 8, "importance")


def test_radiance_samples_is_the_nirf_target_generator(tx, gscene, oracle, cuda_device):
    """models/tracer_o3d.py:73-116: generate_dir(uniform, h*w Hammersley + shift) -> cast -> texture, un-reduced"""
    name, g, sc = gscene
    P, N = 64, 8 * 16  # resolution [8,16] like train.env_res
    pos, nrm, sh = T(g["irt_pos"][:P], cuda_device), T(g["irt_nrm"][:P], cuda_device), T(g["irt_shift"][:P], cuda_device)
    n0 = tx.launch_count()
    rad, dirs = tx.radiance_samples(sc, pos, nrm, sh, N, return_dirs=True)
    assert tx.launch_count() - n0 == 1 and rad.shape == (P, N, 3)
    l = tx.generate_dir(nrm, N, None, "uniform", shift=sh)
    assert torch.equal(dirs, l) and torch.equal(rad, tx.query_irf(sc, pos, l))
    lo = oracle.generate_dir(torch.from_numpy(g["irt_nrm"][:P]), N, None, "uniform", shift=torch.from_numpy(g["irt_shift"][:P]))
    ro = oracle.query_irf(oracle.OracleScene(g["vertices"], g["triangles"]), g["triangle_uvs"], torch.from_numpy(g["texture"]),
                          torch.from_numpy(g["irt_pos"][:P]).unsqueeze(1).expand_as(lo), lo.unsqueeze(-2), N)
    assert rel_l2(rad.cpu().numpy(), ro.numpy()) < FWD_TOL


def test_gbuffer_from_primary_rays(tx, gscene, oracle, cuda_device):
    from texir_code_b200 import synth
    name, g, sc = gscene
    rays = synth.cube_map_rays(16, cam=(0.2, -0.1, 0.3))
    gb = tx.gbuffer(sc, T(rays, cuda_device).reshape(6, 16, 16, 6))
    assert gb["position"].shape == (6, 16, 16, 3) and gb["texc"].shape == (6, 16, 16, 2) and gb["mask"].shape == (6, 16, 16, 1)
    t0, i0, u0 = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays)
    assert np.array_equal(gb["primitive_ids"].cpu().numpy().reshape(-1).view(np.uint32), i0)
    assert np.array_equal(gb["primitive_uvs"].cpu().numpy().reshape(-1, 2), u0)
    hit = i0 != 0xFFFFFFFF
    assert np.array_equal(gb["mask"].cpu().numpy().reshape(-1) > 0, hit) and hit.mean() > 0.95
    x = rays[:, :3].astype(np.float64) + rays[:, 3:].astype(np.float64) * t0[:, None].astype(np.float64)
    got = gb["position"].cpu().numpy().reshape(-1, 3)
    assert np.abs(got[hit] - x[hit]).max() < 2e-6
    tri = g["vertices"][g["triangles"][np.where(hit, i0, 0).astype(np.int64)]].astype(np.float64)
    ng = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    ng /= np.linalg.norm(ng, axis=1, keepdims=True)
    ng *= -np.sign((ng * rays[:, 3:]).sum(1, keepdims=True))  # facing the camera
    assert np.abs(gb["normal"].cpu().numpy().reshape(-1, 3)[hit] - ng[hit]).max() < 1e-5
    # texc = the reference's UV interpolation (mat_nvdiffrast.py:303-311)
    uvc = g["triangle_uvs"].reshape(-1, 3, 2)[np.where(hit, i0, 0).astype(np.int64)]
    b = np.clip(u0, 0, 1)
    uv = uvc[:, 0] * (1 - b[:, 0:1] - b[:, 1:2]) + uvc[:, 1] * b[:, 0:1] + uvc[:, 2] * b[:, 1:2]
    assert np.abs(gb["texc"].cpu().numpy().reshape(-1, 2)[hit] - uv[hit].astype(np.float32)).max() <= 1e-7
    if (~hit).any():
        assert np.all(got[~hit] == [1, 0, 0])
    off = tx.gbuffer(sc, T(rays, cuda_device), offset=1e-2)["position"].cpu().numpy()
    assert np.abs(off[hit] - (got[hit] + 1e-2 * gb["normal"].cpu().numpy().reshape(-1, 3)[hit])).max() < 1e-6


def test_material_textures_receive_scattered_gradients(tx, gscene, oracle, cuda_device):
    """G-buffer -> bilinear material fetch -> fused render -> backward: the gradients land in the trainable
    textures materials_a / materials_r through the atomic scatter, compared with the same chain in
    torch (grid_sample) on top of the oracle's render"""
    name, g, _ = gscene
    mm = tx.MaterialModel(g["vertices"], g["triangles"], g["triangle_uvs"], torch.from_numpy(g["texture"]),
                          sample_l=[16, 16], device=cuda_device, materials_a_res=32, materials_r_res=64)
    rng = np.random.default_rng(7)
    with torch.no_grad():
        mm.materials_a.copy_(T(rng.uniform(0.1, 0.9, size=(32, 32, 3)).astype(np.float32), cuda_device))
        mm.materials_r.copy_(T(rng.uniform(0.05, 0.7, size=(64, 64, 1)).astype(np.float32), cuda_device))
    P = len(g["pos"])
    texc = rng.uniform(0.05, 0.95, size=(P, 2)).astype(np.float32)
    torch.manual_seed(666)
    out = mm.forward_gbuffer(T(g["nrm"], cuda_device), T(g["pos"], cuda_device), T(g["cam"], cuda_device),
                             T(g["irr"], cuda_device), texc=T(texc, cuda_device))
    (out["rgb"] * T(g["g_rgb"], cuda_device)).sum().backward()
    # torch / oracle chain on the CPU
    ta = mm.materials_a.detach().cpu().clone().requires_grad_(True)
    tr = mm.materials_r.detach().cpu().clone().requires_grad_(True)
    grid = (torch.from_numpy(texc) * 2 - 1).reshape(1, P, 1, 2)
    fetch = lambda t: torch.nn.functional.grid_sample(t.permute(2, 0, 1).unsqueeze(0), grid, mode="bilinear",
                                                      padding_mode="border", align_corners=False)[0, :, :, 0].t()
    ref = oracle.render(oracle.OracleScene(g["vertices"], g["triangles"]), g["triangle_uvs"], torch.from_numpy(g["texture"]),
                        [16, 16], ["uniform", "importance"], torch.from_numpy(g["nrm"]), fetch(ta), fetch(tr),
                        torch.from_numpy(g["pos"]), torch.from_numpy(g["cam"]), torch.from_numpy(g["irr"]),
                        shift_spec=torch.from_numpy(g["render_e14_16_shift"]))
    (ref["rgb"] * torch.from_numpy(g["g_rgb"])).sum().backward()
    assert rel_l2(out["rgb"].detach().cpu().numpy(), ref["rgb"].detach().numpy()) < FWD_TOL
    assert rel_l2(mm.materials_a.grad.cpu().numpy(), ta.grad.numpy()) < GRAD_TOL
    assert rel_l2(mm.materials_r.grad.cpu().numpy(), tr.grad.numpy()) < GRAD_TOL
    assert (mm.materials_a.grad != 0).any() and (mm.materials_r.grad != 0).any()


def test_point_order_is_a_morton_permutation_and_changes_no_value(tx, big, cuda_device):
    from texir_code_b200 import ops, synth
    sc, tex, scene = big
    P = 20000
    pos, nrm = synth.sample_surface_points(sc, P, seed=12)
    posd, nrmd = T(pos, cuda_device), T(nrm, cuda_device)
    order = ops.point_order(scene, posd).cpu().numpy()
    assert order.dtype == np.int32 and np.array_equal(np.sort(order), np.arange(P))
    # consecutive points of the order are spatial neighbours: mean step far below the random-order step
    live = ~np.all(nrm == 0, axis=1)
    step_sorted = np.linalg.norm(np.diff(pos[order][live[order]], axis=0), axis=1).mean()
    step_random = np.linalg.norm(np.diff(pos[live], axis=0), axis=1).mean()
    assert step_sorted < 0.1 * step_random
    torch.manual_seed(3)
    sh = torch.rand(P, 1, 2).to(cuda_device)
    a = tx.irradiance(scene, posd, nrmd, sh, 64, coherent=True)
    b = tx.irradiance(scene, posd, nrmd, sh, 64, coherent=False)
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------------
# material texture fetch / scatter
# ---------------------------------------------------------------------------------------------------
def test_texture_fetch_and_scatter(tx, cuda_device):
    rng = np.random.default_rng(2)
    H, W, C, P = 16, 32, 3, 500
    tex = T(rng.normal(size=(H, W, C)).astype(np.float32), cuda_device).requires_grad_(True)
    uv = T(rng.uniform(0.05, 0.95, size=(P, 2)).astype(np.float32), cuda_device)
    out = tx.texture_fetch(tex, uv)
    # torch reference: bilinear with texel centres at (i+0.5)/size == grid_sample(align_corners=False)
    grid = (uv * 2 - 1).reshape(1, P, 1, 2)
    t2 = tex.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.grid_sample(t2.permute(2, 0, 1).unsqueeze(0), grid, mode="bilinear", padding_mode="border",
                                          align_corners=False)[0, :, :, 0].t()
    assert rel_l2(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
    gout = T(rng.normal(size=(P, C)).astype(np.float32), cuda_device)
    (out * gout).sum().backward()
    (ref * gout).sum().backward()
    assert rel_l2(tex.grad.cpu().numpy(), t2.grad.cpu().numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------------
# full-size properties (T = 2^20 triangles, 4096^2 texture: BASELINE config sizes)
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big(tx, cuda_device):
    from texir_code_b200 import synth
    sc = synth.make_room_scene(1 << 20, seed=666)
    tex = torch.from_numpy(synth.make_radiance_texture(4096, seed=666))
    scene = tx.Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=cuda_device)
    scene.set_texture(tex)
    return sc, tex, scene


def test_full_size_scene_properties(tx, oracle, big, cuda_device):
    from texir_code_b200 import synth
    sc, tex, scene = big
    info = scene.info()
    assert info["num_tris"] == 1 << 20 and info["depth8"] < 40 and info["tex_h"] == 4096
    P, N = 4096, 256
    pos, nrm = synth.sample_surface_points(sc, P, seed=1)
    posd, nrmd = T(pos, cuda_device), T(nrm, cuda_device)
    torch.manual_seed(666)
    sh = torch.rand(P, 1, 2).to(cuda_device)
    # (1) BVH2 and BVH8 traversals agree bit-for-bit on 1M rays, and with the oracle's own BVH on a sample
    l = tx.generate_dir(posd * 0 + nrmd, N, None, "uniform", shift=sh)
    rays = torch.cat([posd.unsqueeze(1).expand_as(l), l], -1).reshape(-1, 6)
    scene.set_bvh(2)
    r2 = scene.cast_rays(rays)
    scene.set_bvh(8)
    r8 = scene.cast_rays(rays)
    scene.set_bvh(2)
    for k in ("t_hit", "primitive_ids", "primitive_uvs"):
        assert torch.equal(r2[k], r8[k]), k
    sel = torch.arange(0, rays.shape[0], 53, device=cuda_device)[:20000]
    t0, i0, u0 = oracle.OracleScene(sc.vertices, sc.triangles).cast_rays_np(rays[sel].cpu().numpy())
    assert np.array_equal(r8["primitive_ids"][sel].cpu().numpy().view(np.uint32), i0)
    assert np.array_equal(r8["t_hit"][sel].cpu().numpy(), t0)
    live = torch.from_numpy(~np.all(nrm == 0, axis=1)).to(cuda_device)
    hit = torch.isfinite(r8["t_hit"]).reshape(P, N)
    assert hit[live].float().mean() > 0.995  # closed room (T-junction cracks from the exact-count edge bisection leak ~0.1 %)
    # (2) linearity in the radiance texture: scaling by 2 (exact in fp32) scales E exactly
    E1 = tx.irradiance(scene, posd, nrmd, sh, N)
    scene.set_texture(tex * 2)
    E2 = tx.irradiance(scene, posd, nrmd, sh, N)
    assert torch.equal(E2, E1 * 2)
    # (3) constant radiance 1 in a closed room: E = 2 pi / N * sum(clamp(n.l)) -> pi (cosine integral)
    scene.set_texture(torch.ones(1, 3, 64, 64))
    Ec = tx.irradiance(scene, posd, nrmd, sh, N)[live]
    assert abs(Ec.mean().item() - np.pi) < 0.02 and (Ec.max() < 3.5).item()
    scene.set_texture(tex)
    # (4) the fused kernel against the oracle on a 512-point batch at full scene size
    Eo = oracle.irt_forward(oracle.OracleScene(sc.vertices, sc.triangles), sc.triangle_uvs, tex, torch.from_numpy(pos[:512]),
                            torch.from_numpy(nrm[:512]), N, "uniform", 512, sh[:512].cpu())
    assert rel_l2(E1[:512].cpu().numpy(), Eo.numpy()) < FWD_TOL


def test_full_size_material_step(tx, oracle, big, cuda_device):
    """C3-shaped material step (cube-map G-buffer from primary rays, fwd + bwd) against the oracle on a slice"""
    from texir_code_b200 import synth
    sc, tex, scene = big
    rays = synth.cube_map_rays(32)
    hit = scene.cast_rays(T(rays, cuda_device))
    pos, nrm = synth.gbuffer_from_hits(sc, rays, hit["t_hit"].cpu().numpy(), hit["primitive_ids"].cpu().numpy().view(np.uint32))
    P = pos.shape[0]
    rough, albedo, irr, g_rgb = synth.make_materials(P, seed=8)
    cam = np.zeros(3, np.float32)
    torch.manual_seed(666)
    sh = torch.rand(P, 1, 2)
    N = 64
    a = T(albedo, cuda_device).requires_grad_(True)
    r = T(rough, cuda_device).requires_grad_(True)
    rgb = tx.shade(scene, T(nrm, cuda_device), a, r, T(pos, cuda_device), T(cam, cuda_device), T(irr, cuda_device),
                   sh.to(cuda_device), num_spec=N)
    (rgb * T(g_rgb, cuda_device)).sum().backward()
    # like-for-like oracle: the reference evaluates generate_dir / the BRDF on the CUDA device and only query_irf on the
    # host (mat_nvdiffrast.py:294); on the i.i.d.-noise atlas torch-CPU directions alone differ by ~4e-5 rel-L2
    sl = slice(0, 768)
    ac = T(albedo[sl], cuda_device).requires_grad_(True)
    rc = T(rough[sl], cuda_device).requires_grad_(True)
    ref = oracle.render(oracle.OracleScene(sc.vertices, sc.triangles), sc.triangle_uvs, tex, [N, N], ["uniform", "importance"],
                        T(nrm[sl], cuda_device), ac, rc, T(pos[sl], cuda_device), T(cam, cuda_device),
                        T(irr[sl], cuda_device), shift_spec=sh[sl].to(cuda_device))
    (ref["rgb"] * T(g_rgb[sl], cuda_device)).sum().backward()
    assert rel_l2(rgb[sl].detach().cpu().numpy(), ref["rgb"].detach().cpu().numpy()) < FWD_TOL
    assert rel_l2(a.grad[sl].cpu().numpy(), ac.grad.cpu().numpy()) < GRAD_TOL
    assert rel_l2(r.grad[sl].cpu().numpy(), rc.grad.cpu().numpy()) < GRAD_TOL


# ---------------------------------------------------------------------------------------------------
# system level: the optimisation loop of trainer/train_material.py (stage 2) on top of the ops
# ---------------------------------------------------------------------------------------------------
def test_material_optimisation_loop_recovers_textures(tx, gscene, cuda_device):
    """Adam on materials_a / materials_r through G-buffer -> texture fetch -> fused render -> log-L1 loss
    (models/loss.py:81-99), with the trainer's clamps (train_material.py:458,592-593): the loss must fall."""
    from texir_code_b200 import synth
    name, g, _ = gscene
    if name != "smooth":
        pytest.skip("the i.i.d.-noise atlas texture makes a 32-spp roughness landscape too rough for a 100-step test")
    mm = tx.MaterialModel(g["vertices"], g["triangles"], g["triangle_uvs"], torch.from_numpy(g["texture"]),
                          sample_l=[32, 32], device=cuda_device, materials_a_res=16, materials_r_res=16)
    rays = synth.cube_map_rays(12, cam=(0.1, 0.0, 0.2))
    gb = tx.gbuffer(mm.scene, T(rays, cuda_device), offset=1e-2)
    keep = gb["mask"].reshape(-1) > 0
    nrm, pos, texc = gb["normal"][keep], gb["position"][keep], gb["texc"][keep]
    cam = torch.tensor([0.1, 0.0, 0.2], device=cuda_device)
    irr = torch.full_like(pos, 3.0)
    rng = np.random.default_rng(0)
    gt_a = T(rng.uniform(0.2, 0.9, size=(16, 16, 3)).astype(np.float32), cuda_device)
    gt_r = T(rng.uniform(0.1, 0.6, size=(16, 16, 1)).astype(np.float32), cuda_device)
    shift = torch.rand(pos.shape[0], 1, 2, generator=torch.Generator().manual_seed(1)).to(cuda_device)

    def render(a_tex, r_tex):
        a = tx.texture_fetch(a_tex, texc)
        r = tx.texture_fetch(r_tex, texc)
        return tx.shade(mm.scene, nrm, a, r, pos, cam, irr, shift, num_spec=32)

    mm.scene.set_texture(mm.texture)
    with torch.no_grad():
        target = render(gt_a, gt_r)
    opt = torch.optim.Adam([mm.materials_a, mm.materials_r], lr=3e-2)
    losses = []
    for it in range(100):
        opt.zero_grad()
        loss = (torch.log1p(render(mm.materials_a, mm.materials_r)) - torch.log1p(target)).abs().mean()
        loss.backward()
        opt.step()
        with torch.no_grad():
            mm.materials_r.clamp_(0.01, 0.8)
            mm.materials_a.clamp_(min=0.0)
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])
    assert set(mm.state_dict()) >= {"materials_a", "materials_r"}  # the reference's checkpoint keys


def test_shared_first_descent_switch_changes_no_bit(cuda_device, tmp_path):
    """TEXIR_FRONT=0 (every ray walks from the root) and the default (shared first descent) must give bit-identical
    irradiance and render results; the switch is read once per process, hence two subprocesses"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import texir_code_b200 as tx
from texir_code_b200 import synth
dev = torch.device('cuda:0')
sc = synth.make_room_scene(60000, seed=666, clutter=16)
tex = torch.from_numpy(synth.make_radiance_texture(512, seed=666))
scene = tx.Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev)
scene.set_texture(tex)
pos, nrm = synth.sample_surface_points(sc, 3000, seed=3, seam_fraction=0.05)
pos[:3] = [[0, 0, 0], [50, 0, 0], [3.99, 1.49, 2.99]]
torch.manual_seed(1)
sh = torch.rand(3000, 1, 2).to(dev)
P = torch.from_numpy(pos).to(dev); N = torch.from_numpy(nrm).to(dev)
E = tx.irradiance(scene, P, N, sh, 256)
rough, albedo, irr, g = synth.make_materials(3000, seed=2)
rgb = tx.shade(scene, N, torch.from_numpy(albedo).to(dev), torch.from_numpy(rough).to(dev), P, torch.tensor([0.1, 0.2, 0.3], device=dev),
               torch.from_numpy(irr).to(dev), sh, num_spec=64)
np.savez(sys.argv[1], E=E.cpu().numpy(), rgb=rgb.cpu().numpy())
""" % root
    outs = []
    for flag in ("0", "1"):
        path = str(tmp_path / ("front%s.npz" % flag))
        env = dict(os.environ, TEXIR_FRONT=flag)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
        outs.append(np.load(path))
    assert np.array_equal(outs[0]["E"].view(np.uint32), outs[1]["E"].view(np.uint32))
    assert np.array_equal(outs[0]["rgb"].view(np.uint32), outs[1]["rgb"].view(np.uint32))
    assert np.isfinite(outs[0]["E"]).all() and float(outs[0]["E"].sum()) > 0
