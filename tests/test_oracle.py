"""CPU: the oracle (oracle/texir_oracle.py + oracle/cast_rays_oracle.c) against the golden vectors written by
executing the REFERENCE's own Python (oracle/make_golden.py).  The reference has no tests / KATs of its own
(SURVEY.md section 4), so these goldens are the pin."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2


def test_hammersley_known_answers(oracle):
    g = load_golden("hammersley")
    tab = oracle.hammersley_table(16)
    assert np.array_equal(tab, g["ham16"].astype(np.float32))
    for i, want in zip(g["ri_in"], g["ri_out"]):
        assert oracle.radical_inverse(int(i)) == want
    # SURVEY section 4 probes
    assert oracle.radical_inverse(1) == 0.5 and oracle.radical_inverse(3) == 0.75
    assert list(tab[5]) == [0.3125, 0.625]


@pytest.mark.parametrize("N", [16, 64])
@pytest.mark.parametrize("mode", ["uniform", "cosine", "importance"])
def test_generate_dir_matches_reference(oracle, N, mode):
    g = load_golden("generate_dir")
    nrm = torch.tensor(g["normals"])
    rough = torch.tensor(g["roughness"], requires_grad=True)
    L = oracle.generate_dir(nrm, N, None, mode, rough if mode == "importance" else None, shift=torch.tensor(g["shift"]))
    assert np.array_equal(L.detach().numpy(), g["L_%s_%d" % (mode, N)])
    assert np.all(L.detach().numpy()[3] == 0.0)  # zero normal -> zero directions, no NaN
    if mode == "importance":
        (L * torch.tensor(g["gL_%d" % N])).sum().backward()
        assert rel_l2(rough.grad.numpy(), g["grough_%d" % N]) < 1e-6


def test_generate_dir_draws_reference_shift_stream(oracle):
    g = load_golden("generate_dir")
    torch.manual_seed(666)
    L = oracle.generate_dir(torch.tensor(g["normals"]), 16, None, "uniform")
    assert np.array_equal(L.numpy(), g["L_uniform_16"])
    # consecutive 512-row draws == one big draw (TracerO3d relies on it)
    torch.manual_seed(1)
    a = torch.cat([torch.rand(512, 1, 2) for _ in range(3)])
    torch.manual_seed(1)
    assert torch.equal(a, torch.rand(1536, 1, 2))


def test_query_irf_quad_known_answers(oracle):
    g = load_golden("query_irf_quad")
    for brute in (True, False):
        sc = oracle.OracleScene(g["vertices"], g["triangles"], brute=brute)
        rad = oracle.query_irf(sc, g["triangle_uvs"], torch.tensor(g["texture"]), torch.tensor(g["org"]).unsqueeze(1),
                               torch.tensor(g["dir"]).reshape(6, 1, 1, 3), 1)
        assert np.array_equal(rad.numpy(), g["rad"])
    # SURVEY appendix C probes
    np.testing.assert_allclose(g["rad"][0, 0], [22.5, 23.5, 24.5], rtol=1e-6)
    np.testing.assert_allclose(g["rad"][1, 0], [29.7, 30.7, 31.7], rtol=1e-6)
    np.testing.assert_allclose(g["rad"][2, 0], [37.5, 38.5, 39.5], rtol=1e-6)
    assert np.all(g["rad"][3] == 0) and np.all(g["rad"][4] == 0)


def test_cast_rays_contract_and_bvh_equals_brute(oracle):
    g = load_golden("scene_atlas")
    rng = np.random.default_rng(3)
    R = 4000
    o = rng.uniform([-3.5, -1.2, -2.5], [3.5, 1.2, 2.5], size=(R, 3)).astype(np.float32)
    d = rng.normal(size=(R, 3)).astype(np.float32)
    d[:5] = 0.0  # zero directions never hit
    rays = np.concatenate([o, d], 1)
    tb, ib, ub = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays)
    tv, iv, uv = oracle.OracleScene(g["vertices"], g["triangles"], brute=False).cast_rays_np(rays)
    assert np.array_equal(tb, tv) and np.array_equal(ib, iv) and np.array_equal(ub, uv)
    miss = ib == 0xFFFFFFFF
    assert miss[:5].all() and np.isinf(tb[miss]).all() and (ub[miss] == 0).all()
    hit = ~miss
    assert hit.mean() > 0.9  # closed room
    # hit point = (1-u-v) v0 + u v1 + v v2  (Open3D's primitive_uvs convention)
    tri = g["vertices"][g["triangles"][ib[hit].astype(np.int64)]].astype(np.float64)
    u, v = ub[hit, 0:1].astype(np.float64), ub[hit, 1:2].astype(np.float64)
    x = tri[:, 0] * (1 - u - v) + tri[:, 1] * u + tri[:, 2] * v
    y = o[hit].astype(np.float64) + d[hit].astype(np.float64) * tb[hit, None]
    assert np.abs(x - y).max() < 5e-4
    # Open3D-style front end
    res = oracle.OracleScene(g["vertices"], g["triangles"]).cast_rays(rays.reshape(40, 100, 6))
    assert res["t_hit"].numpy().shape == (40, 100) and res["primitive_uvs"].numpy().shape == (40, 100, 2)
    assert res["primitive_ids"].numpy().dtype == np.uint32


@pytest.mark.parametrize("tag,N,eps", [("e14", 16, 1e-14), ("e6", 64, 1e-6)])
def test_render_matches_reference(oracle, scene_golden, tag, N, eps):
    name, g = scene_golden
    sc = oracle.OracleScene(g["vertices"], g["triangles"])
    k = "render_%s_%d_" % (tag, N)
    a = torch.tensor(g["albedo"], requires_grad=True)
    r = torch.tensor(g["rough"], requires_grad=True)
    i = torch.tensor(g["irr"], requires_grad=True)
    out = oracle.render(sc, g["triangle_uvs"], torch.tensor(g["texture"]), [N, N], ["uniform", "importance"],
                        torch.tensor(g["nrm"]), a, r, torch.tensor(g["pos"]), torch.tensor(g["cam"]), i,
                        shift_spec=torch.tensor(g[k + "shift"]), eps=eps)
    (out["rgb"] * torch.tensor(g["g_rgb"])).sum().backward()
    assert np.array_equal(out["rgb"].detach().numpy(), g[k + "rgb"])
    assert rel_l2(a.grad.numpy(), g[k + "g_albedo"]) < 1e-6
    assert rel_l2(r.grad.numpy().reshape(-1), g[k + "g_rough"]) < 1e-5
    assert rel_l2(i.grad.numpy(), g[k + "g_irr"]) < 1e-6
    # the stand-alone pieces
    spec = oracle.specular_reflectance(torch.tensor(g[k + "lighting"]), torch.tensor(g[k + "h"]), torch.tensor(g["nrm"]),
                                       torch.tensor(g[k + "view"]), torch.tensor(g[k + "l"]), torch.tensor(g["rough"]),
                                       eps=eps)
    assert rel_l2(spec.numpy(), g[k + "spec"]) < 1e-6  # summation order of a strided reduce differs by <= 1 ulp


@pytest.mark.parametrize("dmode", ["uniform", "cosine"])
def test_relighting_matches_reference(oracle, scene_golden, dmode):
    name, g = scene_golden
    sc = oracle.OracleScene(g["vertices"], g["triangles"])
    out = oracle.render(sc, g["triangle_uvs"], torch.tensor(g["texture"]), [32, 16], [dmode, "importance"],
                        torch.tensor(g["nrm"]), torch.tensor(g["albedo"]), torch.tensor(g["rough"]),
                        torch.tensor(g["pos"]), torch.tensor(g["cam"]), torch.tensor(g["irr"]),
                        shift_spec=torch.tensor(g["relight_%s_shift_spec" % dmode]),
                        shift_diff=torch.tensor(g["relight_%s_shift_diff" % dmode]), relighting=True, eps=1e-6)
    assert rel_l2(out["rgb"].numpy(), g["relight_%s_rgb" % dmode]) < 1e-6
    d = oracle.diffuse_reflectance(torch.tensor(g["diffuse_%s_lighting" % dmode]), torch.tensor(g["diffuse_%s_l" % dmode]),
                                   torch.tensor(g["nrm"]), torch.tensor(g["albedo"]), dmode)
    assert rel_l2(d.numpy(), g["diffuse_%s_out" % dmode]) < 1e-6


def test_irt_matches_reference(oracle, scene_golden):
    name, g = scene_golden
    sc = oracle.OracleScene(g["vertices"], g["triangles"])
    E = oracle.irt_forward(sc, g["triangle_uvs"], torch.tensor(g["texture"]), torch.tensor(g["irt_pos"]),
                           torch.tensor(g["irt_nrm"]), int(g["irt_N"]), "uniform", 512, torch.tensor(g["irt_shift"]))
    assert np.array_equal(E.numpy(), g["irt_E"])
    seam = np.all(g["irt_nrm"] == 0, axis=1)
    assert seam.any() and np.all(E.numpy()[seam] == 0)


def test_uv_derivative_formula_vs_finite_differences(oracle):
    """oracle.gbuffer_uv_da (the float64 ray-differential formula the G-buffer kernels are checked against) equals
    central differences of the interpolated UVs along the ray differentials, on the pixels whose displaced rays stay on
    the same triangle"""
    from texir_code_b200 import synth
    sc = synth.make_room_scene(3000, seed=666, clutter=6, tex_res=64, uv_kind="planar")
    scene = oracle.OracleScene(sc.vertices, sc.triangles, brute=True)
    res = 12
    rays = synth.cube_map_rays(res, cam=(0.2, -0.1, 0.3)).astype(np.float64)
    dd = synth.cube_map_ray_differentials(res).astype(np.float64)

    def uv_of(r):
        t, ids, bary = scene.cast_rays_np(r.astype(np.float32))
        hit = ids != 0xFFFFFFFF
        tri = np.where(hit, ids, 0).astype(np.int64)
        v = sc.vertices.astype(np.float64)[sc.triangles[tri]]
        # exact barycentrics of the float64 ray on the hit triangle's plane (the float32 ones are too coarse for a
        # finite difference)
        o, d = r[:, :3], r[:, 3:]
        e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
        n = np.cross(e1, e2)
        tt = np.einsum("ij,ij->i", n, v[:, 0] - o) / np.einsum("ij,ij->i", n, d)
        x = o + tt[:, None] * d - v[:, 0]
        nn = np.einsum("ij,ij->i", n, n)
        b1 = np.einsum("ij,ij->i", np.cross(x, e2), n) / nn
        b2 = np.einsum("ij,ij->i", np.cross(e1, x), n) / nn
        tuv = np.asarray(sc.triangle_uvs, np.float64).reshape(-1, 3, 2)[tri]
        uv = tuv[:, 0] * (1 - b1 - b2)[:, None] + tuv[:, 1] * b1[:, None] + tuv[:, 2] * b2[:, None]
        return uv, ids, hit

    uv0, ids0, hit0 = uv_of(rays)
    an = oracle.gbuffer_uv_da(sc.vertices, sc.triangles, sc.triangle_uvs, rays, dd, ids0)
    h = 1e-3
    for k in range(2):
        step = np.concatenate([np.zeros((len(rays), 3)), dd[:, 3 * k:3 * k + 3] * h], 1)
        up, idp, _ = uv_of(rays + step)
        um, idm, _ = uv_of(rays - step)
        same = hit0 & (idp == ids0) & (idm == ids0)
        assert same.mean() > 0.5
        fd = (up - um) / (2 * h)
        assert rel_l2(an[same][:, [k, 2 + k]], fd[same]) < 1e-5
    assert np.all(an[~hit0] == 0)


def test_closest_hit_oracle_vs_float64_geometry(oracle):
    """The closest-hit oracle (the one function whose parity against real Embree is unpinned) against an independent
    float64 evaluation of the same geometry: same primitive wherever the two nearest candidates are not within 1e-6 of
    each other, hit distance and barycentrics to float32 accuracy, misses agree away from edges"""
    from texir_code_b200 import synth
    sc = synth.make_room_scene(1500, seed=666, clutter=4, tex_res=32, uv_kind="planar")
    scene = oracle.OracleScene(sc.vertices, sc.triangles, brute=True)
    rng = np.random.default_rng(2)
    R = 400
    o = rng.uniform([-3.5, -1.2, -2.5], [3.5, 1.2, 2.5], size=(R, 3))
    d = rng.normal(size=(R, 3))
    rays = np.concatenate([o, d], 1).astype(np.float32)
    t, ids, bary = scene.cast_rays_np(rays)
    o64, d64 = rays[:, :3].astype(np.float64), rays[:, 3:].astype(np.float64)
    v = sc.vertices.astype(np.float64)[sc.triangles]                              # [T,3,3]
    e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
    # Moeller-Trumbore in float64 for every (ray, triangle)
    pvec = np.cross(d64[:, None, :], e2[None])
    det = np.einsum("rtj,tj->rt", pvec, e1)
    tvec = o64[:, None, :] - v[None, :, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.einsum("rtj,rtj->rt", tvec, pvec) / det
        qvec = np.cross(tvec, e1[None])
        w = np.einsum("rj,rtj->rt", d64, qvec) / det
        tt = np.einsum("tj,rtj->rt", e2, qvec) / det
    inside = (np.abs(det) > 0) & (u >= 0) & (w >= 0) & (u + w <= 1) & (tt > 0)
    tt = np.where(inside, tt, np.inf)
    order = np.argsort(tt, axis=1)
    best, second = order[:, 0], order[:, 1]
    tb = tt[np.arange(R), best]
    ts = tt[np.arange(R), second]
    hit64 = np.isfinite(tb)
    margin = np.min(np.stack([u, w, 1 - u - w], -1)[np.arange(R), best], axis=-1)  # distance from the edges, barycentric
    clear = hit64 & (margin > 1e-5) & ((ts - tb) > 1e-6 * np.maximum(tb, 1.0))
    assert clear.mean() > 0.9
    assert np.array_equal(ids[clear], best[clear].astype(np.uint32))
    assert np.max(np.abs(t[clear] - tb[clear]) / tb[clear]) < 1e-5
    ub, wb = u[np.arange(R), best], w[np.arange(R), best]
    assert np.max(np.abs(bary[clear, 0] - ub[clear])) < 1e-4 and np.max(np.abs(bary[clear, 1] - wb[clear])) < 1e-4
    with np.errstate(divide="ignore", invalid="ignore"):
        t_raw = np.einsum("tj,rtj->rt", e2, qvec) / det
    edge_dist = np.abs(np.minimum(np.minimum(u, w), 1 - u - w))                  # 0 on an edge of the triangle
    near_edge = np.min(np.where((np.abs(det) > 0) & (t_raw > 0), edge_dist, np.inf), axis=1) < 1e-5
    miss = ids == 0xFFFFFFFF
    assert np.array_equal(miss[~near_edge], ~hit64[~near_edge])
    assert np.all(np.isinf(t[miss])) and np.all(bary[miss] == 0)


def test_mt19937_restatement_is_torch_rand():
    """oracle/mt19937_oracle.py (the checker of the device-side shift generator) against torch itself: values of
    torch.rand and the generator state it ends in, from block boundaries and from mid-block, chained"""
    import torch
    import mt19937_oracle as mt
    for seed in (666, 1, 2 ** 40 + 7):
        torch.manual_seed(seed)
        st = torch.get_rng_state()
        words, left, nxt = mt.parse_state(st.numpy())
        assert words.shape == (624,) and left == 1 and nxt == 0
        for n in (1, 623, 624, 625, 5000, 196608):
            torch.set_rng_state(st)
            ref = torch.rand(n).numpy()
            got, w2, l2, n2 = mt.rand_float32(words, left, nxt, n)
            assert np.array_equal(ref, got)
            assert np.array_equal(mt.store_state(st.numpy(), w2, l2, n2), torch.get_rng_state().numpy())
        torch.set_rng_state(st)
        torch.rand(777)
        mid = torch.get_rng_state()
        ref = torch.rand(98304, 1, 2).numpy().reshape(-1)
        w, l, x = mt.parse_state(mid.numpy())
        got, w2, l2, n2 = mt.rand_float32(w, l, x, 2 * 98304)
        assert np.array_equal(ref, got) and np.array_equal(mt.store_state(mid.numpy(), w2, l2, n2), torch.get_rng_state().numpy())
