"""CPU: the kernels' host/device headers (texir_math.h, texir_bvh.h, texir_point.h) and the host BVH builder,
compiled with g++ (tests/cpu_emul), against the oracle and the reference goldens.  This catches arithmetic bugs
in the device code on a box without a GPU; the GPU tests (test_gpu_*.py) then check the real kernels."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, rel_l2


class EmulScene:
    def __init__(self, lib, verts, tris, uvs=None, tex=None):
        self.lib = lib
        self.v = np.ascontiguousarray(verts, np.float32)
        self.t = np.ascontiguousarray(tris, np.int32)
        self.uv = None if uvs is None else np.ascontiguousarray(uvs, np.float64)
        self.h = lib.emul_scene_create(self.v.ctypes.data, len(self.v), self.t.ctypes.data, len(self.t),
                                       None if self.uv is None else self.uv.ctypes.data, 2)
        assert self.h
        if tex is not None:
            self.tex = np.ascontiguousarray(tex, np.float32).reshape(3, tex.shape[-2], tex.shape[-1])
            lib.emul_set_texture(self.h, self.tex.ctypes.data, self.tex.shape[1], self.tex.shape[2])

    def __del__(self):
        self.lib.emul_scene_destroy(self.h)

    def stats(self):
        out = np.zeros(7)
        self.lib.emul_scene_stats(self.h, out.ctypes.data)
        return out

    def cast(self, kind, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        R = len(rays)
        t, i, uv = np.empty(R, np.float32), np.empty(R, np.uint32), np.empty((R, 2), np.float32)
        nv, nt = np.empty(R, np.uint32), np.empty(R, np.uint32)
        self.lib.emul_cast_rays(self.h, kind, rays.ctypes.data, R, t.ctypes.data, i.ctypes.data, uv.ctypes.data,
                                nv.ctypes.data, nt.ctypes.data)
        return t, i, uv, nv, nt


def _rays(R, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform([-3.5, -1.2, -2.5], [3.5, 1.2, 2.5], size=(R, 3)).astype(np.float32)
    d = rng.normal(size=(R, 3)).astype(np.float32)
    d[:4] = 0.0
    d[4:8] = [[1, 0, 0], [0, -1, 0], [0, 0, 1], [-1, 0, 0]]  # axis-aligned: zero components in the slab test
    return np.concatenate([o, d], 1)


@pytest.mark.parametrize("kind", [2, 8])
def test_cast_rays_bit_exact_vs_oracle(emul, oracle, kind):
    g = load_golden("scene_atlas")
    rays = _rays(6000, 5)
    t0, i0, u0 = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays)
    t, i, uv, nv, nt = EmulScene(emul, g["vertices"], g["triangles"]).cast(kind, rays)
    assert np.array_equal(i, i0) and np.array_equal(t, t0) and np.array_equal(uv, u0)
    assert nv.mean() < 60 and nt.mean() < 40  # a traversal, not a scan


def test_bvh_build_stats(emul):
    from texir_code_b200 import synth
    sc = synth.make_room_scene(20000, clutter=8, tex_res=256)
    es = EmulScene(emul, sc.vertices, sc.triangles)
    n2, n8, d2, d8, sah2, sah8, secs = es.stats()
    assert n2 > 20000 / 3 - 1 and n8 < n2 / 2 and d8 < d2 <= 60


@pytest.mark.parametrize("N", [16, 64])
@pytest.mark.parametrize("mode", ["uniform", "cosine", "importance"])
def test_generate_dir_vs_reference(emul, N, mode):
    g = load_golden("generate_dir")
    nrm, rough, shift = g["normals"], np.ascontiguousarray(g["roughness"].reshape(-1)), np.ascontiguousarray(g["shift"].reshape(-1, 2))
    out = np.empty((len(nrm), N, 3), np.float32)
    emul.emul_generate_dir(nrm.ctypes.data, rough.ctypes.data, shift.ctypes.data, len(nrm), N,
                           {"uniform": 0, "cosine": 1, "importance": 2}[mode], out.ctypes.data)
    want = g["L_%s_%d" % (mode, N)]
    # identical up to libm sin/cos ulps.  importance mode: torch's vectorised CPU sqrt is not correctly rounded
    # (0.6 % of inputs differ from IEEE sqrt by 1 ulp; CUDA's sqrtf -- what the reference really runs -- is IEEE
    # like ours), and a 1-ulp change of cos(theta) near 1 is amplified in sin(theta) = sqrt(1 - cos^2)
    assert np.abs(out - want).max() <= (2e-6 if mode == "importance" else 2.5e-7)
    assert (out == want).mean() > 0.9
    assert np.all(out[3] == 0)


@pytest.mark.parametrize("kind", [2, 8])
def test_query_irf_vs_reference(emul, scene_golden, kind):
    name, g = scene_golden
    es = EmulScene(emul, g["vertices"], g["triangles"], g["triangle_uvs"], g["texture"])
    for k in ("render_e14_16_", "render_e6_64_"):
        l, want = np.ascontiguousarray(g[k + "l"]), g[k + "lighting"]
        P, N = l.shape[:2]
        rad = np.empty((P, N, 3), np.float32)
        emul.emul_query_irf(es.h, kind, g["pos"].ctypes.data, l.ctypes.data, P, N, rad.ctypes.data)
        assert np.array_equal(rad, want)
    q = load_golden("query_irf_quad")
    eq = EmulScene(emul, q["vertices"], q["triangles"], q["triangle_uvs"], q["texture"])
    rad = np.empty((6, 1, 3), np.float32)
    emul.emul_query_irf(eq.h, kind, q["org"].ctypes.data, q["dir"].ctypes.data, 6, 1, rad.ctypes.data)
    assert np.array_equal(rad, q["rad"])


@pytest.mark.parametrize("kind", [2, 8])
def test_irradiance_vs_reference(emul, scene_golden, kind):
    name, g = scene_golden
    es = EmulScene(emul, g["vertices"], g["triangles"], g["triangle_uvs"], g["texture"])
    P, N = len(g["irt_pos"]), int(g["irt_N"])
    sh = np.ascontiguousarray(g["irt_shift"].reshape(P, 2))
    E = np.empty((P, 3), np.float32)
    emul.emul_irradiance(es.h, kind, g["irt_pos"].ctypes.data, g["irt_nrm"].ctypes.data, sh.ctypes.data, P, N, 0,
                         E.ctypes.data)
    # the atlas scene has a discontinuous texture: a 1-ulp sin/cos difference can move a hit to another
    # atlas cell, so its bound is the north-star's 1e-4; the smooth scene must be much tighter
    assert rel_l2(E, g["irt_E"]) < (1e-4 if name == "atlas" else 2e-6)
    seam = np.all(g["irt_nrm"] == 0, axis=1)
    assert np.all(E[seam] == 0)


@pytest.mark.parametrize("kind", [2, 8])
@pytest.mark.parametrize("tag,N,eps", [("e14", 16, 1e-14), ("e6", 64, 1e-6)])
def test_shade_vs_reference(emul, scene_golden, kind, tag, N, eps):
    name, g = scene_golden
    es = EmulScene(emul, g["vertices"], g["triangles"], g["triangle_uvs"], g["texture"])
    k = "render_%s_%d_" % (tag, N)
    P = len(g["pos"])
    rough = np.ascontiguousarray(g["rough"].reshape(-1))
    sh = np.ascontiguousarray(g[k + "shift"].reshape(P, 2))
    rgb, drgb, ediff = (np.empty((P, 3), np.float32) for _ in range(3))
    emul.emul_shade(es.h, kind, g["nrm"].ctypes.data, g["albedo"].ctypes.data, rough.ctypes.data, g["pos"].ctypes.data,
                    g["cam"].ctypes.data, g["irr"].ctypes.data, sh.ctypes.data, None, P, N, 0, 2, 0,
                    ctypes.c_float(eps), rgb.ctypes.data, drgb.ctypes.data, ediff.ctypes.data)
    assert rel_l2(rgb, g[k + "rgb"]) < 1e-4
    g_rough = (g["g_rgb"] * drgb).sum(1)
    assert rel_l2(g_rough, g[k + "g_rough"]) < 1e-3
    assert rel_l2(g["g_rgb"] * g["irr"] / np.float32(np.pi), g[k + "g_albedo"]) < 1e-6


@pytest.mark.parametrize("dmode", ["uniform", "cosine"])
def test_relighting_vs_reference(emul, scene_golden, dmode):
    name, g = scene_golden
    es = EmulScene(emul, g["vertices"], g["triangles"], g["triangle_uvs"], g["texture"])
    P = len(g["pos"])
    rough = np.ascontiguousarray(g["rough"].reshape(-1))
    shs = np.ascontiguousarray(g["relight_%s_shift_spec" % dmode].reshape(P, 2))
    shd = np.ascontiguousarray(g["relight_%s_shift_diff" % dmode].reshape(P, 2))
    rgb, drgb, ediff = (np.empty((P, 3), np.float32) for _ in range(3))
    emul.emul_shade(es.h, 8, g["nrm"].ctypes.data, g["albedo"].ctypes.data, rough.ctypes.data, g["pos"].ctypes.data,
                    g["cam"].ctypes.data, None, shs.ctypes.data, shd.ctypes.data, P, 16, 32, 2,
                    1 if dmode == "cosine" else 0, ctypes.c_float(1e-6), rgb.ctypes.data, drgb.ctypes.data,
                    ediff.ctypes.data)
    assert rel_l2(rgb, g["relight_%s_rgb" % dmode]) < 1e-4


# ---------------------------------------------------------------------------------------------------
# the walks the kernels use (texir_walk.h): cheaper float node step, 32-byte quantised nodes
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("quantised", [0, 1])
def test_fast_node_step_bit_exact_vs_oracle(emul, oracle, quantised):
    """float nodes with the FMA slab test, and the 32-byte quantised nodes: same hits as brute force"""
    for name in ("scene_atlas", "scene_smooth"):
        g = load_golden(name)
        rays = _rays(20000, 17)
        t0, i0, u0 = oracle.OracleScene(g["vertices"], g["triangles"], brute=True).cast_rays_np(rays)
        es = EmulScene(emul, g["vertices"], g["triangles"])
        t, i, uv = np.empty(len(rays), np.float32), np.empty(len(rays), np.uint32), np.empty((len(rays), 2), np.float32)
        emul.emul_cast_rays_fast(es.h, quantised, rays.ctypes.data, len(rays), t.ctypes.data, i.ctypes.data, uv.ctypes.data)
        assert np.array_equal(i, i0) and np.array_equal(t, t0) and np.array_equal(uv, u0)


def test_quantised_nodes_on_degenerate_scenes(emul, oracle):
    """flat scene (zero extent along z), single triangle, rays inside the plane"""
    q = load_golden("query_irf_quad")
    rng = np.random.default_rng(4)
    rays = np.concatenate([rng.uniform(-2, 2, size=(4000, 3)), rng.normal(size=(4000, 3))], 1).astype(np.float32)
    rays[:50, 5] = 0.0          # travelling parallel to the quad
    rays[50:100, 2] = 2.0       # origins exactly in the quad's plane
    for verts, tris in ((q["vertices"], q["triangles"]), (q["vertices"][:3], q["triangles"][:1])):
        t0, i0, u0 = oracle.OracleScene(verts, tris, brute=True).cast_rays_np(rays)
        es = EmulScene(emul, verts, tris)
        t, i, uv = np.empty(len(rays), np.float32), np.empty(len(rays), np.uint32), np.empty((len(rays), 2), np.float32)
        emul.emul_cast_rays_fast(es.h, 1, rays.ctypes.data, len(rays), t.ctypes.data, i.ctypes.data, uv.ctypes.data)
        assert np.array_equal(i, i0) and np.array_equal(t, t0) and np.array_equal(uv, u0)


def test_shared_first_descent_bit_exact_vs_plain_walk(emul):
    """texir_walk.h: rays that share an origin start from the per-point chain of siblings (qbvh2_point_front /
    trace_qbvh2_front) instead of the root; hits must be bit-identical to the plain QBVH2 walk, for origins on the
    surface (+1e-2 n, the hot path's), in free space, at the scene centre, outside the scene and on seams (0,0,0)"""
    from texir_code_b200 import synth
    rng = np.random.default_rng(0)
    for num_tris, clutter in ((4000, 6), (60000, 16)):
        sc = synth.make_room_scene(num_tris, seed=666, clutter=clutter)
        es = EmulScene(emul, sc.vertices, sc.triangles)
        pos, _ = synth.sample_surface_points(sc, 24, seed=3, seam_fraction=0.1)
        extra = np.array([[0, 0, 0], [100, 0, 0], [3.99, 1.49, 2.99], [0.3, -0.2, 0.5]], np.float32)
        sizes = []
        for p in np.concatenate([pos, extra]).astype(np.float32):
            d = rng.normal(size=(256, 3)).astype(np.float32)
            d[:5] = [[1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, 0], [-1, 0, 0]]
            rays = np.ascontiguousarray(np.concatenate([np.repeat(p[None], 256, 0), d], 1), np.float32)
            t0, i0, u0 = np.empty(256, np.float32), np.empty(256, np.uint32), np.empty((256, 2), np.float32)
            emul.emul_cast_rays_fast(es.h, 1, rays.ctypes.data, 256, t0.ctypes.data, i0.ctypes.data, u0.ctypes.data)
            t1, i1, u1 = np.empty(256, np.float32), np.empty(256, np.uint32), np.empty((256, 2), np.float32)
            info = np.zeros(2, np.int32)
            p = np.ascontiguousarray(p)
            d = np.ascontiguousarray(d)
            emul.emul_cast_rays_front(es.h, p.ctypes.data, d.ctypes.data, 256, t1.ctypes.data, i1.ctypes.data, u1.ctypes.data,
                                      info.ctypes.data)
            assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
            assert np.array_equal(i0, i1) and np.array_equal(u0.view(np.uint32), u1.view(np.uint32))
            sizes.append(int(info[0]))
        assert max(sizes) >= 8 and min(sizes) == 0      # surface points share a long chain; outside points none
    # single-triangle scene: the root's two children are the same leaf
    one = EmulScene(emul, np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.int32))
    p = np.array([0.2, 0.2, 1.0], np.float32)
    d = np.ascontiguousarray(np.array([[0, 0, -1], [0, 0, 1], [0.1, 0.1, -1]], np.float32))
    t1, i1, u1 = np.empty(3, np.float32), np.empty(3, np.uint32), np.empty((3, 2), np.float32)
    emul.emul_cast_rays_front(one.h, p.ctypes.data, d.ctypes.data, 3, t1.ctypes.data, i1.ctypes.data, u1.ctypes.data, None)
    assert i1.tolist() == [0, 0xFFFFFFFF, 0] and t1[0] == 1.0 and np.isinf(t1[1])


def test_free_running_walk_bit_exact_vs_plain_walk(emul):
    """texir_walk.h trace_qbvh2_free (the G-buffer kernel's loop: node step or leaf test per iteration, whichever the ray
    stands at): same visiting order per ray, so hits must be bit-identical to the while-while walk -- primary rays from
    inside the room, rays from outside, degenerate directions, and the single-triangle scene"""
    from texir_code_b200 import synth
    rng = np.random.default_rng(5)
    for num_tris, clutter in ((4000, 6), (60000, 16)):
        sc = synth.make_room_scene(num_tris, seed=666, clutter=clutter)
        es = EmulScene(emul, sc.vertices, sc.triangles)
        rays = np.ascontiguousarray(synth.cube_map_rays(24), np.float32)
        o = rng.uniform(-6, 6, (2000, 3)).astype(np.float32)
        d = rng.normal(size=(2000, 3)).astype(np.float32)
        d[:4] = [[0, 0, 0], [1, 0, 0], [0, np.inf, 0], [np.nan, 0, 1]]
        rays = np.ascontiguousarray(np.concatenate([rays, np.concatenate([o, d], 1)]), np.float32)
        R = len(rays)
        t0, i0, u0 = np.empty(R, np.float32), np.empty(R, np.uint32), np.empty((R, 2), np.float32)
        emul.emul_cast_rays_fast(es.h, 1, rays.ctypes.data, R, t0.ctypes.data, i0.ctypes.data, u0.ctypes.data)
        t1, i1, u1 = np.empty(R, np.float32), np.empty(R, np.uint32), np.empty((R, 2), np.float32)
        emul.emul_cast_rays_free(es.h, rays.ctypes.data, R, t1.ctypes.data, i1.ctypes.data, u1.ctypes.data)
        assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
        assert np.array_equal(i0, i1) and np.array_equal(u0.view(np.uint32), u1.view(np.uint32))
        assert (i0 != 0xFFFFFFFF).mean() > 0.5
    one = EmulScene(emul, np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.int32))
    rays = np.ascontiguousarray(np.array([[0.2, 0.2, 1.0, 0, 0, -1], [0.2, 0.2, 1.0, 0, 0, 1]], np.float32))
    t1, i1, u1 = np.empty(2, np.float32), np.empty(2, np.uint32), np.empty((2, 2), np.float32)
    emul.emul_cast_rays_free(one.h, rays.ctypes.data, 2, t1.ctypes.data, i1.ctypes.data, u1.ctypes.data)
    assert i1.tolist() == [0, 0xFFFFFFFF] and t1[0] == 1.0 and np.isinf(t1[1])


def test_entry_distance_stack_walk_bit_exact_vs_plain_walk(emul):
    """texir_walk.h trace_qbvh2_t (the render kernels' walk: stack entries carry the postponed child's entry distance and a
    pop drops what the closest hit has overtaken): dropping such entries must not change a single hit -- with and without
    the shared first descent, rays from surface points, from outside, degenerate directions"""
    from texir_code_b200 import synth
    rng = np.random.default_rng(9)
    for num_tris, clutter in ((4000, 6), (60000, 16)):
        sc = synth.make_room_scene(num_tris, seed=666, clutter=clutter)
        es = EmulScene(emul, sc.vertices, sc.triangles)
        pos, _ = synth.sample_surface_points(sc, 12, seed=5, seam_fraction=0.0)
        for p in np.concatenate([pos, np.array([[0, 0, 0], [50, 0, 0]], np.float32)]).astype(np.float32):
            d = rng.normal(size=(512, 3)).astype(np.float32)
            d[:4] = [[0, 0, 0], [1, 0, 0], [0, np.inf, 0], [np.nan, 0, 1]]
            rays = np.ascontiguousarray(np.concatenate([np.repeat(p[None], 512, 0), d], 1), np.float32)
            t0, i0, u0 = np.empty(512, np.float32), np.empty(512, np.uint32), np.empty((512, 2), np.float32)
            emul.emul_cast_rays_fast(es.h, 1, rays.ctypes.data, 512, t0.ctypes.data, i0.ctypes.data, u0.ctypes.data)
            for use_front in (0, 1):
                t1, i1, u1 = np.empty(512, np.float32), np.empty(512, np.uint32), np.empty((512, 2), np.float32)
                emul.emul_cast_rays_tnear(es.h, rays.ctypes.data, 512, use_front, t1.ctypes.data, i1.ctypes.data, u1.ctypes.data)
                assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32)), use_front
                assert np.array_equal(i0, i1) and np.array_equal(u0.view(np.uint32), u1.view(np.uint32))
    one = EmulScene(emul, np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.int32))
    rays = np.ascontiguousarray(np.array([[0.2, 0.2, 1.0, 0, 0, -1], [0.2, 0.2, 1.0, 0, 0, 1]], np.float32))
    t1, i1, u1 = np.empty(2, np.float32), np.empty(2, np.uint32), np.empty((2, 2), np.float32)
    emul.emul_cast_rays_tnear(one.h, rays.ctypes.data, 2, 0, t1.ctypes.data, i1.ctypes.data, u1.ctypes.data)
    assert i1.tolist() == [0, 0xFFFFFFFF] and t1[0] == 1.0 and np.isinf(t1[1])
