"""Deterministic synthetic stand-ins for the paper's (non-public) scenes -- SURVEY.md section 8(d).

"Closed room + clutter", metres, y-up like the reference scenes: an inward-facing box
[-4,4] x [-1.5,1.5] x [-3,3] whose faces are tessellated grids with +-0.02 m value-noise
displacement (about 60 % of the triangles) plus tessellated ellipsoids inside it (about 40 %), an
exact triangle count, a light-map style UV atlas (triangle pair k owns cell k of a square grid, with
a half-texel gutter; float64 [3T,2] like Open3D's ``triangle_uvs``) and an HDR radiance texture
(log-normal base, 0.5 % emitter texels, times 2**hdr_exposure).

Pure numpy; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ROOM_LO = np.array([-4.0, -1.5, -3.0])
ROOM_HI = np.array([4.0, 1.5, 3.0])


@dataclass
class SynthScene:
    vertices: np.ndarray       # [V,3] float32
    triangles: np.ndarray      # [T,3] int32
    triangle_uvs: np.ndarray   # [3T,2] float64
    tri_normals: np.ndarray    # [T,3] float32 unit geometric normals (facing the room's air)
    tri_areas: np.ndarray      # [T] float64


def _value_noise(nu: int, nv: int, rng: np.random.Generator, cells: int = 12) -> np.ndarray:
    """smooth noise in [-1,1] on a (nv+1) x (nu+1) vertex grid (bilinear upsample of a coarse lattice)"""
    lat = rng.uniform(-1.0, 1.0, size=(cells + 1, cells + 1))
    y = np.linspace(0.0, cells, nv + 1)
    x = np.linspace(0.0, cells, nu + 1)
    y0 = np.minimum(y.astype(np.int64), cells - 1)
    x0 = np.minimum(x.astype(np.int64), cells - 1)
    fy = (y - y0)[:, None]
    fx = (x - x0)[None, :]
    a = lat[y0][:, x0]
    b = lat[y0][:, x0 + 1]
    c = lat[y0 + 1][:, x0]
    d = lat[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def _grid_face(origin, du, dv, nu, nv, normal, rng, amp):
    """a tessellated rectangle origin + s*du + t*dv, displaced along `normal` (border kept flat so the
    six faces stay sealed); returns vertices [.,3] and triangles [.,3] wound so that the geometric
    normal (v1-v0) x (v2-v0) points along `normal`"""
    s = np.linspace(0.0, 1.0, nu + 1)
    t = np.linspace(0.0, 1.0, nv + 1)
    S, Tt = np.meshgrid(s, t)
    P = origin[None, None, :] + S[..., None] * du[None, None, :] + Tt[..., None] * dv[None, None, :]
    disp = _value_noise(nu, nv, rng) * amp
    disp[0, :] = disp[-1, :] = 0.0
    disp[:, 0] = disp[:, -1] = 0.0
    P = P + disp[..., None] * normal[None, None, :]
    verts = P.reshape(-1, 3)
    j, i = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    v00 = (j * (nu + 1) + i).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (nu + 1)
    v11 = v01 + 1
    if np.dot(np.cross(du, dv), normal) > 0:
        t0 = np.stack([v00, v10, v11], axis=1)
        t1 = np.stack([v00, v11, v01], axis=1)
    else:
        t0 = np.stack([v00, v11, v10], axis=1)
        t1 = np.stack([v00, v01, v11], axis=1)
    tris = np.empty((2 * v00.shape[0], 3), np.int64)
    tris[0::2] = t0
    tris[1::2] = t1
    return verts, tris


def _cube_sphere(n: int, radii, centre):
    """ellipsoid from a subdivided cube (12 n^2 triangles, outward facing)"""
    verts, tris = [], []
    base = 0
    axes = [(0, 1, 2, 1.0), (0, 1, 2, -1.0), (1, 2, 0, 1.0), (1, 2, 0, -1.0), (2, 0, 1, 1.0), (2, 0, 1, -1.0)]
    lin = np.linspace(-1.0, 1.0, n + 1)
    A, B = np.meshgrid(lin, lin)
    for ua, va, wa, sgn in axes:
        p = np.zeros((n + 1, n + 1, 3))
        p[..., ua] = A
        p[..., va] = B
        p[..., wa] = sgn
        p = p / np.linalg.norm(p, axis=-1, keepdims=True)
        verts.append(p.reshape(-1, 3))
        j, i = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        v00 = (j * (n + 1) + i).reshape(-1) + base
        v10, v01 = v00 + 1, v00 + (n + 1)
        v11 = v01 + 1
        # (e_ua x e_va) . e_wa = +1 for the cyclic triples used above, so outward iff sgn > 0
        if sgn > 0:
            t0 = np.stack([v00, v10, v11], axis=1)
            t1 = np.stack([v00, v11, v01], axis=1)
        else:
            t0 = np.stack([v00, v11, v10], axis=1)
            t1 = np.stack([v00, v01, v11], axis=1)
        tt = np.empty((2 * v00.shape[0], 3), np.int64)
        tt[0::2] = t0
        tt[1::2] = t1
        tris.append(tt)
        base += (n + 1) * (n + 1)
    v = np.concatenate(verts) * np.asarray(radii)[None, :] + np.asarray(centre)[None, :]
    return v, np.concatenate(tris)


def make_uv_atlas(num_tris: int, tex_res: int) -> np.ndarray:
    """triangle pair k -> cell k of a ceil(sqrt(T/2))^2 grid, half-texel gutter; float64 [3T,2]"""
    pairs = (num_tris + 1) // 2
    side = int(np.ceil(np.sqrt(pairs)))
    k = np.arange(num_tris) // 2
    cx = (k % side).astype(np.float64)
    cy = (k // side).astype(np.float64)
    g = 0.5 / float(tex_res)
    u0, u1 = cx / side + g, (cx + 1) / side - g
    v0, v1 = cy / side + g, (cy + 1) / side - g
    odd = (np.arange(num_tris) % 2) == 1
    uv = np.empty((num_tris, 3, 2), np.float64)
    uv[:, 0, 0], uv[:, 0, 1] = u0, v0
    uv[:, 1, 0] = np.where(odd, u1, u1)
    uv[:, 1, 1] = np.where(odd, v1, v0)
    uv[:, 2, 0] = np.where(odd, u0, u1)
    uv[:, 2, 1] = v1
    return uv.reshape(-1, 2)


def make_planar_uvs(vertices: np.ndarray, triangles: np.ndarray) -> np.ndarray:
    """a continuous (position-based) UV map: no seams between neighbouring triangles"""
    p = vertices[triangles.reshape(-1)].astype(np.float64)
    ext = ROOM_HI - ROOM_LO
    q = (p - ROOM_LO[None, :]) / ext[None, :]
    u = 0.08 + 0.84 * (0.8 * q[:, 0] + 0.2 * q[:, 1])
    v = 0.08 + 0.84 * (0.75 * q[:, 2] + 0.25 * q[:, 1])
    return np.stack([u, v], axis=1)


def make_room_scene(num_tris: int, seed: int = 666, clutter: int = 64, tex_res: int = 4096,
                    uv_kind: str = "atlas") -> SynthScene:
    if num_tris < 64:
        raise ValueError("num_tris must be >= 64")
    rng = np.random.default_rng(seed)
    clutter = int(min(clutter, max(0, num_tris // 400)))
    # clutter: ~40 % of the budget in `clutter` ellipsoids of 12 n^2 triangles
    n_sph = 0
    if clutter > 0:
        n_sph = max(2, int(np.sqrt(0.4 * num_tris / clutter / 12.0)))
        while clutter * 12 * n_sph * n_sph > 0.45 * num_tris and n_sph > 2:
            n_sph -= 1
    tri_clutter = clutter * 12 * n_sph * n_sph
    quads = max(6, (num_tris - tri_clutter) // 2)
    ext = ROOM_HI - ROOM_LO
    faces = [  # origin, du, dv, inward normal
        (np.array([ROOM_LO[0], ROOM_LO[1], ROOM_LO[2]]), np.array([ext[0], 0, 0]), np.array([0, 0, ext[2]]), np.array([0.0, 1, 0])),
        (np.array([ROOM_LO[0], ROOM_HI[1], ROOM_LO[2]]), np.array([ext[0], 0, 0]), np.array([0, 0, ext[2]]), np.array([0.0, -1, 0])),
        (np.array([ROOM_LO[0], ROOM_LO[1], ROOM_LO[2]]), np.array([ext[0], 0, 0]), np.array([0, ext[1], 0]), np.array([0.0, 0, 1])),
        (np.array([ROOM_LO[0], ROOM_LO[1], ROOM_HI[2]]), np.array([ext[0], 0, 0]), np.array([0, ext[1], 0]), np.array([0.0, 0, -1])),
        (np.array([ROOM_LO[0], ROOM_LO[1], ROOM_LO[2]]), np.array([0, 0, ext[2]]), np.array([0, ext[1], 0]), np.array([1.0, 0, 0])),
        (np.array([ROOM_HI[0], ROOM_LO[1], ROOM_LO[2]]), np.array([0, 0, ext[2]]), np.array([0, ext[1], 0]), np.array([-1.0, 0, 0])),
    ]
    areas = np.array([np.linalg.norm(np.cross(f[1], f[2])) for f in faces])
    rho = quads / areas.sum()
    while True:
        dims = []
        for f in faces:
            a, b = np.linalg.norm(f[1]), np.linalg.norm(f[2])
            dims.append((max(1, int(a * np.sqrt(rho))), max(1, int(b * np.sqrt(rho)))))
        if sum(2 * nu * nv for nu, nv in dims) + tri_clutter <= num_tris:
            break
        rho *= 0.98
    verts, tris = [], []
    base = 0
    for f, (nu, nv) in zip(faces, dims):
        v, t = _grid_face(f[0], f[1], f[2], nu, nv, f[3], rng, 0.02)
        verts.append(v)
        tris.append(t + base)
        base += v.shape[0]
    for _ in range(clutter):
        radii = rng.uniform(0.15, 0.8, size=3)
        centre = rng.uniform(ROOM_LO + radii + 0.05, ROOM_HI - radii - 0.05)
        v, t = _cube_sphere(n_sph, radii, centre)
        verts.append(v)
        tris.append(t + base)
        base += v.shape[0]
    V = np.concatenate(verts)
    T = np.concatenate(tris)
    # exact count: bisect the longest edge of randomly chosen triangles (one extra triangle each)
    while T.shape[0] < num_tris:
        need = num_tris - T.shape[0]
        pick = rng.choice(T.shape[0], size=min(need, T.shape[0]), replace=False)
        tri = T[pick]
        p = V[tri]
        elen = np.stack([np.linalg.norm(p[:, 1] - p[:, 0], axis=1), np.linalg.norm(p[:, 2] - p[:, 1], axis=1),
                         np.linalg.norm(p[:, 0] - p[:, 2], axis=1)], axis=1)
        e = np.argmax(elen, axis=1)
        rows = np.arange(tri.shape[0])
        a, b, c = tri[rows, e], tri[rows, (e + 1) % 3], tri[rows, (e + 2) % 3]
        mid = 0.5 * (V[a] + V[b])
        m = V.shape[0] + rows
        V = np.concatenate([V, mid])
        T[pick] = np.stack([a, m, c], axis=1)
        T = np.concatenate([T, np.stack([m, b, c], axis=1)])
    V32 = np.ascontiguousarray(V, dtype=np.float32)
    T32 = np.ascontiguousarray(T, dtype=np.int32)
    p = V32[T32].astype(np.float64)
    cr = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    nrm = np.linalg.norm(cr, axis=1)
    tri_normals = (cr / np.maximum(nrm, 1e-30)[:, None]).astype(np.float32)
    uvs = make_uv_atlas(T32.shape[0], tex_res) if uv_kind == "atlas" else make_planar_uvs(V32, T32)
    return SynthScene(V32, T32, uvs, tri_normals, 0.5 * nrm)


def make_radiance_texture(res: int, seed: int = 666, hdr_exposure: float = 5.0, kind: str = "noise") -> np.ndarray:
    """[1,3,res,res] float32, already 'flipped and exposed' like self.texture (mat_nvdiffrast.py:89-93).
    kind='noise': i.i.d. log-normal texels + 0.5 % emitters; kind='smooth': band-limited."""
    rng = np.random.default_rng(seed + 1)
    if kind == "noise":
        tex = np.exp(rng.normal(-1.0, 1.0, size=(res, res, 3))).astype(np.float32)
        emit = rng.random((res, res)) < 0.005
        gain = rng.uniform(2.0 ** 5, 2.0 ** 8, size=(res, res)).astype(np.float32)
        tex[emit] *= gain[emit][:, None]
    elif kind == "smooth":
        y, x = np.meshgrid(np.linspace(0, 1, res), np.linspace(0, 1, res), indexing="ij")
        tex = np.empty((res, res, 3), np.float32)
        for c in range(3):
            ph = rng.uniform(0, 2 * np.pi, size=4)
            f = rng.uniform(1.0, 5.0, size=4)
            tex[..., c] = np.exp(-1.0 + 0.8 * np.sin(2 * np.pi * f[0] * x + ph[0]) * np.cos(2 * np.pi * f[1] * y + ph[1])
                                 + 0.5 * np.sin(2 * np.pi * (f[2] * x + f[3] * y) + ph[2]))
        gx, gy = rng.uniform(0.2, 0.8, size=2)
        tex *= (1.0 + 60.0 * np.exp(-((x - gx) ** 2 + (y - gy) ** 2) / 0.002))[..., None].astype(np.float32)
    else:
        raise ValueError(kind)
    tex *= np.float32(2.0 ** hdr_exposure)
    return np.ascontiguousarray(tex.transpose(2, 0, 1)[None])


def sample_surface_points(scene: SynthScene, num_points: int, seed: int = 666, seam_fraction: float = 0.005):
    """IrT texels: area-weighted surface samples, pos = x + 1e-2 n (tracer_o3d_irt.py:110), a few
    all-zero 'seam' texels (:138-139).  Returns pos, nrm float32 [P,3]."""
    rng = np.random.default_rng(seed + 2)
    cdf = np.cumsum(scene.tri_areas)
    t = np.searchsorted(cdf, rng.random(num_points) * cdf[-1]).clip(0, scene.triangles.shape[0] - 1)
    r1, r2 = np.sqrt(rng.random(num_points)), rng.random(num_points)
    b0, b1, b2 = 1 - r1, r1 * (1 - r2), r1 * r2
    p = scene.vertices[scene.triangles[t]].astype(np.float64)
    x = p[:, 0] * b0[:, None] + p[:, 1] * b1[:, None] + p[:, 2] * b2[:, None]
    n = scene.tri_normals[t].astype(np.float64)
    pos = (x + 1e-2 * n).astype(np.float32)
    nrm = n.astype(np.float32)
    seam = rng.random(num_points) < seam_fraction
    pos[seam] = 0.0
    nrm[seam] = 0.0
    return np.ascontiguousarray(pos), np.ascontiguousarray(nrm)


def cube_map_rays(res: int, cam=(0.0, 0.0, 0.0)) -> np.ndarray:
    """[6*res*res, 6] primary rays through the pixel centres of a cube map (the six 90-degree views the
    reference rasterises, datasets/dataset.py:402-469)"""
    lin = (np.arange(res) + 0.5) / res * 2.0 - 1.0
    a, b = np.meshgrid(lin, lin)
    one = np.ones_like(a)
    dirs = [np.stack([one, -b, -a], -1), np.stack([-one, -b, a], -1), np.stack([a, one, b], -1),
            np.stack([a, -one, -b], -1), np.stack([a, -b, one], -1), np.stack([-a, -b, -one], -1)]
    d = np.concatenate([x.reshape(-1, 3) for x in dirs])
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(np.asarray(cam, np.float64)[None, :], d.shape)
    return np.ascontiguousarray(np.concatenate([o, d], axis=1), dtype=np.float32)


def cube_map_ray_differentials(res: int) -> np.ndarray:
    """[6*res*res, 6] = (d dir/dX, d dir/dY) of cube_map_rays' normalised directions per pixel step (X = column,
    Y = row): what texir_gbuffer_uv_da turns into the texd of dr.interpolate(..., diff_attrs='all')"""
    lin = (np.arange(res) + 0.5) / res * 2.0 - 1.0
    a, b = np.meshgrid(lin, lin)
    one, zero = np.ones_like(a), np.zeros_like(a)
    step = 2.0 / res
    dirs = [np.stack([one, -b, -a], -1), np.stack([-one, -b, a], -1), np.stack([a, one, b], -1),
            np.stack([a, -one, -b], -1), np.stack([a, -b, one], -1), np.stack([-a, -b, -one], -1)]
    # derivatives of the un-normalised directions with respect to a (columns) and b (rows)
    d_da = [(0, 0, -1), (0, 0, 1), (1, 0, 0), (1, 0, 0), (1, 0, 0), (-1, 0, 0)]
    d_db = [(0, -1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (0, -1, 0), (0, -1, 0)]
    out = []
    for face in range(6):
        d = dirs[face].reshape(-1, 3)
        ln = np.linalg.norm(d, axis=1, keepdims=True)
        dh = d / ln
        cols = []
        for dv in (d_da[face], d_db[face]):
            g = np.broadcast_to(np.asarray(dv, np.float64) * step, d.shape)
            cols.append((g - dh * np.sum(dh * g, axis=1, keepdims=True)) / ln)
        out.append(np.concatenate(cols, axis=1))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)


def gbuffer_from_hits(scene: SynthScene, rays: np.ndarray, t_hit: np.ndarray, prim_ids: np.ndarray):
    """pos = hit + 1e-2 n, geometric normal; uncovered pixels get pos = n = (1,0,0)
    (mat_nvdiffrast.py:125, :179)"""
    hit = np.isfinite(t_hit) & (prim_ids != 0xFFFFFFFF)
    ids = np.where(hit, prim_ids, 0).astype(np.int64)
    n = scene.tri_normals[ids]
    x = rays[:, 0:3] + rays[:, 3:6] * np.where(hit, t_hit, 0.0)[:, None].astype(np.float32)
    pos = (x + np.float32(1e-2) * n).astype(np.float32)
    nrm = n.astype(np.float32).copy()
    pos[~hit] = np.array([1.0, 0.0, 0.0], np.float32)
    nrm[~hit] = np.array([1.0, 0.0, 0.0], np.float32)
    return np.ascontiguousarray(pos), np.ascontiguousarray(nrm)


def make_materials(num_points: int, seed: int = 666):
    """roughness ~ U(0.01, 0.8) (trainer clamp range, train_material.py:458), albedo ~ U(0,1),
    irr ~ exp(N(0,1)), upstream gradient ~ N(0,1)"""
    rng = np.random.default_rng(seed + 3)
    rough = rng.uniform(0.01, 0.8, size=(num_points, 1)).astype(np.float32)
    albedo = rng.uniform(0.0, 1.0, size=(num_points, 3)).astype(np.float32)
    irr = np.exp(rng.normal(0.0, 1.0, size=(num_points, 3))).astype(np.float32)
    g_rgb = rng.normal(0.0, 1.0, size=(num_points, 3)).astype(np.float32)
    return rough, albedo, irr, g_rgb
