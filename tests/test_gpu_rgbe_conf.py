"""GPU (-m gpu): the RGBE-packed radiance texture (texir_scene_set_texture_rgbe) and the config-level model classes.

The packed texture must reproduce the float path BIT FOR BIT: the reference's radiance texture always comes out of a
Radiance .hdr file (models/mat_nvdiffrast.py:89-93), i.e. 8-bit mantissas with a shared exponent, times a power of two."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tx(cuda_device):
    import texir_code_b200 as tx
    tx._lib.load()
    return tx


def _float_texture(tx, rgbe, exposure):
    """what the reference builds from the file: cv2.imread (BGR) -> RGB -> flip rows -> * 2**exposure, [1,3,H,W]"""
    from texir_code_b200 import postprocess
    bgr = postprocess.rgbe_decode(rgbe.cuda(), bgr=True)
    return (bgr.flip(-1).flip(0).permute(2, 0, 1).unsqueeze(0) * (2.0 ** exposure)).contiguous()


@pytest.mark.parametrize("exposure", [5, 0, -3])
def test_rgbe_texture_is_bit_identical_to_the_float_texture(tx, cuda_device, scene_golden, exposure):
    name, g = scene_golden
    rng = np.random.default_rng(4)
    H = W = 128
    rgbe = rng.integers(0, 256, (H, W, 4)).astype(np.uint8)
    rgbe[..., 3] = rng.integers(100, 150, (H, W))                  # plausible radiance exponents
    rgbe[rng.random((H, W)) < 0.05] = 0                            # black texels (exponent byte 0)
    rgbe[0, :8, 3] = np.arange(1, 9)                               # exponents whose scale is a float32 denormal
    rgbe_t = torch.from_numpy(rgbe)
    tex = _float_texture(tx, rgbe_t, exposure)
    sc = tx.Scene(g["vertices"], g["triangles"], g["triangle_uvs"], device=cuda_device)
    pos, nrm = torch.from_numpy(g["pos"]).to(cuda_device), torch.from_numpy(g["nrm"]).to(cuda_device)
    torch.manual_seed(1)
    shift = torch.rand(pos.shape[0], 1, 2).to(cuda_device)
    sc.set_texture(tex)
    E_f = tx.irradiance(sc, pos, nrm, shift, 256).clone()
    d = tx.generate_dir(nrm, 64, None, "uniform", shift=shift)
    rad_f = tx.query_irf(sc, pos.unsqueeze(1).expand_as(d), d.unsqueeze(-2), 64).clone()
    bytes_f = sc.info()["device_bytes"]
    sc.set_texture_rgbe(rgbe_t, exposure, flip=True)
    E_p = tx.irradiance(sc, pos, nrm, shift, 256)
    rad_p = tx.query_irf(sc, pos.unsqueeze(1).expand_as(d), d.unsqueeze(-2), 64)
    assert torch.equal(rad_p, rad_f) and torch.equal(E_p, E_f)
    assert float(rad_f.abs().max()) > 0
    sc.set_texture(tex)                                             # the float texture takes over again
    assert torch.equal(tx.irradiance(sc, pos, nrm, shift, 256), E_f)
    with pytest.raises(ValueError):
        sc.set_texture_rgbe(rgbe_t, 2.5)


def _write_scene_files(tmp_path, sc, rgbe_tex, rgbe_irt, index_png=None):
    import cv2
    from texir_code_b200 import postprocess
    d = tmp_path / "hdr_texture"
    d.mkdir()
    obj = d / "out1.obj"
    uv = sc.triangle_uvs.reshape(-1, 2)
    nrm = sc.tri_normals
    with open(obj, "w") as f:
        for v in sc.vertices:
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        for t in uv:
            f.write("vt %.17g %.17g\n" % tuple(t))
        for t in range(len(sc.triangles)):
            a, b, c = (int(x) + 1 for x in sc.triangles[t])
            f.write("f %d/%d %d/%d %d/%d\n" % (a, 3 * t + 1, b, 3 * t + 2, c, 3 * t + 3))
    for name, rgbe in (("hdr_texture.hdr", rgbe_tex), ("irt.hdr", rgbe_irt)):
        img = postprocess.rgbe_decode(torch.from_numpy(rgbe).cuda(), bgr=True)
        postprocess.imwrite_hdr(str(d / name), img)
    if index_png is not None:
        cv2.imwrite(str(d / "0.png"), index_png)
    return str(obj)


def test_conf_driven_material_model_equals_the_array_built_one(tx, cuda_device, tmp_path):
    """texir_code_b200.conf_models.MaterialModel(conf, ids, extrinsics, optim_cam) -- the constructor the reference's
    trainer calls (trainer/train_material.py:110-111) -- reading out1.obj / hdr_texture.hdr / irt.hdr from disk,
    against models.MaterialModel built from the same arrays"""
    from test_host_logic import _reference_cube_mvps
    from texir_code_b200 import conf_models, models, synth
    sc = synth.make_room_scene(5000, seed=7, clutter=4, tex_res=64, uv_kind="planar")
    rng = np.random.default_rng(9)
    mk = lambda n: np.concatenate([rng.integers(1, 256, (n, n, 3)), rng.integers(120, 132, (n, n, 1))], -1).astype(np.uint8)
    rgbe_tex, rgbe_irt = mk(64), mk(32)
    obj = _write_scene_files(tmp_path, sc, rgbe_tex, rgbe_irt)
    conf = conf_models.ConfigLite.parse_string("""
        train{ path_mesh_open3d = %s
               pano_img_res = [32, 64]
               sample_light = [32, 16]
               hdr_exposure = 5 }
        models{ render{ sample_type = [uniform, importance] } }
    """ % obj)
    m = conf_models.MaterialModel(conf=conf, ids=[0], extrinsics=None, optim_cam=False)
    assert m.cube_res == 16 and m.sample_l == [32, 16] and m.max_mip_level == 13
    assert tuple(m.materials_a.shape) == (2048, 2048, 3) and tuple(m.materials_r.shape) == (4096, 4096, 1)
    assert sorted(k for k in m.state_dict()) == ["irrt", "materials_a", "materials_r"]      # no vn in this OBJ
    tex = _float_texture(tx, torch.from_numpy(rgbe_tex), 5)
    assert torch.equal(m.texture, tex)
    from texir_code_b200 import postprocess
    irt = postprocess.rgbe_decode(torch.from_numpy(rgbe_irt).cuda(), bgr=True).flip(-1)       # NOT flipped vertically (:73)
    assert torch.equal(m.irrt, irt)
    ref = models.MaterialModel(sc.vertices, sc.triangles, sc.triangle_uvs, tex, sample_l=[32, 16], device=cuda_device,
                               materials_a_res=2048, materials_r_res=4096, cube_res=16, irrt=irt, max_mip_level=13)
    m._warned_normals = ref._warned_normals = True
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 3] = [0.3, 0.0, -0.2]
    mvp, cam = _reference_cube_mvps(c2w)
    outs = []
    for model in (m, ref):
        torch.manual_seed(3)
        out = model(mvp.to(cuda_device), 0, cam, 2)
        out["rgb"].sum().backward()
        outs.append((out["rgb"].detach().clone(), model.materials_a.grad.clone(), model.materials_r.grad.clone()))
    assert m.scene.texture_key()[0] == "rgbe"                                                # the packed texels were used
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-7) and torch.allclose(outs[0][2], outs[1][2], rtol=1e-4, atol=1e-7)
    out = m(mvp.to(cuda_device), 0, cam, -1)                                                 # stage -1 swaps the texture
    assert bool(torch.isfinite(out["rgb"]).all()) and m.scene.texture_key()[0] != "rgbe"
    m(mvp.to(cuda_device), 0, cam, 2)
    assert m.scene.texture_key()[0] == "rgbe"


def test_conf_driven_tracer_runs_from_textures(tx, cuda_device, tmp_path):
    """conf_models.TracerO3d: constructor of tracer_o3d_irt.py:36-96 from files; forward() on position / normal textures
    left by calcute_position_normal_texture equals the array-built tracer (the panorama producer itself needs the
    reference's utils.Cube2Pano and is exercised where that is importable)"""
    from texir_code_b200 import conf_models, models, synth
    sc = synth.make_room_scene(5000, seed=7, clutter=4, tex_res=64, uv_kind="planar")
    rng = np.random.default_rng(10)
    mk = lambda n: np.concatenate([rng.integers(1, 256, (n, n, 3)), rng.integers(120, 132, (n, n, 1))], -1).astype(np.uint8)
    rgbe_tex = mk(64)
    idx = np.zeros((1024, 1024, 3), np.uint16)
    idx[..., 0] = rng.integers(0, 50001, (1024, 1024))
    idx[..., 1] = rng.integers(0, 50001, (1024, 1024))
    idx[:4] = 0
    obj = _write_scene_files(tmp_path, sc, rgbe_tex, mk(8), index_png=idx)
    conf = conf_models.ConfigLite.parse_string("""
        train{ path_mesh_open3d = %s
               sample_light = [64, 16]
               hdr_exposure = 5 }
        models{ render{ sample_type = [uniform, importance] } }
    """ % obj)
    t = conf_models.TracerO3d(conf=conf, ids=[0], extrinsics=[None], optim_cam=False)
    assert t.index_texture.shape == (1024, 1024, 3) and t.index_texture.dtype == np.uint16 and t.sample_l == [64, 16]
    pos, nrm = synth.sample_surface_points(sc, 512 * 1024, seed=3)
    pano = torch.from_numpy(np.concatenate([pos, nrm], -1).reshape(512, 1024, 6)).to(cuda_device)
    t.calcute_position_normal_texture([pano], t.index_texture)
    torch.manual_seed(5)
    out = t()                                                     # the runner's call: self.model()
    assert out.shape == (1024, 1024, 3) and bool((out[:4] == 0).all()) and float(out.abs().max()) > 0
    ref = models.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, t.texture, sample_l=[64, 16], device=cuda_device)
    torch.manual_seed(5)
    want = ref.forward(t.position_texture, t.normal_texture, t.index_texture)
    assert torch.equal(out, want) and t.scene.texture_key()[0] == "rgbe"


def test_device_shift_stream_is_torch_rand(tx, cuda_device):
    """draw_shift for a cube map's 98 304 pixels runs the CPU generator's mt19937 stream on the device
    (texir_mt19937_uniform): the values AND the state the host generator is left in must be those of
    torch.rand(b, 1, 2) -- from a fresh seed, from mid-block, and across consecutive draws of different sizes"""
    from texir_code_b200 import sample_util
    assert sample_util.MT_MIN_DRAWS <= 2 * 98304
    keep = torch.get_rng_state()
    sample_util.draw_shift(98304, cuda_device)                       # first use runs the one-time self-check (one extra launch)
    torch.set_rng_state(keep)
    for seed in (666, 12345):
        torch.manual_seed(seed)
        torch.rand(333)                                              # not on a block boundary
        st = torch.get_rng_state()
        want = [torch.rand(98304, 1, 2), torch.rand(7, 1, 2), torch.rand(98304, 1, 2), torch.rand(49152, 1, 2)]
        end = torch.get_rng_state()
        torch.set_rng_state(st)
        n0 = tx.launch_count()
        got = [sample_util.draw_shift(98304, cuda_device), sample_util.draw_shift(7, cuda_device),
               sample_util.draw_shift(98304, cuda_device), sample_util.draw_shift(49152, cuda_device)]
        assert tx.launch_count() - n0 == 3 and sample_util._mt_ok is True   # three device draws, the 7-pixel one on the host
        for g, w in zip(got, want):
            assert g.shape == w.shape and torch.equal(g.cpu(), w)
        assert torch.equal(torch.get_rng_state(), end)
    # deferred hand-back (what MaterialModel.forward wraps itself in): same values, the state arrives when the block ends,
    # and a second draw inside the block starts from the first one's state
    torch.manual_seed(99)
    want = [torch.rand(98304, 1, 2), torch.rand(49152, 1, 2)]
    end = torch.get_rng_state()
    torch.manual_seed(99)
    with sample_util.deferred_rng_commit():
        a = sample_util.draw_shift(98304, cuda_device)
        assert sample_util._mt_pending is not None
        b = sample_util.draw_shift(49152, cuda_device)
    assert sample_util._mt_pending is None and torch.equal(torch.get_rng_state(), end)
    assert torch.equal(a.cpu(), want[0]) and torch.equal(b.cpu(), want[1])
    # the model's step draws through it: same colours as with the host draw
    os.environ["TEXIR_SHIFT_DEVICE"] = "0"
    try:
        torch.manual_seed(5)
        a = sample_util.draw_shift(98304, cuda_device)
    finally:
        del os.environ["TEXIR_SHIFT_DEVICE"]
    torch.manual_seed(5)
    assert torch.equal(sample_util.draw_shift(98304, cuda_device), a)
