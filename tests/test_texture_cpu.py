"""CPU: the mip-mapped material-texture fetch.  texir_tex.h compiled with g++ (tests/cpu_emul) against
the oracle's torch restatement of dr.texture('linear-mipmap-linear'); nvdiffrast itself is not installable here, so
this pins the device arithmetic to the oracle, not the oracle to nvdiffrast (oracle header: parity unpinned)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2


def _case(H, W, C, P, seed):
    rng = np.random.default_rng(seed)
    tex = rng.uniform(0, 1, (H, W, C)).astype(np.float32)
    uv = rng.uniform(-0.3, 1.3, (P, 2)).astype(np.float32)                  # wraps on both sides
    scale = np.exp(rng.uniform(np.log(0.05), np.log(40.0), (P, 1))) / max(H, W)
    uv_da = (rng.normal(size=(P, 4)) * scale).astype(np.float32)            # magnified ... strongly minified
    uv_da[:8] = 0.0                                                         # zero footprint: log2(0) = -inf
    return tex, uv, uv_da


def _emul_fetch(emul, tex, uv, uv_da, max_level):
    H, W, C = tex.shape
    L = emul.emul_mip_levels(H, W, max_level)
    mips = np.zeros(max(int(emul.emul_mip_elems(H, W, C, L)), 1), np.float32)
    emul.emul_mip_build(tex.ctypes.data, H, W, C, L, mips.ctypes.data)
    out = np.empty((len(uv), C), np.float32)
    lod = np.empty(len(uv), np.float32)
    emul.emul_texture_fetch_mip(tex.ctypes.data, mips.ctypes.data, H, W, C, L, uv.ctypes.data, uv_da.ctypes.data,
                                len(uv), out.ctypes.data, lod.ctypes.data)
    return out, lod, L, mips


@pytest.mark.parametrize("shape,max_level", [((64, 64, 3), -1), ((32, 128, 1), -1), ((64, 64, 3), 2), ((12, 8, 2), 2),
                                             ((1, 16, 3), -1)])
def test_mip_fetch_device_header_vs_oracle(emul, oracle, shape, max_level):
    tex, uv, uv_da = _case(*shape, 4000, 1)
    out, lod, L, mips = _emul_fetch(emul, tex, uv, uv_da, max_level)
    ml = None if max_level < 0 else max_level
    levels = oracle.mip_pyramid(torch.from_numpy(tex), ml)
    assert L == len(levels) - 1
    off = 0
    for l in range(1, L + 1):                                                # the pyramid itself, level by level
        n = levels[l].numel()
        assert rel_l2(mips[off:off + n], levels[l].numpy().reshape(-1)) < 1e-6
        off += n
    ref_lod = oracle.texture_lod(torch.from_numpy(uv_da), shape[0], shape[1], L).numpy()
    assert np.max(np.abs(lod - ref_lod)) < 1e-4
    ref = oracle.texture_mipmap_linear(torch.from_numpy(tex), torch.from_numpy(uv), torch.from_numpy(uv_da), ml).numpy()
    assert rel_l2(out, ref) < 1e-5
    assert np.max(np.abs(out - ref)) < 1e-4


def test_mip_level_rules(emul):
    assert emul.emul_mip_levels(2048, 2048, 13) == 11      # materials_a with get_mip_level(8192) = 13
    assert emul.emul_mip_levels(4096, 4096, 13) == 12      # materials_r
    assert emul.emul_mip_levels(4096, 4096, 3) == 3
    assert emul.emul_mip_levels(6, 4, -1) == -1            # 6 -> 3 is odd and would have to be halved again
    assert emul.emul_mip_levels(6, 4, 1) == 1
    assert emul.emul_mip_levels(1, 1, -1) == 0
    assert emul.emul_mip_levels(1, 8, -1) == 3


def test_mip_fetch_limits(emul, oracle):
    """zero derivatives select level 0 (= the plain bilinear fetch); a footprint beyond the coarsest level returns the
    texture mean"""
    tex, uv, _ = _case(32, 32, 3, 256, 2)
    zero = np.zeros((256, 4), np.float32)
    out, lod, L, _ = _emul_fetch(emul, tex, uv, zero, -1)
    assert np.all(lod == 0)
    ref = oracle.texture_linear(torch.from_numpy(tex), torch.from_numpy(uv)).numpy()
    assert rel_l2(out, ref) < 1e-6
    huge = np.full((256, 4), 10.0, np.float32)
    out, lod, L, _ = _emul_fetch(emul, tex, uv, huge, -1)
    assert np.all(lod == L)
    assert np.allclose(out, tex.reshape(-1, 3).mean(0)[None], rtol=1e-5)


def test_level_of_detail_is_log2_of_the_footprint_major_axis(oracle):
    """the closed form in texture_lod / mip_lod (l2b + sqrt(l2n)) is the largest eigenvalue of J J^T, J = d(s,t)/d(X,Y)
    in texels: the squared length of the major axis of the pixel's footprint ellipse"""
    rng = np.random.default_rng(4)
    H, W, L = 256, 512, 8
    da = (rng.normal(size=(2000, 4)) * np.exp(rng.uniform(-8, -2, (2000, 1)))).astype(np.float32)
    lod = oracle.texture_lod(torch.from_numpy(da), H, W, L).numpy()
    J = np.stack([np.stack([da[:, 0] * W, da[:, 1] * W], -1), np.stack([da[:, 2] * H, da[:, 3] * H], -1)], -2).astype(np.float64)
    smax = np.linalg.svd(J, compute_uv=False)[:, 0]
    ref = np.clip(np.log2(smax), 0, L)
    assert np.max(np.abs(lod - ref)) < 1e-4
