"""GPU (-m gpu): IrT post-processing through the package's binding of the C ABI, against the goldens
written by the reference's own tools/padding_texture.py + cv2, against the oracle (scipy + torch on the CPU) at
sizes it finishes in seconds, and through size-independent properties at the 4096^2 texture size.
Everything on this row is index / byte work: the bar is bit-exact."""

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

TAGS = ["pow2", "odd", "sparse"]


@pytest.fixture(scope="module")
def pp(cuda_device):
    import texir_code_b200 as tx
    tx._lib.load()
    return tx.postprocess


@pytest.mark.parametrize("tag", TAGS)
def test_pad_texture_matches_reference_tool(pp, cuda_device, tag):
    g = load_golden("post_" + tag)
    out, idx = pp.pad_texture(torch.from_numpy(g["img_in"]).to(cuda_device), return_indices=True)
    assert np.array_equal(idx.cpu().numpy().transpose(2, 0, 1), g["indices"])
    assert np.array_equal(out.cpu().numpy(), g["img_out"])


@pytest.mark.parametrize("tag", TAGS)
def test_padding_texture_tool_writes_the_reference_file(pp, cuda_device, tmp_path, tag):
    """the whole tool: .hdr in -> pad -> .hdr out, byte for byte the file the reference's tool wrote"""
    g = load_golden("post_" + tag)
    src = tmp_path / "0_irr_texture.hdr"
    src.write_bytes(g["hdr_in"].tobytes())
    called = []
    dst = pp.padding_texture(str(src), denoise=lambda a, b: called.append((a, b)), device=cuda_device)
    assert dst == str(tmp_path / "t.hdr")
    assert open(dst, "rb").read() == g["hdr_out"].tobytes()
    assert called == [(dst, str(tmp_path / "0_irr_texture_denoised_padding.hdr"))]


@pytest.mark.parametrize("tag", TAGS)
def test_hdr_io_matches_cv2(pp, cuda_device, tmp_path, tag):
    g = load_golden("post_" + tag)
    path = str(tmp_path / "w.hdr")
    pp.imwrite_hdr(path, torch.from_numpy(g["img"]).to(cuda_device))
    assert open(path, "rb").read() == g["hdr_in"].tobytes()
    back = pp.imread_hdr(path, cuda_device)
    assert np.array_equal(back.cpu().numpy(), g["img_in"])
    # generate_ir_texture.py:80-82 hands cv2 the channel-reversed model output
    pp.save_irr_texture(torch.from_numpy(g["img"][:, :, ::-1].copy()).to(cuda_device), path)
    assert open(path, "rb").read() == g["hdr_in"].tobytes()


def test_rgbe_codec_vs_oracle(pp, cuda_device, oracle):
    rng = np.random.default_rng(3)
    img = np.exp(rng.normal(0, 6, (1 << 16, 3))).astype(np.float32)
    img[:64] = 0
    img[64:128] = rng.uniform(0, 2e-32, (64, 3)).astype(np.float32)
    img[128:136] = [[1, 1, 1], [0.5, 0.5, 0.5], [255, 1, 0], [256, 0.999, 2], [1e30, 1, 1], [3e38, 3e38, 1],
                    [0.99999994, 1, 1.0000001], [1e-31, 2e-31, 0]]
    q = pp.rgbe_encode(torch.from_numpy(img).to(cuda_device))
    assert np.array_equal(q.cpu().numpy(), oracle.float2rgbe(img))
    assert np.array_equal(pp.rgbe_decode(q).cpu().numpy(), oracle.rgbe2float(q.cpu().numpy()))
    assert np.array_equal(pp.rgbe_decode(q, bgr=False).cpu().numpy(), oracle.rgbe2float(q.cpu().numpy())[:, ::-1])


def _islands(H, W, seed, fill):
    rng = np.random.default_rng(seed)
    img = np.exp(rng.normal(0.0, 1.0, (H, W, 3))).astype(np.float32)
    coarse = rng.random((H // 8 + 1, W // 8 + 1)) < fill
    valid = np.kron(coarse, np.ones((8, 8), bool))[:H, :W] | (rng.random((H, W)) < 0.002)
    img[~valid] = 0.0
    return img


@pytest.mark.parametrize("shape,fill", [((1024, 1024), 0.5), ((600, 1000), 0.3), ((777, 333), 0.05), ((1, 64), 0.5),
                                        ((64, 1), 0.5)])
def test_pad_texture_vs_oracle(pp, cuda_device, oracle, shape, fill):
    """the reference's IrT is 1024^2 (tracer_o3d_irt.py); also non-power-of-two sizes, where grid_sample's
    float32 un-normalisation decides which neighbour an odd index reads"""
    img = _islands(shape[0], shape[1], 11, fill)
    if not img.any():
        img[0, 0] = 1.0
    out, idx = pp.pad_texture(torch.from_numpy(img).to(cuda_device), return_indices=True)
    ref, ref_idx = oracle.padding_texture_array(img)
    assert np.array_equal(idx.cpu().numpy().transpose(2, 0, 1), ref_idx)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_pad_texture_properties_at_full_size(pp, cuda_device):
    """4096^2 (the radiance-texture size of BASELINE configs[1]): non-empty texels are untouched; the reported site of
    every texel is a non-empty texel; the site distances form an exact Euclidean distance field (0 on non-empty
    texels, 1-Lipschitz between neighbours, never beaten by a neighbour's site); an empty texel holds the value
    grid_sample('nearest') reads for its site -- at a power-of-two size that is texel (row & ~1, col & ~1), the
    reference's half-to-even rounding (texir_post.h), which may itself be empty; the op is deterministic"""
    H = W = 4096
    img = torch.from_numpy(_islands(H, W, 21, 0.4)).to(cuda_device)
    out, idx = pp.pad_texture(img, return_indices=True)
    empty = img.sum(-1) == 0
    assert torch.equal(out[~empty], img[~empty])
    sr, sc = idx[..., 0].long(), idx[..., 1].long()
    assert not bool(empty[sr, sc].any())
    assert torch.equal(out[empty], img[sr & ~1, sc & ~1][empty])
    ii = torch.arange(H, device=cuda_device).view(H, 1).expand(H, W)
    jj = torch.arange(W, device=cuda_device).view(1, W).expand(H, W)
    d2 = (sr - ii) ** 2 + (sc - jj) ** 2
    assert bool((d2[~empty] == 0).all())
    for dy, dx in ((0, 1), (1, 0)):
        a, b = d2[dy:, dx:], d2[:H - dy, :W - dx]
        assert float((a.double().sqrt() - b.double().sqrt()).abs().max()) <= 1.0 + 1e-9
        # my neighbour's site is a candidate for me as well: it can never be strictly nearer than my own site
        nsr, nsc = sr[dy:, dx:], sc[dy:, dx:]
        cand = (nsr - ii[:H - dy, :W - dx]) ** 2 + (nsc - jj[:H - dy, :W - dx]) ** 2
        assert bool((b <= cand).all())
        nsr, nsc = sr[:H - dy, :W - dx], sc[:H - dy, :W - dx]
        cand = (nsr - ii[dy:, dx:]) ** 2 + (nsc - jj[dy:, dx:]) ** 2
        assert bool((a <= cand).all())
    again, idx2 = pp.pad_texture(img, return_indices=True)
    assert torch.equal(again, out) and torch.equal(idx2, idx)


def test_pad_texture_rejects_cpu_tensors(pp):
    import texir_code_b200 as tx
    with pytest.raises(tx.TexirError):
        pp.pad_texture(torch.zeros(4, 4, 3))


def test_padding_index_texture_tool(pp, cuda_device, tmp_path):
    """tools/padding_texture.py:8-43: 16-bit PNG in -> pad -> '<name>_padding.png', the pixels the reference's tool wrote;
    a fourth channel stays out of the mask and is sampled like the others"""
    g = load_golden("post_index")
    src = tmp_path / "idx.png"
    src.write_bytes(g["png_in"].tobytes())
    dst = pp.padding_index_texture(str(src), device=cuda_device)
    assert dst == str(tmp_path / "idx_padding.png")
    out = pp.imread_png(dst)
    assert out.dtype == np.uint16 and np.array_equal(out, g["img_out"])
    rng = np.random.default_rng(2)
    alpha = rng.integers(0, 65536, g["img_in"].shape[:2]).astype(np.uint16)
    four = np.concatenate([g["img_in"], alpha[:, :, None]], -1)
    src4 = tmp_path / "four.png"
    pp.imwrite_png(str(src4), four)
    out4 = pp.imread_png(pp.padding_index_texture(str(src4), device=cuda_device))
    assert np.array_equal(out4[:, :, :3], g["img_out"])
    empty = g["img_in"].astype(np.int64).sum(-1) == 0
    assert np.array_equal(out4[~empty, 3], alpha[~empty])
    # the fourth channel of a padded texel comes from the texel its colour came from (colours here are unique enough)
    key = {tuple(c): a for c, a in zip(g["img_in"][~empty].tolist(), alpha[~empty].tolist())}
    picked = out4[empty]
    assert all(key[tuple(p[:3])] == p[3] for p in picked.tolist() if tuple(p[:3]) != (0, 0, 0))
