"""CPU: the IrT post-processing row.  The goldens tests/golden/post_*.npz were written by the
reference's own tools/padding_texture.py and cv2 (oracle/make_golden_post.py).  Checked here without a GPU:
  * the oracle restatement against those goldens,
  * the device header texir_post.h (compiled with g++, tests/cpu_emul) against goldens, scipy and the oracle,
  * the host-only .hdr container entry points of the C ABI against cv2's bytes."""
import ctypes
import os

import numpy as np
import pytest

from conftest import load_golden

TAGS = ["pow2", "odd", "sparse"]


def _emul_pad(emul, img):
    img = np.ascontiguousarray(img, np.float32)
    H, W, C = img.shape
    out = np.empty_like(img)
    idx = np.empty((H, W, 2), np.int32)
    emul.emul_pad_texture(img.ctypes.data, H, W, C, out.ctypes.data, idx.ctypes.data)
    return out, idx


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_padding_matches_reference_tool(oracle, tag):
    g = load_golden("post_" + tag)
    res, idx = oracle.padding_texture_array(g["img_in"])
    assert np.array_equal(idx, g["indices"])
    # the reference wrote `res` through cv2.imwrite: compare after the same RGBE quantisation
    assert np.array_equal(oracle.rgbe2float(oracle.float2rgbe(res)), g["img_out"])
    assert np.array_equal(res, g["img_out"])  # padded values are copies of already-quantised texels


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_rgbe_codec_matches_cv2(oracle, tag):
    g = load_golden("post_" + tag)
    assert np.array_equal(oracle.rgbe2float(oracle.float2rgbe(g["img"])), g["img_in"])
    cv2 = pytest.importorskip("cv2")
    dec = cv2.imdecode(g["hdr_in"], -1)
    assert np.array_equal(dec, g["img_in"])


@pytest.mark.parametrize("tag", TAGS)
def test_device_header_padding_matches_reference_tool(emul, tag):
    g = load_golden("post_" + tag)
    out, idx = _emul_pad(emul, g["img_in"])
    assert np.array_equal(idx.transpose(2, 0, 1), g["indices"])
    assert np.array_equal(out, g["img_out"])


def test_device_header_padding_ties_and_shapes_vs_scipy(emul, oracle):
    """random masks on small integer grids are full of equal-distance cases; every index must be scipy's"""
    pytest.importorskip("scipy")
    rng = np.random.default_rng(7)
    shapes = [(1, 1), (1, 17), (23, 1), (2, 2), (9, 31), (40, 40), (33, 64), (64, 7), (57, 129)]
    for H, W in shapes:
        for p in (0.02, 0.2, 0.6, 0.95):
            img = rng.uniform(0.5, 2.0, (H, W, 3)).astype(np.float32)
            img[rng.random((H, W)) > p] = 0.0
            if not img.any():
                img[rng.integers(H), rng.integers(W)] = 1.0
            out, idx = _emul_pad(emul, img)
            ref, ref_idx = oracle.padding_texture_array(img)
            assert np.array_equal(idx.transpose(2, 0, 1), ref_idx), (H, W, p)
            assert np.array_equal(out, ref), (H, W, p)


def test_device_header_padding_degenerate_inputs(emul):
    full = np.random.default_rng(0).uniform(0.1, 1, (8, 9, 3)).astype(np.float32)
    out, idx = _emul_pad(emul, full)                       # nothing to pad: identity
    assert np.array_equal(out, full)
    assert np.array_equal(idx[..., 0], np.arange(8)[:, None].repeat(9, 1))
    empty = np.zeros((5, 6, 3), np.float32)                # nothing to pad FROM: stays empty
    out, _ = _emul_pad(emul, empty)
    assert np.array_equal(out, empty)
    cancel = np.zeros((4, 4, 3), np.float32)               # channel sum 0 but non-zero channels: "empty" like the tool
    cancel[1, 1] = [1.0, -1.0, 0.0]
    cancel[3, 2] = [0.25, 0.5, 0.75]
    out, idx = _emul_pad(emul, cancel)
    assert tuple(idx[1, 1]) == (3, 2)


def test_device_header_rgbe_codec(emul, oracle):
    rng = np.random.default_rng(3)
    img = np.exp(rng.normal(0, 6, (4096, 3))).astype(np.float32)
    img[:64] = 0
    img[64:128] = rng.uniform(0, 2e-32, (64, 3)).astype(np.float32)   # around the 1e-32 cut-off
    img[128:192, 1:] = 0                                              # one bright channel
    img[192:200] = [[1, 1, 1], [0.5, 0.5, 0.5], [255, 1, 0], [256, 0.999, 2], [1e30, 1, 1], [3e38, 3e38, 1],
                    [0.99999994, 1, 1.0000001], [1e-31, 2e-31, 0]]
    q = np.empty((len(img), 4), np.uint8)
    emul.emul_rgbe_encode(img.ctypes.data, len(img), 1, q.ctypes.data)
    assert np.array_equal(q, oracle.float2rgbe(img))
    back = np.empty_like(img)
    emul.emul_rgbe_decode(q.ctypes.data, len(img), 1, back.ctypes.data)
    assert np.array_equal(back, oracle.rgbe2float(q))
    rgb = np.empty_like(img)
    emul.emul_rgbe_decode(q.ctypes.data, len(img), 0, rgb.ctypes.data)
    assert np.array_equal(rgb, back[:, ::-1])


@pytest.fixture(scope="module")
def lib():
    from texir_code_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize("tag", TAGS)
def test_hdr_container_bytes_match_cv2(lib, oracle, tmp_path, tag):
    g = load_golden("post_" + tag)
    H, W = g["img"].shape[:2]
    rgbe = np.ascontiguousarray(oracle.float2rgbe(g["img"]))
    path = os.fsencode(str(tmp_path / "w.hdr"))
    assert lib.texir_hdr_write_host(path, rgbe.ctypes.data, H, W) == 0, lib.texir_last_error()
    assert open(path, "rb").read() == g["hdr_in"].tobytes()
    # and back: cv2's own file through the reader
    src = tmp_path / "r.hdr"
    src.write_bytes(g["hdr_out"].tobytes())
    h, w = ctypes.c_int(), ctypes.c_int()
    assert lib.texir_hdr_read_info(os.fsencode(str(src)), ctypes.byref(h), ctypes.byref(w)) == 0
    assert (h.value, w.value) == (H, W)
    back = np.zeros((H, W, 4), np.uint8)
    assert lib.texir_hdr_read_host(os.fsencode(str(src)), back.ctypes.data, H, W) == 0, lib.texir_last_error()
    assert np.array_equal(oracle.rgbe2float(back), g["img_out"])


def test_hdr_container_flat_and_long_scanlines(lib, oracle, tmp_path):
    """widths below 8 are stored flat; runs are capped at 127 and literals at 128 bytes; checked against cv2"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for H, W in ((3, 5), (2, 7), (4, 8), (3, 700)):
        img = np.exp(rng.normal(0, 1, (H, W, 3))).astype(np.float32)
        img[:, W // 4: W // 4 + W // 2] = 0.75            # a run longer than 127 when W = 700
        ref = tmp_path / "ref.hdr"
        cv2.imwrite(str(ref), img)
        rgbe = np.ascontiguousarray(oracle.float2rgbe(img))
        mine = tmp_path / "mine.hdr"
        assert lib.texir_hdr_write_host(os.fsencode(str(mine)), rgbe.ctypes.data, H, W) == 0
        assert mine.read_bytes() == ref.read_bytes(), (H, W)
        back = np.zeros((H, W, 4), np.uint8)
        assert lib.texir_hdr_read_host(os.fsencode(str(ref)), back.ctypes.data, H, W) == 0
        assert np.array_equal(back, rgbe)


def test_hdr_container_rejects_bad_files(lib, tmp_path):
    h, w = ctypes.c_int(), ctypes.c_int()
    assert lib.texir_hdr_read_info(os.fsencode(str(tmp_path / "missing.hdr")), ctypes.byref(h), ctypes.byref(w)) != 0
    assert b"cannot open" in lib.texir_last_error()
    bad = tmp_path / "bad.hdr"
    bad.write_bytes(b"P6\n2 2\n255\n" + bytes(12))
    assert lib.texir_hdr_read_info(os.fsencode(str(bad)), ctypes.byref(h), ctypes.byref(w)) != 0
    trunc = tmp_path / "trunc.hdr"
    trunc.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 4 +X 16\n\x02\x02\x00\x10\x90")
    buf = np.zeros((4, 16, 4), np.uint8)
    assert lib.texir_hdr_read_host(os.fsencode(str(trunc)), buf.ctypes.data, 4, 16) != 0
    assert b"truncated" in lib.texir_last_error()
    assert lib.texir_hdr_read_host(os.fsencode(str(trunc)), buf.ctypes.data, 4, 8) != 0  # size mismatch
    assert lib.texir_pad_texture(None, 4, 4, 3, None, None, None, None) != 0              # validation, no GPU touched
    assert lib.texir_pad_texture(None, 4, 4, 5, None, None, None, None) != 0


def test_index_texture_padding_vs_reference_tool(emul, oracle):
    """tools/padding_texture.py:8-43 (16-bit PNG index texture): oracle and device header against what the tool wrote"""
    g = load_golden("post_index")
    img = g["img_in"].astype(np.float32)
    res, _ = oracle.padding_texture_array(img)
    assert np.array_equal(res.astype(np.uint16), g["img_out"])
    out, _ = _emul_pad(emul, img)
    assert np.array_equal(out.astype(np.uint16), g["img_out"])


def test_png_codec_reads_cv2_files_and_cv2_reads_ours(lib, tmp_path):
    """imread_png / imwrite_png (chunks + zlib in Python, scanline filters in the library) against cv2's PNG codec:
    the committed 16-bit file of the golden, then every supported layout, smooth images (cv2 picks Sub / Up / Average /
    Paeth filters on them) and noise"""
    from texir_code_b200 import postprocess as pp
    g = load_golden("post_index")
    src = tmp_path / "idx.png"
    src.write_bytes(g["png_in"].tobytes())
    got = pp.imread_png(str(src))
    assert got.dtype == np.uint16 and np.array_equal(got, g["img_in"])
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:37, 0:53]
    for dtype, top in ((np.uint8, 255), (np.uint16, 65535)):
        for C in (1, 3, 4):
            smooth = ((yy * 3 + xx * 5)[:, :, None] * (np.arange(C) + 1) * (top // 600 + 1) % (top + 1)).astype(dtype)
            noise = rng.integers(0, top + 1, (37, 53, C)).astype(dtype)
            for img in (smooth, noise):
                img = img[:, :, 0] if C == 1 else img
                a, b = str(tmp_path / "a.png"), str(tmp_path / "b.png")
                cv2.imwrite(a, img)
                assert np.array_equal(pp.imread_png(a), cv2.imread(a, -1))
                pp.imwrite_png(b, img)
                back = cv2.imread(b, -1)
                assert back.dtype == img.dtype and np.array_equal(back, img)
                assert np.array_equal(pp.imread_png(b), img)
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    from texir_code_b200._lib import TexirError
    with pytest.raises(TexirError):
        pp.imread_png(str(bad))
    raw = np.zeros((2, 1 + 6), np.uint8)
    raw[1, 0] = 9                                                     # no such filter type
    out = np.zeros((2, 6), np.uint8)
    assert lib.texir_png_unfilter_host(raw.ctypes.data, 2, 6, 6, out.ctypes.data) != 0
