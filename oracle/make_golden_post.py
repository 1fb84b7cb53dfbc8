"""make_golden_post.py -- TEST INFRASTRUCTURE.  Goldens of the IrT post-processing row, written by
executing the REFERENCE's own tools/padding_texture.py (from /root/reference, read-only; it needs only cv2,
scipy, numpy and torch, all present in the build container) and cv2's own .hdr codec:

    tests/golden/post_<tag>.npz
        img        float32 [H,W,3] BGR, what generate_ir_texture.py:80-82 hands to cv2.imwrite
        hdr_in     bytes of cv2.imwrite('0_irr_texture.hdr', img)
        img_in     cv2.imread(hdr_in, -1)            (the RGBE-quantised image the padding tool sees)
        indices    scipy's nearest-non-empty indices for img_in's mask ([2,H,W], as the tool computes them)
        hdr_out    bytes of the file the reference's padding_texture() wrote ('t.hdr')
        img_out    cv2.imread(hdr_out, -1)

    tests/golden/post_index.npz     the 16-bit PNG index texture through the tool's padding_index_texture (:8-43)
        png_in     bytes of cv2.imwrite('idx.png', uint16 [H,W,3])
        img_in     cv2.imread(png_in, -1)
        img_out    cv2.imread of the 'idx_padding.png' the reference's tool wrote

The external `oidnDenoise` command the tool shells out to afterwards does not exist here; os.system just reports
that and the padded file is already on disk.

Run:  python oracle/make_golden_post.py      (needs /root/reference; the GPU box never runs this)
"""
from __future__ import annotations

import os
import sys
import tempfile

import cv2
import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("TEXIR_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def island_texture(H, W, seed, fill=0.55):
    """an atlas-like irradiance texture: blobs of valid texels (HDR values) on an empty (all-zero) background"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    valid = np.zeros((H, W), bool)
    for _ in range(max(3, int(fill * H * W / 60))):
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        ry, rx = rng.uniform(1, 6), rng.uniform(1, 6)
        valid |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    valid[rng.random((H, W)) < 0.02] = True   # isolated texels: many equal-distance cases on the integer grid
    img = np.exp(rng.normal(0.0, 1.5, (H, W, 3))).astype(np.float32)
    img[rng.random((H, W)) < 0.01] *= 200.0   # emitters
    img[~valid] = 0.0
    return img


def main():
    sys.path.insert(0, os.path.join(REF, "tools"))
    import padding_texture as ref_tool          # the reference's own module
    os.makedirs(OUT, exist_ok=True)
    for tag, H, W, seed in (("pow2", 64, 64, 1), ("odd", 45, 83, 2), ("sparse", 96, 40, 3)):
        img = island_texture(H, W, seed, fill=0.15 if tag == "sparse" else 0.55)
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "0_irr_texture.hdr")
            cv2.imwrite(src, img)
            hdr_in = np.frombuffer(open(src, "rb").read(), np.uint8)
            img_in = cv2.imread(src, -1)
            ref_tool.padding_texture(src)        # writes <tmp>/t.hdr, then tries to run oidnDenoise (absent)
            dst = os.path.join(tmp, "t.hdr")
            hdr_out = np.frombuffer(open(dst, "rb").read(), np.uint8)
            img_out = cv2.imread(dst, -1)
        mask = np.asarray(img_in[:, :, 0] + img_in[:, :, 1] + img_in[:, :, 2] == 0.0, np.uint8)
        _, indices = ndimage.distance_transform_edt(mask, return_indices=True)
        np.savez_compressed(os.path.join(OUT, "post_%s.npz" % tag), img=img, hdr_in=hdr_in, img_in=img_in,
                            indices=indices.astype(np.int32), hdr_out=hdr_out, img_out=img_out)
        print("wrote post_%s.npz: %dx%d, %d empty texels, hdr %d -> %d bytes" % (tag, H, W, int(mask.sum()), len(hdr_in),
                                                                               len(hdr_out)))
    # the index texture (view id, pixel column, pixel row of the panorama a texel was seen in), 16-bit PNG
    rng = np.random.default_rng(7)
    H, W = 72, 56
    valid = island_texture(H, W, 4)[:, :, 0] > 0
    idx = np.stack([rng.integers(1, 40, (H, W)), rng.integers(0, 65536, (H, W)), rng.integers(0, 65536, (H, W))], -1).astype(np.uint16)
    idx[~valid] = 0
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "idx.png")
        cv2.imwrite(src, idx)
        png_in = np.frombuffer(open(src, "rb").read(), np.uint8)
        img_in = cv2.imread(src, -1)
        ref_tool.padding_index_texture(src)
        img_out = cv2.imread(os.path.join(tmp, "idx_padding.png"), -1)
    assert img_in.dtype == np.uint16 and img_out.dtype == np.uint16 and np.array_equal(img_in, idx)
    np.savez_compressed(os.path.join(OUT, "post_index.npz"), png_in=png_in, img_in=img_in, img_out=img_out)
    print("wrote post_index.npz: %dx%d, %d empty texels, png %d bytes" % (H, W, int((~valid).sum()), len(png_in)))


if __name__ == "__main__":
    main()
