"""Times the IrT post-processing ops on cuda:0 (CUDA events, after warm-up):  python profiles/profile_postprocess.py"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texir_code_b200 import postprocess as pp  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(0)
for res in (1024, 4096):
    img = np.exp(rng.normal(0, 1, (res, res, 3))).astype(np.float32)
    coarse = rng.random((res // 8, res // 8)) < 0.5
    img[~np.kron(coarse, np.ones((8, 8), bool))] = 0.0
    t = torch.from_numpy(img).to(dev)
    ms_pad = timed(lambda: pp.pad_texture(t))
    ms_enc = timed(lambda: pp.rgbe_encode(t))
    q = pp.rgbe_encode(t)
    ms_dec = timed(lambda: pp.rgbe_decode(q))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "0_irr_texture.hdr")
        ms_write = timed(lambda: pp.imwrite_hdr(path, t), reps=3)
        ms_read = timed(lambda: pp.imread_hdr(path, dev), reps=3)
        size = os.path.getsize(path)
    print("%d^2: pad_texture %.3f ms | rgbe_encode %.3f ms (%.0f GB/s) | rgbe_decode %.3f ms | imwrite_hdr %.1f ms | "
          "imread_hdr %.1f ms (file %.1f MB)" % (res, ms_pad, ms_enc, res * res * 16 / ms_enc / 1e6, ms_dec, ms_write, ms_read,
                                                size / 1e6))
