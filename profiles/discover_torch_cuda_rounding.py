"""Which float32 formula does torch-CUDA evaluate for each step of generate_dir (utils/sample_util.py:84-143)?

The reference runs generate_dir on the GPU (every tensor is .cuda()), so the like-for-like direction oracle is torch
on CUDA.  This script takes each building block, evaluates candidate formulas with exactly rounded float32 steps
(emulated through float64: a product of two float32 is exact in float64, so fma(a, b, c) = float32(a*b + c) up to a
2^-29 double-rounding chance) and prints the fraction of inputs on which a candidate reproduces torch-CUDA bit for bit.
The winners are what csrc/texir_math.h implements on the device.      python profiles/discover_torch_cuda_rounding.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

dev = torch.device("cuda:0")
torch.manual_seed(3)
n = 400000


def f32(x):
    return x.to(torch.float32)


def fma(a, b, c):  # float32 fma through float64
    return f32(a.double() * b.double() + c.double())


def rate(a, b):
    return float((a == b).float().mean())


x = torch.randn(n, 3, device=dev)
X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
ref = torch.norm(x.unsqueeze(1).expand(n, 4, 3), dim=-1, keepdim=True)[:, 0, 0]
cands = {
    "sqrt(fma(z,z,fma(y,y,x*x)))": torch.sqrt(fma(Z, Z, fma(Y, Y, X * X))),
    "sqrt((x*x+y*y)+z*z)": torch.sqrt((X * X + Y * Y) + Z * Z),
    "sqrt(fma(x,x,fma(y,y,z*z)))": torch.sqrt(fma(X, X, fma(Y, Y, Z * Z))),
    "sqrt(x*x+(y*y+z*z))": torch.sqrt(X * X + (Y * Y + Z * Z)),
    "sqrt(fma(y,y,x*x)+z*z)": torch.sqrt(fma(Y, Y, X * X) + Z * Z),
    "f32(sqrt64(x^2+y^2+z^2))": f32(torch.sqrt(X.double() ** 2 + Y.double() ** 2 + Z.double() ** 2)),
    "sqrt32(f32(exact sum))": torch.sqrt(f32(X.double() ** 2 + Y.double() ** 2 + Z.double() ** 2)),
}
print("torch.norm(dim=-1) on CUDA:")
for k, v in cands.items():
    print("   %-34s %.6f" % (k, rate(v, ref)))
# contiguous (non-expanded) input -- what norm_axis(torch.cross(...)) and F.normalize(cam - p) see: another reduction
# schedule (two lanes share the three elements), so possibly another summation order
def norm_cands(X, Y, Z):
    return {
        "sqrt((x*x+y*y)+z*z)": torch.sqrt((X * X + Y * Y) + Z * Z),
        "sqrt((x*x+z*z)+y*y)": torch.sqrt((X * X + Z * Z) + Y * Y),
        "sqrt(x*x+(y*y+z*z))": torch.sqrt(X * X + (Y * Y + Z * Z)),
        "sqrt(fma(z,z,fma(y,y,x*x)))": torch.sqrt(fma(Z, Z, fma(Y, Y, X * X))),
        "sqrt(fma(z,z,x*x)+y*y)": torch.sqrt(fma(Z, Z, X * X) + Y * Y),
        "sqrt(fma(y,y,fma(z,z,x*x)))": torch.sqrt(fma(Y, Y, fma(Z, Z, X * X))),
        "sqrt(fma(y,y,x*x)+z*z)": torch.sqrt(fma(Y, Y, X * X) + Z * Z),
        "sqrt(fma(x,x,z*z)+y*y)": torch.sqrt(fma(X, X, Z * Z) + Y * Y),
    }
for shape in ((n, 3), (512, 16, 3), (512, 2048, 3), (98304, 16, 3), (98304, 3), (512, 3)):
    xs = torch.randn(*shape, device=dev)
    r2 = torch.norm(xs, dim=-1, keepdim=True).reshape(-1)
    f = xs.reshape(-1, 3)
    print("torch.norm(dim=-1) on CUDA, CONTIGUOUS input %s:" % (shape,))
    for k, v in norm_cands(f[:, 0], f[:, 1], f[:, 2]).items():
        print("   %-34s %.6f" % (k, rate(v, r2)))
    if len(shape) == 2:
        fn = torch.nn.functional.normalize(xs, eps=1e-4)
        for k, v in norm_cands(f[:, 0], f[:, 1], f[:, 2]).items():
            print("   F.normalize == x / max(%s, eps): %.6f" % (k, rate(f / torch.clamp(v, min=1e-4).unsqueeze(-1), fn)))
for shape in ((512, 16, 3), (512, 2048, 3), (98304, 16, 3)):
    xs = torch.randn(shape[0], 3, device=dev).unsqueeze(1).expand(*shape)
    r2 = torch.norm(xs, dim=-1, keepdim=True).reshape(-1)
    f = xs.reshape(-1, 3)
    print("torch.norm(dim=-1) on CUDA, EXPANDED input %s:" % (shape,))
    for k, v in norm_cands(f[:, 0], f[:, 1], f[:, 2]).items():
        print("   %-34s %.6f" % (k, rate(v, r2)))

# torch.sum(a * b, dim=-1, keepdim=True): vdh = sum(h * view.unsqueeze(1)), ndl = sum(n.unsqueeze(1) * l) ...
for shape in ((512, 2048, 3), (98304, 16, 3), (98304, 1, 3)):
    hh = torch.randn(*shape, device=dev)
    vv = torch.randn(shape[0], 3, device=dev)
    prod = hh * vv.unsqueeze(1)
    r = torch.sum(prod, dim=-1, keepdim=True).reshape(-1)
    f = prod.reshape(-1, 3)
    p0, p1, p2 = f[:, 0], f[:, 1], f[:, 2]
    print("torch.sum(h * v.unsqueeze(1), dim=-1) on CUDA %s:  (p0+p1)+p2 %.6f   (p0+p2)+p1 %.6f   p0+(p1+p2) %.6f" % (
        shape, rate((p0 + p1) + p2, r), rate((p0 + p2) + p1, r), rate(p0 + (p1 + p2), r)))

a, b = torch.randn(n, 3, device=dev), torch.randn(n, 3, device=dev)
refc = torch.cross(a, b, dim=-1)
ay, az, bx, by, bz, ax = a[:, 1], a[:, 2], b[:, 0], b[:, 1], b[:, 2], a[:, 0]
cc = {
    "fma(ay,bz,-(az*by))": fma(ay, bz, -(az * by)),
    "fma(-az,by,ay*bz)": fma(-az, by, ay * bz),
    "ay*bz - az*by": ay * bz - az * by,
}
print("torch.cross x component on CUDA:")
for k, v in cc.items():
    print("   %-34s %.6f" % (k, rate(v, refc[:, 0])))

# division x / (norm + 1e-6): IEEE division?
d = ref + 1e-6
print("division: f32(x/d in float64) == torch-CUDA x/d: %.6f" % rate(f32(X.double() / d.double()), X / d))

# sin / cos / sqrt against the library's kernels
import ctypes
import texir_code_b200 as tx
lib = tx._lib.load()
phi = (torch.rand(n, device=dev) * 2 * np.pi - np.pi).contiguous()
if hasattr(lib, "texir_debug_sincos"):
    lib.texir_debug_sincos.restype = ctypes.c_int
    lib.texir_debug_sincos.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    for variant, name in ((0, "sincosf"), (1, "sinf / cosf")):
        s, c = torch.empty_like(phi), torch.empty_like(phi)
        rc = lib.texir_debug_sincos(phi.data_ptr(), n, variant, s.data_ptr(), c.data_ptr(), None)
        torch.cuda.synchronize()
        print("device %-12s == torch.sin: %.6f   == torch.cos: %.6f" % (name, rate(s, torch.sin(phi)), rate(c, torch.cos(phi))))
q = torch.rand(n, device=dev)
print("sqrt: f32(sqrt64) == torch.sqrt CUDA: %.6f" % rate(f32(torch.sqrt(q.double())), torch.sqrt(q)))

# compose: L = V*(sin*st) + N*c + U*(-(cos*st)) -- separate torch kernels, so separately rounded; check the whole chain
import texir_oracle as O
P = 4096
rng = np.random.default_rng(1)
nrm = rng.normal(size=(P, 3)).astype(np.float32)
nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
rough = rng.uniform(0.01, 0.8, size=(P, 1)).astype(np.float32)
torch.manual_seed(5)
sh = torch.rand(P, 1, 2)
for mode in ("uniform", "cosine", "importance"):
    for N in (64, 2048):
        ours = tx.generate_dir(torch.from_numpy(nrm).to(dev), N, None, mode, torch.from_numpy(rough).to(dev), shift=sh.to(dev))
        tc = O.generate_dir(torch.from_numpy(nrm).to(dev), N, None, mode, torch.from_numpy(rough).to(dev), shift=sh.to(dev))
        cpu = O.generate_dir(torch.from_numpy(nrm), N, None, mode, torch.from_numpy(rough), shift=sh)
        print("generate_dir %-10s N=%4d: ours vs torch-CUDA mismatching directions %.6f (max abs %.2e) | ours vs torch-CPU %.6f | "
              "torch-CUDA vs torch-CPU %.6f" % (mode, N, float((ours != tc).any(-1).float().mean()), float((ours - tc).abs().max()),
                                                float((ours.cpu() != cpu).any(-1).float().mean()),
                                                float((tc.cpu() != cpu).any(-1).float().mean())))
# frame only (normalised normal, U, V) through the first sample of the uniform mode at s0 = tiny
