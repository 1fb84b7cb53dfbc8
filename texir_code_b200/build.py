"""Builds texir_code_b200/libtexir_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built
file travels with the repository snapshot to the GPU box).

    python -m texir_code_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libtexir_b200.so")
SOURCES = [os.path.join(CSRC, "texir_kernels.cu"), os.path.join(CSRC, "texir_post.cu"), os.path.join(CSRC, "texir_texture.cu"), os.path.join(CSRC, "bvh_build.cpp")]
HEADERS = [os.path.join(CSRC, h) for h in ("texir_math.h", "texir_bvh.h", "texir_walk.h", "texir_point.h", "texir_post.h", "texir_tex.h", "texir_host.h", "bvh_build.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "texir_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # host code must not contract a*b+c: the triangle precomputation has to round like the oracle
    "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread,-fvisibility=hidden,-O2",
    "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO + ".tmp"] + SOURCES
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    os.replace(SO + ".tmp", SO)
    return SO


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
