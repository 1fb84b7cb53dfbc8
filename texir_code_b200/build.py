"""Builds texir_code_b200/libtexir_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built
file travels with the repository snapshot to the GPU box).

    python -m texir_code_b200.build [--force] [-v]

Every source is compiled to its own object file (in parallel, objects cached under csrc/_obj and re-used
while neither the source nor any header changed), then linked.  Concurrent builders (one process per GPU
under torchrun on a fresh checkout) serialise on a file lock and each writes its own temporary file.
"""
from __future__ import annotations

import concurrent.futures
import fcntl
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(HERE, "libtexir_b200.so")
SOURCES = [os.path.join(CSRC, f) for f in (
    "texir_kernels.cu", "texir_post.cu", "texir_texture.cu", "texir_matstage.cu", "texir_peer.cu", "texir_rng.cu",
    "bvh_build_gpu.cu", "bvh_build.cpp")]
HEADERS = [os.path.join(CSRC, h) for h in (
    "texir_math.h", "texir_bvh.h", "texir_walk.h", "texir_point.h", "texir_post.h", "texir_tex.h", "texir_host.h",
    "texir_scene.h", "bvh_build.h")] + [os.path.join(os.path.dirname(HERE), "include", "texir_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # host code must not contract a*b+c: the triangle precomputation has to round like the oracle
    "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread,-fvisibility=hidden,-O2",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _sources():
    return [s for s in SOURCES if os.path.exists(s)]


def _newest_header() -> float:
    return max((os.path.getmtime(h) for h in HEADERS if os.path.exists(h)), default=0.0)


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in _sources()) or _newest_header() > t


def _compile_one(nvcc: str, src: str, obj: str, extra, verbose: bool) -> str:
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    return proc.stdout


def build_library(force: bool = False, verbose: bool = False, out: str = SO, defines=()) -> str:
    """compile + link; `defines` (e.g. ["-DTX_SPECULATIVE=1"]) builds a variant into `out` without the object cache"""
    if not force and out == SO and not is_stale():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)          # one builder at a time (torchrun starts one process per GPU)
        if not force and out == SO and not is_stale():
            return SO                             # another process built it while we waited
        nvcc = find_nvcc()
        hdr_t = _newest_header()
        tag = "" if not defines else "." + str(abs(hash(tuple(defines))) % 10 ** 8)
        jobs = []
        for src in _sources():
            obj = os.path.join(OBJ, os.path.basename(src) + tag + ".o")
            fresh = (not force and not defines and os.path.exists(obj)
                     and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_t))
            jobs.append((src, obj, fresh))
        with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
            futs = [pool.submit(_compile_one, nvcc, s, o, defines, verbose) for s, o, fresh in jobs if not fresh]
            for f in futs:
                log = f.result()
                if verbose and log:
                    print(log)
        tmp = "%s.%d.tmp" % (out, os.getpid())
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-pthread", "-o", tmp] \
            + [o for _, o, _ in jobs]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
        os.replace(tmp, out)
        if defines:
            for _, o, _ in jobs:
                try:
                    os.remove(o)
                except OSError:
                    pass
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
