"""pytest configuration: `gpu` marker, import paths, shared fixtures (goldens, oracle, CPU emulation)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpu_emul")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def oracle():
    import texir_oracle
    texir_oracle.build_oracle_lib()
    return texir_oracle


@pytest.fixture(scope="session")
def emul():
    import build as emul_build
    return emul_build.load()


@pytest.fixture(scope="session", params=["atlas", "smooth"])
def scene_golden(request):
    return request.param, load_golden("scene_" + request.param)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
