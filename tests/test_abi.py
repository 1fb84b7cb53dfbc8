"""CPU: the C-ABI library loads and exports every symbol include/texir_b200.h declares; argument validation
that happens before any CUDA call; the Python binding types every declared symbol.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "texir_b200.h")).read()
    return sorted(set(re.findall(r"TEXIR_API\s+[\w\s\*]+?\b(texir_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from texir_code_b200 import _lib, build
    build.build_library()
    return _lib.load()


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("texir_scene_create", "texir_scene_set_texture", "texir_cast_rays", "texir_generate_dir",
                 "texir_query_irf", "texir_irradiance_fwd", "texir_shade_fwd", "texir_shade_bwd",
                 "texir_specular_reflectance", "texir_diffuse_reflectance", "texir_texture_scatter"):
        assert must in syms
    assert len(syms) >= 22


def test_library_exports_every_declared_symbol(lib):
    from texir_code_b200 import _lib
    raw = ctypes.CDLL(_lib.SO_PATH)
    for s in declared_symbols():
        assert hasattr(raw, s), "libtexir_b200.so lacks %s" % s
    # and the binding types exactly the declared set
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert lib.texir_version() >= 100
    assert lib.texir_launch_count() >= 0


def test_no_cpp_symbols_leak(lib):
    import subprocess
    from texir_code_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("texir_") for s in exported), exported


def test_argument_validation_without_gpu(lib):
    z = ctypes.c_void_p(0)
    assert lib.texir_generate_dir(z, z, z, 4, 0, 0, z, z) != 0        # N = 0
    assert b"N > 0" in lib.texir_last_error()
    assert lib.texir_generate_dir(z, z, z, 4, 16, 7, z, z) != 0       # unknown mode
    assert lib.texir_generate_dir(z, z, z, -1, 16, 0, z, z) != 0      # negative P
    assert lib.texir_query_irf(z, z, z, 1, 1, z, z) != 0              # NULL scene
    assert b"NULL scene" in lib.texir_last_error()
    assert lib.texir_irradiance_fwd(z, z, z, z, z, 1, 1, 0, z, z) != 0
    assert lib.texir_point_order(z, z, 4, z, z) != 0
    assert lib.texir_scene_set_bvh(z, 8) != 0
    h = ctypes.c_void_p()
    assert lib.texir_scene_create(z, 0, z, 0, z, 0, 1, ctypes.byref(h)) != 0   # empty mesh, fails before CUDA
    assert b"empty mesh" in lib.texir_last_error()
    assert lib.texir_scene_destroy(z) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "texir_code_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # comments may mention the test infrastructure; code must not import / include / dlopen it
                code = "\n".join(l for l in src.splitlines()
                                 if ("import" in l or "#include" in l or "CDLL" in l or "dlopen" in l))
                assert "oracle" not in code and "cpu_emul" not in code and "emul" not in code, f


def test_ops_fail_loudly_on_cpu_tensors():
    import torch
    from texir_code_b200 import TexirError, ops, sample_util
    with pytest.raises(TexirError):
        sample_util.generate_dir(torch.zeros(4, 3), 8)
    with pytest.raises(TexirError):
        ops.specular_reflectance(*(torch.zeros(2, 4, 3),) * 2, torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 4, 3),
                                 torch.zeros(2, 1))
    if not torch.cuda.is_available():
        from texir_code_b200 import Scene
        with pytest.raises(TexirError):
            Scene([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]])
