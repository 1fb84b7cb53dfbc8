"""CPU: the drop-in boundary from the consumers' side.

  * a C program compiled against include/texir_b200.h and linked with libtexir_b200.so (argument validation only:
    no GPU needed) -- the header, the library and INTEGRATION.md section C cannot drift apart unnoticed;
  * the reference's OWN model class (imported from /root/reference with the third-party modules stubbed, in a
    subprocess) with HotPathMixin put in front of it, as INTEGRATION.md patch B does: the four hot-path methods
    must resolve to the mixin and keep the reference's signatures;
  * the config-level seam: ConfigLite on the reference's .conf files, the OBJ reader, the constructor signatures.
"""
import inspect
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

C_PROGRAM = r"""
#include <stdio.h>
#include <string.h>
#include "texir_b200.h"

/* INTEGRATION.md section C, with NULL device pointers: every call must be REJECTED by argument validation */
int main(void) {
    texir_scene* sc = NULL;
    int fails = 0;
    void* stream = NULL;
    if (texir_version() < 100) { printf("version\n"); fails++; }
    if (texir_scene_create(NULL, 0, NULL, 0, NULL, 0, 0, &sc) == 0) { printf("create accepted an empty mesh\n"); fails++; }
    if (!strstr(texir_last_error(), "empty mesh")) { printf("error text: %s\n", texir_last_error()); fails++; }
    if (texir_scene_create_ex(NULL, 0, NULL, 0, NULL, 0, 0, TEXIR_BUILDER_DEVICE, &sc) == 0) fails++;
    if (texir_scene_set_texture(NULL, NULL, 4, 4, TEXIR_TEX_CHW, 0, stream) == 0) fails++;
    if (texir_irradiance_fwd(NULL, NULL, NULL, NULL, /*order*/ NULL, 512, 2048, TEXIR_MODE_UNIFORM, NULL, stream) == 0) fails++;
    if (!strstr(texir_last_error(), "NULL scene")) { printf("error text: %s\n", texir_last_error()); fails++; }
    float* peers[2] = {NULL, NULL};
    if (texir_irradiance_fwd_peers(NULL, NULL, NULL, NULL, NULL, 512, 2048, TEXIR_MODE_UNIFORM, peers, 2, stream) == 0) fails++;
    if (texir_shade_fwd(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 16, 256, 0, TEXIR_MODE_IMPORTANCE,
                        TEXIR_MODE_UNIFORM, 1e-14f, NULL, NULL, NULL, stream) == 0) fails++;
    if (texir_shade_bwd(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 16, 256, 0, TEXIR_MODE_IMPORTANCE,
                        TEXIR_MODE_UNIFORM, 1e-14f, NULL, NULL, NULL, NULL, NULL, NULL, stream) == 0) fails++;
    if (texir_generate_dir(NULL, NULL, NULL, 4, 0, TEXIR_MODE_UNIFORM, NULL, stream) == 0) fails++;
    if (texir_texture_mip_levels(4096, 4096, 13) != 12) { printf("mip levels\n"); fails++; }
    if (texir_texture_mip_levels(6, 6, -1) != -1) { printf("odd side must be refused\n"); fails++; }
    if (texir_texture_mip_elems(4, 4, 3, 2) != (2 * 2 + 1) * 3) { printf("mip elems\n"); fails++; }
    if (texir_peer_barrier(NULL, 0, 2, 1, 0, stream) == 0) fails++;
    if (texir_scene_destroy(NULL) != 0) fails++;
    texir_scene_info info;
    if (sizeof(info.builder) != 4 || TEXIR_MAX_PEERS != 8 || TEXIR_PEER_HANDLE_BYTES != 64) fails++;
    printf(fails ? "FAILED %d\n" : "OK\n", fails);
    return fails;
}
"""


def test_c_consumer_compiles_against_the_header_and_links(tmp_path):
    from texir_code_b200 import _lib, build
    build.build_library()
    src = tmp_path / "consumer.c"
    src.write_text(C_PROGRAM)
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.SO_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:libtexir_b200.so", "-Wl,-rpath," + libdir]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.strip().endswith("OK"), run.stdout + run.stderr


def test_integration_md_examples_use_the_declared_signatures():
    """the C snippet of INTEGRATION.md passes as many arguments as the header declares for each call it shows"""
    import re
    hdr = open(os.path.join(ROOT, "include", "texir_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("## C. Calling the C ABI"):]
    block = block[block.index("```c"):block.index("```", block.index("```c") + 4)]

    def nargs(text, name):
        m = re.search(name + r"\s*\(", text)
        if not m:
            return None
        depth, i, n, seen = 1, m.end(), 1, False
        while depth:
            ch = text[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 1:
                n += 1
            elif not ch.isspace():
                seen = True
            i += 1
        return n if seen else 0

    for name in ("texir_scene_create", "texir_scene_set_texture", "texir_irradiance_fwd", "texir_shade_fwd"):
        assert nargs(block, name) == nargs(hdr[hdr.index("TEXIR_API int " + name + "("):], name), name


MRO_SCRIPT = r"""
import inspect, json, sys, types
sys.path.insert(0, %(ref)r)
sys.path.insert(0, %(root)r)
import torch
class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)
    def __call__(self, *a, **k):
        return a[0] if a else None
for m in ["open3d", "nvdiffrast", "nvdiffrast.torch", "pyredner", "pyhocon", "siren_pytorch", "tinycudann",
          "tensorboardX", "GPUtil", "pytorch_msssim"]:
    sys.modules[m] = _Stub(m)
import models.mat_nvdiffrast as M
import models.tracer_o3d_irt as T
import utils.sample_util as su
from texir_code_b200.models import HotPathMixin
from texir_code_b200 import models as ours, sample_util as ours_su, conf_models

class Patched(HotPathMixin, M.MaterialModel):        # INTEGRATION.md patch B
    pass

res = {"mro_ok": {}, "sig_ok": {}}
for name in ("render", "query_irf", "specular_reflectance", "diffuse_reflectance"):
    res["mro_ok"][name] = getattr(Patched, name) is getattr(HotPathMixin, name)
    a = list(inspect.signature(getattr(M.MaterialModel, name)).parameters)
    b = list(inspect.signature(getattr(HotPathMixin, name)).parameters)
    res["sig_ok"][name] = a == b
# generate_dir keeps the reference's positional parameters (one extra keyword at the end)
a = list(inspect.signature(su.generate_dir).parameters)
b = list(inspect.signature(ours_su.generate_dir).parameters)
res["generate_dir"] = b[:len(a)] == a
# the config-level seam: same constructor / forward parameters as the reference classes
def params(f):
    return [p for p in inspect.signature(f).parameters if p != "self"]
res["ctor_mat"] = params(conf_models.MaterialModel.__init__)[:6] == params(M.MaterialModel.__init__)
res["ctor_irt"] = params(conf_models.TracerO3d.__init__)[:5] == params(T.TracerO3d.__init__)
res["fwd_mat"] = params(ours.MaterialModel.forward) == params(M.MaterialModel.forward)
# TracerO3d.forward() takes no required argument, like the reference's
res["fwd_irt"] = all(p.default is not inspect.Parameter.empty for n, p in inspect.signature(conf_models.TracerO3d.forward).parameters.items() if n != "self")
res["state_keys"] = sorted(k for k in ("materials_a", "materials_r", "irrt"))
print("RESULT " + json.dumps(res))
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_mixin_in_front_of_the_reference_class():
    script = MRO_SCRIPT % {"ref": REF, "root": ROOT}
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert all(res["mro_ok"].values()), res
    assert all(res["sig_ok"].values()), res
    assert res["generate_dir"] and res["ctor_mat"] and res["ctor_irt"] and res["fwd_mat"] and res["fwd_irt"], res


def test_config_lite_reads_the_hocon_subset():
    from texir_code_b200.conf_models import ConfigLite
    c = ConfigLite.parse_string(textwrap.dedent("""
        train{
            expname = default   # a comment
            model_class = texir_code_b200.conf_models.MaterialModel
            pano_img_res = [256,512]
            sample_light = [32, 16]
            hdr_exposure = 5
            mat_learning_rate = 3e-2
            optim_cam = False
            path_mesh_open3d = ../data/x/out1.obj
        }
        render_loss
        {
            loss_type = L1
            w_gradient = 1
        }
        models{ render{ sample_type = [ uniform, importance] } tracer{ } }
    """))
    assert c.get_string("train.model_class") == "texir_code_b200.conf_models.MaterialModel"
    assert c.get_list("train.pano_img_res") == [256, 512] and c.get_list("train.sample_light") == [32, 16]
    assert c.get_float("train.hdr_exposure") == 5.0 and c.get_float("train.mat_learning_rate") == 0.03
    assert c.get_bool("train.optim_cam") is False
    assert c.get_list("models.render.sample_type") == ["uniform", "importance"]
    assert c.get_config("render_loss") == {"loss_type": "L1", "w_gradient": 1} and c.get_config("models.tracer") == {}
    assert c.get_list("train.env_res", default=[8, 16]) == [8, 16]
    with pytest.raises(KeyError):
        c.get_string("train.nope")
    if os.path.isdir(REF):
        import glob
        for f in glob.glob(os.path.join(REF, "configs", "*.conf")):
            conf = ConfigLite.parse_file(f)
            top = "train" if "train" in conf else "test"          # the tester confs use a `test` block
            assert conf.get_string(top + ".model_class").startswith("models.") and conf.get_string(top + ".path_mesh_open3d").endswith(".obj")


def test_obj_reader(tmp_path):
    from texir_code_b200.conf_models import load_obj
    p = tmp_path / "out1.obj"
    p.write_text(textwrap.dedent("""
        # a quad and a triangle
        v 0 0 0
        v 1 0 0
        v 1 1 0
        v 0 1 0
        v 0 0 1
        vt 0 0
        vt 1 0
        vt 1 1
        vt 0 1
        vn 0 0 1
        vn 0 1 0
        f 1/1/1 2/2/1 3/3/1 4/4/1
        f -5/1/2 -1/2/2 2/3/2
    """))
    m = load_obj(str(p))
    assert m["vertices"].shape == (5, 3) and m["vertices"].dtype == np.float32
    assert m["triangles"].tolist() == [[0, 1, 2], [0, 2, 3], [0, 4, 1]] and m["triangles"].dtype == np.int32
    assert m["triangle_uvs"].shape == (9, 2) and m["triangle_uvs"].dtype == np.float64
    assert m["triangle_uvs"][:3].tolist() == [[0, 0], [1, 0], [1, 1]] and m["triangle_uvs"][5].tolist() == [0, 1]   # raw v
    assert m["vertex_normals"].shape == (5, 3) and m["vertex_normals"][4].tolist() == [0, 1, 0]
