"""make_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by executing the REFERENCE's own
Python (from /root/reference, read-only) in the build container.

The reference cannot run stock here (open3d, nvdiffrast, pyredner, pyhocon ... are absent), but its
hot-path functions do once:
  * the absent third-party modules are registered as permissive stubs in sys.modules,
  * torch.Tensor.cuda is shimmed to identity (the reference hard-codes .cuda()),
  * `self.scene` is an oracle/texir_oracle.OracleScene, whose cast_rays follows Open3D's contract.
So every number below comes out of utils/sample_util.generate_dir,
models/mat_nvdiffrast.MaterialModel.{render,specular_reflectance,diffuse_reflectance,query_irf},
models/test_nvdiffrast.MaterialModel.render and the body of models/tracer_o3d_irt.TracerO3d.forward,
verbatim -- only the ray/triangle closest hit is ours.

Run:  python oracle/make_golden.py      (needs /root/reference; the GPU box never runs this)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("TEXIR_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def import_reference():
    """returns (sample_util, mat_nvdiffrast, tracer_o3d_irt, test_nvdiffrast) reference modules"""
    if not os.path.isdir(REF):
        raise RuntimeError("reference checkout not found at %s" % REF)
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self

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
    import utils.sample_util as su
    import models.mat_nvdiffrast as M
    import models.tracer_o3d_irt as T
    import models.test_nvdiffrast as TN
    return su, M, T, TN


def small_scene(num_tris=3000, tex_res=128, uv_kind="atlas", tex_kind="noise"):
    from texir_code_b200 import synth
    sc = synth.make_room_scene(num_tris, seed=666, clutter=6, tex_res=tex_res, uv_kind=uv_kind)
    tex = synth.make_radiance_texture(tex_res, seed=666, hdr_exposure=5.0, kind=tex_kind)
    return sc, tex


def fake_self(M, scene, triangle_uvs, texture, sample_l, sample_type, cube_res, model_cls=None):
    cls = model_cls or M.MaterialModel
    fake = types.SimpleNamespace(sample_l=list(sample_l), sample_type=list(sample_type), scene=scene,
                                 triangle_uvs=triangle_uvs, texture=texture, cube_res=cube_res, relighting=False)
    fake.query_irf = lambda *a: cls.query_irf(fake, *a)
    fake.specular_reflectance = lambda *a, **k: cls.specular_reflectance(fake, *a, **k)
    fake.diffuse_reflectance = lambda *a, **k: cls.diffuse_reflectance(fake, *a, **k)
    return fake


def main():
    import texir_oracle as O
    from texir_code_b200 import synth
    su, M, T, TN = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)

    # ---- 1. Hammersley known answers -----------------------------------------------------------
    ham = np.array([su.Hammersley(i, 16) for i in range(16)], np.float64)
    ri = np.array([su.RadicalInverse(i) for i in [0, 1, 2, 3, 5, 255, 2047, 65535, 2 ** 20 + 1]], np.float64)
    np.savez_compressed(os.path.join(OUT, "hammersley.npz"), ham16=ham,
                        ri_in=np.array([0, 1, 2, 3, 5, 255, 2047, 65535, 2 ** 20 + 1]), ri_out=ri)

    # ---- 2. generate_dir, three modes, forward + d/d(roughness) ---------------------------------
    rng = np.random.default_rng(11)
    nrm = rng.normal(size=(40, 3)).astype(np.float32)
    nrm[:20] /= np.linalg.norm(nrm[:20], axis=1, keepdims=True)       # unit and non-unit normals
    nrm[0] = [0, 1, 0]; nrm[1] = [1, 0, 0]; nrm[2] = [0.995, 0.05, 0.05]; nrm[3] = [0, 0, 0]; nrm[4] = [-1, 0, 0]
    rough = rng.uniform(0.01, 0.8, size=(40, 1)).astype(np.float32)
    gd = {"normals": nrm, "roughness": rough}
    for N in (16, 64):
        for mode in ("uniform", "cosine", "importance"):
            torch.manual_seed(666)
            shift = torch.rand(40, 1, 2)
            torch.manual_seed(666)
            r = torch.tensor(rough, requires_grad=True)
            L = su.generate_dir(torch.tensor(nrm), N, None, mode, r if mode == "importance" else None)
            gd["shift"] = shift.numpy()
            gd["L_%s_%d" % (mode, N)] = L.detach().numpy()
            if mode == "importance":
                g = torch.tensor(rng.normal(size=L.shape).astype(np.float32))
                (L * g).sum().backward()
                gd["gL_%d" % N] = g.numpy()
                gd["grough_%d" % N] = r.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "generate_dir.npz"), **gd)

    # ---- 3. query_irf known answers on the 2-triangle quad ------------------
    qv = np.array([[-1, -1, 2], [1, -1, 2], [1, 1, 2], [-1, 1, 2]], np.float32)
    qt = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    quv = np.array([[0, 0], [1, 0], [1, 1], [0, 0], [1, 1], [0, 1]], np.float64)
    qtex = torch.arange(48, dtype=torch.float32).reshape(4, 4, 3).permute(2, 0, 1).unsqueeze(0).contiguous()
    qs = O.OracleScene(qv, qt, brute=True)
    org = np.array([[0, 0, 0], [0, 0, 0], [.5, .5, 0], [0, 0, 0], [0, 0, 0], [0.9, -0.9, 0]], np.float32)
    dr = np.array([[0, 0, 1], [.2, .1, 1], [0, 0, 1], [0, 0, -1], [0, 0, 0], [0.05, -0.05, 1]], np.float32)
    fake = fake_self(T, qs, quv, qtex, [1, 1], ["uniform", "importance"], 1, model_cls=T.TracerO3d)
    rad = T.TracerO3d.query_irf(fake, torch.tensor(org).unsqueeze(1), torch.tensor(dr).reshape(6, 1, 1, 3), 1)
    np.savez_compressed(os.path.join(OUT, "query_irf_quad.npz"), vertices=qv, triangles=qt, triangle_uvs=quv,
                        texture=qtex.numpy(), org=org, dir=dr, rad=rad.numpy())

    # ---- 4. scene-level goldens: query_irf, specular/diffuse reductions, render fwd+bwd, IrT ----
    for tag, uv_kind, tex_kind in (("atlas", "atlas", "noise"), ("smooth", "planar", "smooth")):
        sc, tex = small_scene(uv_kind=uv_kind, tex_kind=tex_kind)
        ttex = torch.tensor(tex)
        scene = O.OracleScene(sc.vertices, sc.triangles, brute=True)
        cube_res = 4
        P = 6 * cube_res * cube_res
        pos, nrm = synth.sample_surface_points(sc, P, seed=5)
        rough, albedo, irr, g_rgb = synth.make_materials(P, seed=5)
        cam = np.array([0.3, -0.2, 0.5], np.float32)
        out = {"vertices": sc.vertices, "triangles": sc.triangles, "triangle_uvs": sc.triangle_uvs, "texture": tex,
               "pos": pos, "nrm": nrm, "rough": rough, "albedo": albedo, "irr": irr, "g_rgb": g_rgb, "cam": cam}
        for N, eps_tag, cls in ((16, "e14", M.MaterialModel), (64, "e6", TN.MaterialModel)):
            fake = fake_self(M, scene, sc.triangle_uvs, ttex, [N, N], ["uniform", "importance"], cube_res, cls)
            shp = (6, cube_res, cube_res)
            t_n = torch.tensor(nrm).reshape(*shp, 3)
            t_a = torch.tensor(albedo).reshape(*shp, 3).requires_grad_(True)
            t_r = torch.tensor(rough).reshape(*shp, 1).requires_grad_(True)
            t_p = torch.tensor(pos).reshape(*shp, 3)
            t_i = torch.tensor(irr).reshape(*shp, 3).requires_grad_(True)
            torch.manual_seed(666)
            shift = torch.rand(P, 1, 2)
            torch.manual_seed(666)
            res = cls.render(fake, t_n, t_a, t_r, t_p, torch.tensor(cam), t_i)
            (res["rgb"] * torch.tensor(g_rgb).reshape(*shp, 3)).sum().backward()
            k = "render_%s_%d_" % (eps_tag, N)
            out[k + "shift"] = shift.numpy()
            out[k + "rgb"] = res["rgb"].detach().numpy().reshape(P, 3)
            out[k + "g_albedo"] = t_a.grad.numpy().reshape(P, 3)
            out[k + "g_rough"] = t_r.grad.numpy().reshape(P)
            out[k + "g_irr"] = t_i.grad.numpy().reshape(P, 3)
            # the pieces, for the stand-alone ops: h, l, lighting, specular(lighting, ...)
            torch.manual_seed(666)
            h = su.generate_dir(torch.tensor(nrm), N, None, "importance", torch.tensor(rough))
            view = torch.nn.functional.normalize(torch.tensor(cam).unsqueeze(0) - torch.tensor(pos), eps=1e-4)
            vdh = torch.clamp(torch.sum(h * view.unsqueeze(1), dim=-1, keepdim=True), 0.0, 1.0)
            l = 2 * vdh * h - view.unsqueeze(1)
            lighting = fake.query_irf(torch.tensor(pos).unsqueeze(1).expand_as(l), l.unsqueeze(-2), N)
            out[k + "h"] = h.numpy()
            out[k + "l"] = l.numpy()
            out[k + "view"] = view.numpy()
            out[k + "lighting"] = lighting.numpy()
            Lg, hg, lg = lighting.clone().requires_grad_(True), h.clone().requires_grad_(True), l.clone().requires_grad_(True)
            rg = torch.tensor(rough).requires_grad_(True)
            spec = cls.specular_reflectance(fake, Lg, hg, torch.tensor(nrm), view, lg, rg)
            (spec * torch.tensor(g_rgb)).sum().backward()
            out[k + "spec"] = spec.detach().numpy()
            out[k + "spec_gL"] = Lg.grad.numpy()
            out[k + "spec_gh"] = hg.grad.numpy()
            out[k + "spec_gl"] = lg.grad.numpy()
            out[k + "spec_grough"] = rg.grad.numpy().reshape(P)
        # relighting render (test_nvdiffrast.py:275-295): MC diffuse + specular, eps 1e-6
        for dmode in ("uniform", "cosine"):
            fake = fake_self(M, scene, sc.triangle_uvs, ttex, [32, 16], [dmode, "importance"], cube_res, TN.MaterialModel)
            fake.relighting = True
            shp = (6, cube_res, cube_res)
            torch.manual_seed(666)
            sh_d = torch.rand(P, 1, 2)
            sh_s = torch.rand(P, 1, 2)
            torch.manual_seed(666)
            res = TN.MaterialModel.render(fake, torch.tensor(nrm).reshape(*shp, 3), torch.tensor(albedo).reshape(*shp, 3),
                                          torch.tensor(rough).reshape(*shp, 1), torch.tensor(pos).reshape(*shp, 3),
                                          torch.tensor(cam), torch.tensor(irr).reshape(*shp, 3))
            out["relight_%s_shift_diff" % dmode] = sh_d.numpy()
            out["relight_%s_shift_spec" % dmode] = sh_s.numpy()
            out["relight_%s_rgb" % dmode] = res["rgb"].numpy().reshape(P, 3)
            # diffuse_reflectance on its own
            torch.manual_seed(666)
            ld = su.generate_dir(torch.tensor(nrm), 32, None, dmode)
            rad = fake.query_irf(torch.tensor(pos).unsqueeze(1).expand_as(ld), ld.unsqueeze(-2), 32)
            out["diffuse_%s_l" % dmode] = ld.numpy()
            out["diffuse_%s_lighting" % dmode] = rad.numpy()
            out["diffuse_%s_out" % dmode] = TN.MaterialModel.diffuse_reflectance(fake, rad, ld, torch.tensor(nrm),
                                                                                 torch.tensor(albedo), dmode).numpy()
        # IrT loop body (tracer_o3d_irt.py:165-173) on 512-point batches, 128 spp
        Pi = 1024
        ipos, inrm = synth.sample_surface_points(sc, Pi, seed=9, seam_fraction=0.02)
        fake = fake_self(T, scene, sc.triangle_uvs, ttex, [128, 128], ["uniform", "importance"], 0, T.TracerO3d)
        torch.manual_seed(666)
        shifts = torch.cat([torch.rand(512, 1, 2) for _ in range(Pi // 512)])
        torch.manual_seed(666)
        irr_out = np.zeros((Pi, 3), np.float32)
        batch = 512
        for i in range(Pi // batch):
            position = torch.tensor(ipos[i * batch:(i + 1) * batch])
            normal = torch.tensor(inrm[i * batch:(i + 1) * batch])
            ldir = su.generate_dir(normal, 128, None, "uniform", None)
            gt_ir = T.TracerO3d.query_irf(fake, position.unsqueeze(1).expand_as(ldir), ldir.unsqueeze(-2), 128)
            ndl = torch.clamp(torch.sum(normal.unsqueeze(1) * ldir.reshape(batch, -1, 3), dim=-1, keepdim=True), 0.0, 1.0)
            irr_out[i * batch:(i + 1) * batch] = (torch.sum(gt_ir.reshape(batch, -1, 3) * ndl, dim=1) * 2 * np.pi / 128).numpy()
        out.update({"irt_pos": ipos, "irt_nrm": inrm, "irt_shift": shifts.numpy(), "irt_N": np.array(128),
                    "irt_E": irr_out})
        np.savez_compressed(os.path.join(OUT, "scene_%s.npz" % tag), **out)
        print("wrote scene_%s.npz" % tag)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
