"""The reference's two stages end to end on the synthetic room, with this library in place of Open3D / nvdiffrast / cv2:

    1. IrT generation   (trainer/generate_ir_texture.py + models/tracer_o3d_irt.py): one fused launch over all texels,
                        written as a Radiance .hdr, padded like tools/padding_texture.py, read back
    2. Mat stage        (trainer/train_material.py + models/mat_nvdiffrast.py): MaterialModel.forward(mvp, id, cam, stage)
                        with the reference's stages 0 -> 1 -> 2, Adam on materials_a / materials_r, log-L1 image loss,
                        the trainer's clamps

    python examples/pipeline_synthetic.py [--tris 200000] [--irt 256] [--spp 256] [--steps 30]
Needs a CUDA device; prints timings and the loss curve."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import texir_code_b200 as tx  # noqa: E402
from texir_code_b200 import postprocess, synth  # noqa: E402
from texir_code_b200.graphed import GraphedMatStep  # noqa: E402


def cube_mvps(cam):
    """mvp [6,4,4] of the six 90-degree views around `cam` in the reference's row-vector convention"""
    n, f = 1e-4, 100.0
    proj = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, (f + n) / (f - n), -(2 * f * n) / (f - n)], [0, 0, 1.0, 0]])
    out = []
    for fwd, up in (((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 1, 0), (0, 0, -1)), ((0, -1, 0), (0, 0, 1)),
                    ((0, 0, 1), (0, 1, 0)), ((0, 0, -1), (0, 1, 0))):
        fz, uy = torch.tensor(fwd, dtype=torch.float32), torch.tensor(up, dtype=torch.float32)
        w2c = torch.eye(4)
        w2c[:3, :3] = torch.stack([torch.linalg.cross(uy, fz), uy, fz])
        w2c[:3, 3] = -w2c[:3, :3] @ cam
        out.append((proj @ w2c).t())
    return torch.stack(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=200000)
    ap.add_argument("--irt", type=int, default=256, help="IrT resolution (the reference uses 1024)")
    ap.add_argument("--spp", type=int, default=256, help="IrT samples per texel (the reference uses 2048)")
    ap.add_argument("--steps", type=int, default=30, help="optimiser steps per stage")
    ap.add_argument("--graphed", action="store_true",
                    help="forward + loss + backward of every stage replayed as ONE CUDA graph (texir_code_b200.GraphedMatStep)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sc = synth.make_room_scene(args.tris, seed=666, clutter=16)
    tex = torch.from_numpy(synth.make_radiance_texture(1024, seed=666))

    # ---- stage IrrT ----------------------------------------------------------------------------------------------
    tracer = tx.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, tex, sample_l=[args.spp, 64], device=dev)
    P = args.irt * args.irt
    pos, nrm = synth.sample_surface_points(sc, P, seed=3, seam_fraction=0.2)      # 20 % empty texels, like an atlas
    position_texture = torch.from_numpy(pos).to(dev).reshape(args.irt, args.irt, 3)
    normal_texture = torch.from_numpy(nrm).to(dev).reshape(args.irt, args.irt, 3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    irr_texture = tracer(position_texture, normal_texture)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("IrT %dx%d x %d spp: %.1f ms (%.2e ray-samples/s)" % (args.irt, args.irt, args.spp, (t1 - t0) * 1e3,
                                                              P * args.spp / (t1 - t0)))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "0_irr_texture.hdr")
        postprocess.save_irr_texture(irr_texture, path)                            # generate_ir_texture.py:80-82
        padded_path = postprocess.padding_texture(path)                            # tools/padding_texture.py
        irrt = postprocess.imread_hdr(padded_path, dev).flip(-1)                   # mat_nvdiffrast.py:73 ([:,:,::-1])
        print("wrote %s (%d bytes), padded -> %s; empty texels before %d, after %d" % (
            os.path.basename(path), os.path.getsize(path), os.path.basename(padded_path),
            int((irr_texture.sum(-1) == 0).sum()), int((irrt.sum(-1) == 0).sum())))

    # ---- stage Mat -----------------------------------------------------------------------------------------------
    cam = torch.tensor([0.3, -0.2, 0.4])
    mvp = cube_mvps(cam).to(dev)
    model = tx.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[32, 64], scene=tracer.scene, materials_a_res=256,
                             materials_r_res=512, cube_res=64, irrt=irrt, max_mip_level=13)
    # the "photograph": the same model with ground-truth materials
    rng = np.random.default_rng(1)
    gt_a = torch.from_numpy(rng.uniform(0.2, 0.9, (256, 256, 3)).astype(np.float32)).to(dev)
    gt_r = torch.from_numpy(rng.uniform(0.1, 0.6, (512, 512, 1)).astype(np.float32)).to(dev)
    init_a, init_r = model.materials_a.detach().clone(), model.materials_r.detach().clone()
    with torch.no_grad():
        model.materials_a.copy_(gt_a)
        model.materials_r.copy_(gt_r)
        torch.manual_seed(7)
        target = model(mvp, 0, cam, 2)["rgb"]
        model.materials_a.copy_(init_a)
        model.materials_r.copy_(init_r)
    opt = torch.optim.Adam([model.materials_a, model.materials_r], lr=2e-2)
    loss_fn = lambda out, gt: (torch.log1p(out["rgb"]) - torch.log1p(gt)).abs().mean()   # models/loss.py:81-99
    for stage in (0, 1, 2):                                                         # train_material.py's schedule
        losses = []
        warm = min(3, args.steps - 1)                                               # first steps load kernels / grow pools
        step = None
        if args.graphed:
            # captured once per stage; the gradients live in the graph's memory and every replay overwrites them, so the
            # loop below must not call zero_grad(set_to_none=True)
            step = GraphedMatStep(model, stage, loss_fn, example=(mvp, cam, target))
        for it in range(args.steps):
            if it == warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            torch.manual_seed(7)                                                    # same sample set as the target
            if step is not None:
                loss = step(mvp, cam, target)
            else:
                opt.zero_grad(set_to_none=True)
                out = model(mvp, 0, cam, stage)
                loss = loss_fn(out, target)
                loss.backward()
            opt.step()
            with torch.no_grad():
                model.materials_r.clamp_(0.01, 0.8)                                # train_material.py:458
                model.materials_a.clamp_(min=0.0)
            losses.append(float(loss.detach()))
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / (args.steps - warm)
        print("stage %d: %d steps, %.2f ms/step, loss %.4f -> %.4f" % (stage, args.steps, ms, losses[0], losses[-1]))


if __name__ == "__main__":
    main()
