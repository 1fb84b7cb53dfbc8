"""NCCL check of the two multi-GPU paths, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/check_multi_gpu.py
  1. IrT: parallel.sharded_irradiance (point slices + one all_gather) == the single-GPU result, bit for bit.
  2. Mat stage: every rank renders its slice of the pixels, back-propagates into its own copy of materials_a /
     materials_r through the texture fetch's atomic scatter, then ONE all-reduce over both gradients
     (parallel.allreduce_gradients); compared with the gradients of all pixels computed on one GPU.
Prints one line per check on rank 0; exits non-zero on a mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texir_code_b200 import models, ops, parallel, synth  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sc = synth.make_room_scene(200000, seed=666, clutter=16)
    tex = torch.from_numpy(synth.make_radiance_texture(1024, seed=666))
    ok = True

    # 1. IrT
    P, N = 8192, 256
    pos, nrm = synth.sample_surface_points(sc, P, seed=3)
    torch.manual_seed(666)
    shift = torch.rand(P, 1, 2).to(dev)
    tracer = models.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, tex, sample_l=[N, 64], device=dev, build_threads=8)
    posd, nrmd = torch.from_numpy(pos).to(dev), torch.from_numpy(nrm).to(dev)
    full = tracer.irradiance(posd, nrmd, shift)
    shard = parallel.sharded_irradiance(tracer, posd, nrmd, shift)
    same = bool(torch.equal(full, shard))
    ok &= same
    if rank == 0:
        print("IrT: %d ranks, %d texels x %d spp, sharded == single GPU bit for bit: %s" % (world, P, N, same))

    # 2. Mat stage
    rays = torch.from_numpy(synth.cube_map_rays(32, cam=(0.2, -0.1, 0.3))).to(dev)
    dd = torch.from_numpy(synth.cube_map_ray_differentials(32)).to(dev)
    mm = models.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[32, 64], scene=tracer.scene, materials_a_res=256,
                              materials_r_res=512)
    rng = np.random.default_rng(1)
    with torch.no_grad():
        mm.materials_a.copy_(torch.from_numpy(rng.uniform(0.1, 0.9, (256, 256, 3)).astype(np.float32)))
        mm.materials_r.copy_(torch.from_numpy(rng.uniform(0.05, 0.7, (512, 512, 1)).astype(np.float32)))
    gb = ops.gbuffer(mm._scene(), rays, offset=1e-2, ray_differentials=dd)
    R = rays.shape[0]
    cam = torch.tensor([0.2, -0.1, 0.3], device=dev)
    irr = torch.full((R, 3), 2.0, device=dev)
    torch.manual_seed(5)
    sh = torch.rand(R, 1, 2).to(dev)
    g_rgb = torch.from_numpy(rng.normal(size=(R, 3)).astype(np.float32)).to(dev)

    def grads(b, e):
        mm.zero_grad(set_to_none=True)
        a = ops.texture_fetch(mm.materials_a, gb["texc"][b:e], gb["texd"][b:e], max_mip_level=13)
        r = ops.texture_fetch(mm.materials_r, gb["texc"][b:e], gb["texd"][b:e], max_mip_level=13)
        rgb = ops.shade(mm._scene(), gb["normal"][b:e], a, r, gb["position"][b:e], cam, irr[b:e], sh[b:e], num_spec=64)
        rgb.backward(g_rgb[b:e])

    grads(0, R)
    ref_a, ref_r = mm.materials_a.grad.clone(), mm.materials_r.grad.clone()
    b, e = parallel.shard_bounds(R, rank, world)
    grads(b, e)
    nbytes = parallel.allreduce_gradients([mm.materials_a, mm.materials_r])
    err_a = float(torch.linalg.norm(mm.materials_a.grad - ref_a) / torch.linalg.norm(ref_a))
    err_r = float(torch.linalg.norm(mm.materials_r.grad - ref_r) / torch.linalg.norm(ref_r))
    good = err_a < 1e-5 and err_r < 1e-5       # float atomics and the reduction add in a different order
    ok &= good
    if rank == 0:
        print("Mat stage: %d pixels split over %d ranks, one all-reduce of %d bytes; texture gradients vs single GPU: "
              "rel-L2 %.2e (materials_a) %.2e (materials_r): %s" % (R, world, nbytes, err_a, err_r, good))
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
