"""A small pass over the kernels added in round 2, for compute-sanitizer (memcheck / racecheck):
device BVH build, G-buffer (free-running walk, tiles), the Mat-stage step (fetch, render, backward into texture space,
pyramid build / fold), the device-side mt19937 shift stream (eager and through a CUDA graph), RGBE texture, padding.
    compute-sanitizer --tool memcheck python profiles/r02_sanitizer_step.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_host_logic import _reference_cube_mvps
from texir_code_b200 import Scene, graphed, models, ops, postprocess, sample_util, synth
dev = torch.device("cuda:0")
sc = synth.make_room_scene(6000, seed=3, tex_res=128, clutter=6)
tex = torch.from_numpy(synth.make_radiance_texture(128))
scene = Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev, builder="device")
mvp, cam = _reference_cube_mvps(np.eye(4, dtype=np.float32))
mvp, cam = mvp.to(dev), cam.to(dev)
mm = models.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[32, 16], scene=scene, materials_a_res=64,
                          materials_r_res=128, cube_res=96, irrt=torch.rand(64, 64, 3) + 0.5)   # 55 296 pixels: device shift
mm._warned_normals = True
gt = torch.rand(6, 96, 96, 3, device=dev)
for stage in (1, 2):
    mm.zero_grad(set_to_none=True)
    ((mm(mvp, 0, cam, stage)["rgb"] - gt) ** 2).mean().backward()
torch.cuda.synchronize()
step = graphed.GraphedMatStep(mm, 2, lambda o, g: ((o["rgb"] - g) ** 2).mean(), example=(mvp, cam, gt), warmup=1)
for _ in range(2):
    loss = step(mvp, cam, gt)
torch.cuda.synchronize()
pos, nrm = synth.sample_surface_points(sc, 2048, seed=1)
e = ops.irradiance(scene, torch.from_numpy(pos).to(dev), torch.from_numpy(nrm).to(dev), torch.rand(2048, 1, 2).to(dev), 64)
img = torch.rand(40, 56, 3, device=dev)
img[torch.rand(40, 56, device=dev) < 0.5] = 0
postprocess.pad_texture(img)
torch.cuda.synchronize()
print("ok", float(loss), float(e.sum()), sample_util._mt_ok)
