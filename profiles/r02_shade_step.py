"""One Mat-stage forward + backward at the configs[2] shape (98 304 pixels x 256 spp), stash then recompute mode -- the
command the ncu captures of k_shade_lean are taken on:   python profiles/r02_shade_step.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from texir_code_b200 import Scene, ops, synth
dev = torch.device("cuda:0")
sc = synth.make_room_scene(1 << 20, seed=666)
tex = torch.from_numpy(synth.make_radiance_texture(4096, seed=666))
scene = Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev)
scene.set_texture(tex)
pos, nrm, rough, albedo, irr, g_rgb = bench.mat_inputs(scene, sc, dev)
t = lambda x: torch.from_numpy(x).to(dev)
torch.manual_seed(666)
sh = torch.rand(pos.shape[0], 1, 2).to(dev)
for mode in ("stash", "recompute"):
    a, r = t(albedo).requires_grad_(True), t(rough).requires_grad_(True)
    rgb = ops.shade(scene, t(nrm), a, r, t(pos), torch.zeros(3, device=dev), t(irr), sh, num_spec=256, backward=mode)
    rgb.backward(t(g_rgb))
    torch.cuda.synchronize()
    print(mode, float(rgb.sum()), float(r.grad.sum()))
