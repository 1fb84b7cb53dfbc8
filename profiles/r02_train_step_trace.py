"""torch-profiler kernel table of the Mat-stage training step (stage 2) at the training sample count (16 spp,
configs/syn.conf:24) and at 256:   python profiles/r02_train_step_trace.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from texir_code_b200 import models, synth, Scene
dev = torch.device('cuda:0')
sc = synth.make_room_scene(1 << 20, seed=666)
tex = torch.from_numpy(synth.make_radiance_texture(4096, seed=666))
scene = Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev)
mvp, cam = bench.cube_mvps(dev), torch.zeros(3, device=dev)
g = torch.randn(6, 128, 128, 3, device=dev)
from torch.profiler import profile, ProfilerActivity
for nspec in (16, 256):
    mm = models.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[32, nspec], scene=scene, materials_a_res=2048,
                              materials_r_res=4096, cube_res=128, irrt=torch.rand(1024, 1024, 3) * 3.0, max_mip_level=13)
    mm._warned_normals = True
    def step():
        mm.zero_grad(set_to_none=True)
        res = mm(mvp, 0, cam, 2)
        res["rgb"].backward(g)
    for _ in range(5): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): step()
    torch.cuda.synchronize()
    print("== %d spp: %.3f ms per step (wall, 50 steps)" % (nspec, (time.perf_counter() - t0) / 50 * 1e3))
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(10): step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
