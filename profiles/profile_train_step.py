"""Times the pieces of one stage-2 Mat training step (MaterialModel.forward + backward) on cuda:0 and prints the torch profiler
kernel table:  python profiles/profile_train_step.py"""
import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texir_code_b200 import models, synth, ops
dev = torch.device('cuda:0')
sc = synth.make_room_scene(1 << 20, seed=666)
tex = torch.from_numpy(synth.make_radiance_texture(4096, seed=666))
from texir_code_b200 import Scene
scene = Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev)
n_, f_ = 1e-4, 100.0
proj = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, (f_ + n_) / (f_ - n_), -(2 * f_ * n_) / (f_ - n_)], [0, 0, 1.0, 0]])
faces = []
for fwd, up in (((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 1, 0), (0, 0, -1)), ((0, -1, 0), (0, 0, 1)), ((0, 0, 1), (0, 1, 0)), ((0, 0, -1), (0, 1, 0))):
    fz, uy = torch.tensor(fwd, dtype=torch.float32), torch.tensor(up, dtype=torch.float32)
    rx = torch.linalg.cross(uy, fz); w2c = torch.eye(4); w2c[:3, :3] = torch.stack([rx, uy, fz]); faces.append((proj @ w2c).t())
mvp = torch.stack(faces).to(dev)
cam = torch.zeros(3, device=dev)
mm = models.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[32, 256], scene=scene, materials_a_res=2048, materials_r_res=4096, cube_res=128, irrt=torch.rand(1024, 1024, 3) * 3.0, max_mip_level=13)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("rays_from_mvp %.3f ms" % timed(lambda: models.rays_from_mvp(mvp, cam, 128)))
rays, dd = models.rays_from_mvp(mvp, cam, 128)
print("gbuffer op    %.3f ms" % timed(lambda: ops.gbuffer(mm._scene(), rays, ray_differentials=dd)))
print("gbuffers()    %.3f ms" % timed(lambda: mm.gbuffers(mvp, cam)))
gb = mm.gbuffers(mvp, cam)
texc, texd = gb["texc"].reshape(-1, 2), gb["texd"].reshape(-1, 4)
print("fetch a mip   %.3f ms" % timed(lambda: ops.texture_fetch(mm.materials_a, texc, texd, max_mip_level=13)))
print("fetch r mip   %.3f ms" % timed(lambda: ops.texture_fetch(mm.materials_r, texc, texd, max_mip_level=13)))
print("fetch r lin   %.3f ms" % timed(lambda: ops.texture_fetch(mm.materials_r, texc)))
print("build mips r  %.3f ms" % timed(lambda: ops.build_mips(mm.materials_r.detach(), 13)))
def fb():
    mm.zero_grad(set_to_none=True)
    r = ops.texture_fetch(mm.materials_r, texc, texd, max_mip_level=13)
    r.sum().backward()
print("fetch r fwd+bwd %.3f ms" % timed(fb))
def step():
    mm.zero_grad(set_to_none=True)
    res = mm(mvp, 0, cam, 2)
    res["rgb"].sum().backward()
print("train step    %.3f ms" % timed(step))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
