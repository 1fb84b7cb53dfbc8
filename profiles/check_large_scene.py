"""One-off parity check at the larger BASELINE scene sizes (configs[3]/[4]): the fused IrT kernel against the
oracle on 512 points, and BVH2(QBVH2) vs BVH8 hit equality.   python profiles/check_large_scene.py <log2 tris> <tex>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import texir_oracle as O
import texir_code_b200 as tx
from texir_code_b200 import synth

lt, res = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
t0 = time.time(); sc = synth.make_room_scene(1 << lt, tex_res=res); tex = torch.from_numpy(synth.make_radiance_texture(res))
print("synth %.1fs" % (time.time() - t0))
t0 = time.time(); scene = tx.Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev); scene.set_texture(tex)
print("scene build+upload %.1fs" % (time.time() - t0), scene.info())
P, N = 512, 256
pos, nrm = synth.sample_surface_points(sc, P, seed=1)
torch.manual_seed(666); sh = torch.rand(P, 1, 2)
E = tx.irradiance(scene, torch.from_numpy(pos).to(dev), torch.from_numpy(nrm).to(dev), sh.to(dev), N).cpu()
t0 = time.time(); osc = O.OracleScene(sc.vertices, sc.triangles); print("oracle bvh %.1fs" % (time.time() - t0))
Eo = O.irt_forward(osc, sc.triangle_uvs, tex, torch.from_numpy(pos), torch.from_numpy(nrm), N, "uniform", 512, sh)
print("IrT rel-L2 vs oracle: %.3e" % float(torch.linalg.norm(E - Eo) / torch.linalg.norm(Eo)))
l = tx.generate_dir(torch.from_numpy(nrm).to(dev), N, None, "uniform", shift=sh.to(dev))
rays = torch.cat([torch.from_numpy(pos).to(dev).unsqueeze(1).expand_as(l), l], -1).reshape(-1, 6)
a = scene.cast_rays(rays); scene.set_bvh(8); b = scene.cast_rays(rays)
t, i, u = osc.cast_rays_np(rays[::7].cpu().numpy())
print("QBVH2 == BVH8:", all(torch.equal(a[k], b[k]) for k in a), "| == oracle:",
      np.array_equal(a["primitive_ids"][::7].cpu().numpy().view(np.uint32), i) and np.array_equal(a["t_hit"][::7].cpu().numpy(), t))

# where does the residual come from?  per-sample comparison of directions and radiance
scene.set_bvh(2)
rad, dirs = tx.radiance_samples(scene, torch.from_numpy(pos).to(dev), torch.from_numpy(nrm).to(dev), sh.to(dev), N, return_dirs=True)
lo = O.generate_dir(torch.from_numpy(nrm), N, None, "uniform", shift=sh)
ro = O.query_irf(osc, sc.triangle_uvs, tex, torch.from_numpy(pos).unsqueeze(1).expand_as(lo), lo.unsqueeze(-2), N)
dd = (dirs.cpu() != lo).any(-1)
rr = (rad.cpu() != ro).any(-1)
print("samples %d | direction differs (sin/cos ulp) %d | radiance differs %d | radiance differs with equal direction %d"
      % (dd.numel(), int(dd.sum()), int(rr.sum()), int((rr & ~dd).sum())))
# radiance of the oracle for OUR directions: isolates everything after direction generation
ro2 = O.query_irf(osc, sc.triangle_uvs, tex, torch.from_numpy(pos).unsqueeze(1).expand_as(lo), dirs.cpu().unsqueeze(-2), N)
print("radiance for identical directions: max abs diff %.3e, mismatching samples %d" % (float((rad.cpu() - ro2).abs().max()), int((rad.cpu() != ro2).any(-1).sum())))
