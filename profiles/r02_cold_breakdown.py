"""Where the time of the COLD IrT job goes (bench.py e2e_cold): one IrT from nothing, three times in one process, every
phase bracketed by a device synchronise; plus the warm kernel for comparison.   python profiles/r02_cold_breakdown.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texir_code_b200 import models, ops, synth
dev = torch.device("cuda:0")
T, P, SPP = 1 << 20, 1 << 20, 2048
sc = synth.make_room_scene(T, seed=666)
tex = synth.make_radiance_texture(4096, seed=666)
pos, nrm = synth.sample_surface_points(sc, P, seed=3)
h_pos, h_nrm = torch.from_numpy(pos).pin_memory(), torch.from_numpy(nrm).pin_memory()
torch.manual_seed(1)
h_shift = torch.rand(P, 1, 2).pin_memory()
h_out = torch.empty(P, 3).pin_memory()
torch.zeros(1, device=dev); torch.cuda.synchronize()
from texir_code_b200 import _lib; _lib.load()


def sync():
    torch.cuda.synchronize(dev)
    return time.perf_counter()


for rep in range(3):
    t0 = sync()
    tr = models.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, torch.from_numpy(tex), sample_l=[SPP, 64], device=dev)
    t1 = sync()
    tr.scene.set_texture(tr.texture)
    t2 = sync()
    dp, dn, ds = h_pos.to(dev, non_blocking=True), h_nrm.to(dev, non_blocking=True), h_shift.to(dev, non_blocking=True)
    t3 = sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e = tr.irradiance(dp, dn, ds)
    b.record()
    t4 = sync()
    h_out.copy_(e, non_blocking=True)
    t5 = sync()
    e2 = tr.irradiance(dp, dn, ds)
    t6 = sync()
    print("rep %d: scene %.3f | texture %.3f | h2d %.4f | first irradiance %.3f (device %.3f) | d2h %.4f | total %.3f || second irradiance %.3f s"
          % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, a.elapsed_time(b) * 1e-3, t5 - t4, t5 - t0, t6 - t5), flush=True)
    tr.scene.close()
    del tr, e, e2
