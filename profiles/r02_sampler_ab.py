"""Does the clock sampler perturb the timed region?  The IrT step (device-resident, 3 steps) timed with (a) no sampler,
(b) the recipe's `nvidia-smi -lms 200` process, (c) NVML polled from a Python thread every 200 ms; twice each.
    python profiles/r02_sampler_ab.py"""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from texir_code_b200 import models, ops, synth
dev = torch.device("cuda:0")
sc = synth.make_room_scene(1 << 20, seed=666)
tex = torch.from_numpy(synth.make_radiance_texture(4096, seed=666))
tr = models.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, tex, sample_l=[2048, 64], device=dev)
P = 1 << 20
pos, nrm = synth.sample_surface_points(sc, P, seed=3)
d_pos, d_nrm = torch.from_numpy(pos).to(dev), torch.from_numpy(nrm).to(dev)
d_shift = torch.rand(P, 1, 2).to(dev)
timer = bench.Timer(dev, 1)
step = lambda: tr.irradiance(d_pos, d_nrm, d_shift)
for _ in range(2):
    step()


class NvmlThread:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.n, self.h, self.stop_, self.samples = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0), False, []
    def run(self):
        n = self.n
        while not self.stop_:
            self.samples.append((n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM), n.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                 n.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            time.sleep(0.2)
    def start(self):
        self.t = threading.Thread(target=self.run, daemon=True); self.t.start()
    def stop(self):
        self.stop_ = True; self.t.join(); return self.samples


for rep in range(2):
    ms, _ = timer.run(step, 3, 1)
    print("no sampler          %.2f ms/step" % (ms / 3), flush=True)
    s = bench.ClockSampler(0); s.start()
    ms, _ = timer.run(step, 3, 1)
    print("nvidia-smi -lms 200 %.2f ms/step  %s" % (ms / 3, s.stop()), flush=True)
    t = NvmlThread(); t.start()
    ms, _ = timer.run(step, 3, 1)
    sm = t.stop()
    print("NVML thread         %.2f ms/step  %d samples, sm %s" % (ms / 3, len(sm), sorted(set(x[0] for x in sm))), flush=True)
