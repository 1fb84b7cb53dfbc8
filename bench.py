#!/usr/bin/env python
"""bench.py -- ray-samples/s of the TexIR hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one process per GPU)
    python bench.py --impl reference [--steps K] [--warmup W]      the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): IrT generation on the synthetic closed room,
T = 2^20 triangles, 4096^2 HDR radiance texture, P = 1 048 576 texels, 2048 uniform samples per texel
-> 2 147 483 648 ray-samples per step.  A "step" is ONE pass over all P texels (the reference walks them in 2048
batches of 512, models/tracer_o3d_irt.py:156-173).

Multi-GPU (torchrun, one rank per GPU): the SAME job is split over the ranks (strong scaling): the texels are dealt out
block-cyclically along a Morton curve, every rank's kernel stores its results into every rank's buffer over NVLink peer
memory (the all-gather is the kernel's epilogue) and a flag barrier closes the step -- all inside the timed region.
`mat_stage_sharded` is one stage-2 optimisation step of the Mat stage with the pixels split over the ranks and the
single all-reduce of the texture gradients inside the timed region; `relight_sharded` the relighting forward.
--tris / --tex select the larger scenes (configs[3]: 4194304 / 8192; configs[4]: 8388608 / 8192).

One JSON line on stdout (rank 0); see the task contract for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray-samples/sec (fwd+bwd) at 1M tris / 2048 spp; HBM roofline %"
UNIT = "ray-samples/s"
NUM_TRIS = 1 << 20
TEX_RES = 4096
SPP = 2048
ISSUE_SLOTS_PER_S = 148 * 4 * 1.965e9   # warp-instruction issue slots of one B200 at its max SM clock


def b_alg(num_tris: int) -> int:
    """algorithmic bytes per ray-sample: single root-to-leaf BVH2 path of 64-byte nodes,
    one 4-triangle leaf (192 B), per-corner UVs (24 B), 4 bilinear taps of 12 B"""
    import math
    return 64 * math.ceil(math.log2(num_tris / 4)) + 192 + 24 + 48


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key: str):
    """dram bytes per ray-sample of a kernel from the committed ncu --set full captures (profiles/traffic.json)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        return float(d[key])
    except Exception:
        return None


def issue_frac(rate: float, key: str, clocks) -> dict:
    """instruction-issue yardstick next to the HBM one: warp instructions per ray-sample from the
    committed ncu capture x measured rate, against 148 SMs x 4 schedulers x the SM clock sampled during the run"""
    inst = ncu_traffic(key)
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    slots = 148 * 4 * mhz * 1e6
    return {"issue_frac": None if inst is None else rate * inst / slots, "warp_inst_per_ray_sample": inst,
            "issue_slots_per_s": slots}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = "/tmp/texir_clocks_%d_%d.csv" % (os.getpid(), index)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update({"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                        "power_w_max": max(power), "samples": len(sm)})
        return out


def make_workload():
    from texir_code_b200 import synth
    sc = synth.make_room_scene(NUM_TRIS, seed=666, tex_res=TEX_RES)
    tex = synth.make_radiance_texture(TEX_RES, seed=666)
    return sc, tex


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port; Open3D/Embree cannot be installed here)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(sc, tex, points: int, spp: int, steps: int, warmup: int, threads: int):
    """ray-samples/s of the reference IrT loop body on the host: reference Python restated in torch
    (single torch thread, as the reference sets) + the C BVH closest hit on `threads` cores"""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import texir_oracle as O
    from texir_code_b200 import synth
    torch.set_num_threads(1)  # trainer/generate_ir_texture.py:34
    oscene = O.OracleScene(sc.vertices, sc.triangles, threads=threads)
    pos, nrm = synth.sample_surface_points(sc, points, seed=21)
    tpos, tnrm, ttex = torch.from_numpy(pos), torch.from_numpy(nrm), torch.from_numpy(tex)
    torch.manual_seed(666)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.irt_forward(oscene, sc.triangle_uvs, ttex, tpos, tnrm, spp, "uniform", 512, None)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return points * spp * len(times) / total, total / len(times)


def workload_text(P: int) -> str:
    return ("IrT generation (BASELINE configs[1]): synthetic closed room, T=%d triangles, %d^2 fp32 radiance texture, "
            "ONE job of P=%d texels x 2048 spp uniform per step, forward only (the reference's IrT stage has no "
            "backward; the Mat-stage fwd+bwd is under 'mat_stage')" % (NUM_TRIS, TEX_RES, P))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sc, tex = make_workload()
    points = 2048  # four 512-texel batches of the reference loop (tracer_o3d_irt.py:156) = 4 194 304 ray-samples per step
    rate, sec = cpu_reference_rate(sc, tex, points, SPP, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.points),
                   "step": "four 512-texel batches of the reference loop (bounded sample of the %d-texel job)" % args.points},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "2048 texels x 2048 spp per step; reference Python restated in torch (1 torch thread, "
                                   "as the reference sets) + C BVH closest hit on %d threads; NOT Embree" % threads},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_line(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class Timer:
    """CUDA-event timing on the current stream with a barrier + synchronize on both sides, max over ranks"""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def run(self, fn, steps: int, warmup: int, per_step: bool = False):
        """-> (total ms of `steps` calls (max over ranks), [per-call ms on this rank])"""
        import torch
        import torch.distributed as dist
        for _ in range(warmup):
            fn()
        self.barrier()
        stream = torch.cuda.current_stream(self.dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if per_step else []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            if per_step:
                ev[i][0].record(stream)
            fn()
            if per_step:
                ev[i][1].record(stream)
        e1.record(stream)
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), [a.elapsed_time(b) for a, b in ev]


def cube_mvps(dev):
    """the six cube-face matrices of a camera at the origin, reference convention (datasets/dataset.py:402-469)"""
    import torch
    n_, f_ = 1e-4, 100.0
    proj = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, (f_ + n_) / (f_ - n_), -(2 * f_ * n_) / (f_ - n_)],
                         [0, 0, 1.0, 0]])
    faces = []
    for fwd, up in (((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 1, 0), (0, 0, -1)), ((0, -1, 0), (0, 0, 1)),
                    ((0, 0, 1), (0, 1, 0)), ((0, 0, -1), (0, 1, 0))):
        fz, uy = torch.tensor(fwd, dtype=torch.float32), torch.tensor(up, dtype=torch.float32)
        rx = torch.linalg.cross(uy, fz)
        w2c = torch.eye(4)
        w2c[:3, :3] = torch.stack([rx, uy, fz])          # camera at the origin looking along fwd
        faces.append((proj @ w2c).t())
    return torch.stack(faces).to(dev)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from texir_code_b200 import launch_count, models, ops, parallel, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    timer = Timer(dev, world)

    P = args.points
    sc, tex = make_workload()
    torch.zeros(1, device=dev)          # CUDA context + allocator are up before anything is timed
    from texir_code_b200 import _lib
    _lib.load().texir_version()
    # one-time costs of a fresh process (and, on a fresh box, of a cold page cache) are not what scene_create_s reports:
    # a small scene goes through the device builder first, so that its kernels are loaded before the clock starts
    try:
        from texir_code_b200 import Scene
        warm = synth.make_room_scene(8192, seed=1, tex_res=64, clutter=4)
        Scene(warm.vertices, warm.triangles, warm.triangle_uvs, device=dev).close()
    except Exception:
        pass
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter()
    tracer = models.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, torch.from_numpy(tex), sample_l=[SPP, 64],
                              device=dev, build_threads=max(1, (os.cpu_count() or 1) // world))
    tracer.scene.set_texture(tracer.texture)
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t_build
    if args.bvh != 2:
        tracer.scene.set_bvh(args.bvh)
    info = tracer.scene.info()
    # ONE texel set for the whole job, identical on every rank (seeded): rank r computes its share of it
    pos, nrm = synth.sample_surface_points(sc, P, seed=666)
    torch.manual_seed(666)
    shift = torch.rand(P, 1, 2)
    h_pos, h_nrm, h_shift = (torch.from_numpy(pos).pin_memory(), torch.from_numpy(nrm).pin_memory(), shift.pin_memory())
    h_out = torch.empty(P, 3).pin_memory()
    d_pos, d_nrm, d_shift = h_pos.to(dev), h_nrm.to(dev), h_shift.to(dev)
    collective = "none"
    if world > 1:
        collective = args.collective
        try:
            job = parallel.ShardedIrradiance(tracer, P, collective=collective, block=args.block)
        except Exception as exc:  # no peer mapping between these GPUs: the library collective does the same step
            if collective != "peer":
                raise
            sys.stderr.write("peer memory unavailable (%s); using the NCCL formulation\n" % exc)
            collective = "nccl"
            job = parallel.ShardedIrradiance(tracer, P, collective="nccl", block=args.block)
    else:
        job = parallel.ShardedIrradiance(tracer, P)

    def step_resident():
        job(d_pos, d_nrm, d_shift)

    def step_e2e():
        # the call a user makes: host buffers in, host buffer out
        p = h_pos.to(dev, non_blocking=True)
        n = h_nrm.to(dev, non_blocking=True)
        s = h_shift.to(dev, non_blocking=True)
        e = job(p, n, s)
        h_out.copy_(e, non_blocking=True)

    # ---- device-resident timing (value) ------------------------------------------------------------
    # the clock sampler (an nvidia-smi process) is started BEFORE the warm-up steps: spawning it takes tens of
    # milliseconds on a busy host, and a GPU left idle that long re-enters the timed region at ramping clocks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    timer.barrier()
    n_launch0 = launch_count()
    elapsed_ms, step_ms = timer.run(step_resident, args.steps, 0, per_step=True)
    launches = launch_count() - n_launch0
    clocks = sampler.stop() if rank == 0 else {}
    value = P * SPP * args.steps / (elapsed_ms * 1e-3)

    # the dominant kernel alone (this rank's share), CUDA events around its launch on the launching stream
    job.kernel_events = []
    timer.run(step_resident, args.steps, 0)
    torch.cuda.synchronize(dev)
    kernel_ms = [a.elapsed_time(b) for a, b in job.kernel_events]
    job.kernel_events = None

    # ---- end-to-end timing through the public API with host buffers --------------------------------
    e2e_ms, _ = timer.run(step_e2e, args.steps, max(1, min(args.warmup, 2)))
    e2e_value = P * SPP * args.steps / (e2e_ms * 1e-3)
    h2d = (h_pos.numel() + h_nrm.numel() + h_shift.numel()) * 4
    d2h = h_out.numel() * 4
    checksum = float(h_out.double().sum().item())

    # ---- the library formulation of the same collective, for comparison (N > 1, peer path only) -----
    nccl_cmp = None
    if world > 1 and collective == "peer":
        job2 = parallel.ShardedIrradiance(tracer, P, collective="nccl", block=args.block)
        ms2, _ = timer.run(lambda: job2(d_pos, d_nrm, d_shift), args.steps, 1)
        same = bool(torch.equal(job2.result, job.result))
        nccl_cmp = {"ms_per_step": ms2 / args.steps, "value": P * SPP * args.steps / (ms2 * 1e-3), "bit_identical_to_peer": same,
                    "what": "zeroed [P,3] buffer + kernel + ONE NCCL all-reduce(sum) per step"}

    extras = {}
    if world > 1 and not args.no_mat:
        extras["mat_stage_sharded"] = bench_mat_sharded(tracer.scene, sc, dev, args, tracer.texture, timer, rank, world)
        extras["relight_sharded"] = bench_relight_sharded(tracer.scene, sc, dev, args, timer, rank, world)

    # ---- the same job on the RGBE-packed texture (4 B / texel; exact for a texture that came out of a .hdr file) ----
    if world == 1 and not args.no_cold:
        from texir_code_b200 import postprocess
        t_cpu = tracer.texture
        bgr = t_cpu[0].permute(1, 2, 0).flip(0).flip(-1).contiguous().to(dev) / 32.0      # back to "file" order and scale
        rgbe = postprocess.rgbe_encode(bgr, bgr=True)                                     # quantise like cv2.imwrite would
        tex_q = (postprocess.rgbe_decode(rgbe, bgr=True).flip(-1).flip(0).permute(2, 0, 1).unsqueeze(0) * 32.0).contiguous()
        tracer.scene.set_texture(tex_q)
        ref_q = ops.irradiance(tracer.scene, d_pos[:65536], d_nrm[:65536], d_shift[:65536], SPP).clone()
        tracer.scene.set_texture_rgbe(rgbe, 5, flip=True)
        same = bool(torch.equal(ops.irradiance(tracer.scene, d_pos[:65536], d_nrm[:65536], d_shift[:65536], SPP), ref_q))
        order_q = ops.point_order(tracer.scene, d_pos)
        out_q = torch.empty(P, 3, device=dev)
        ms_q, _ = timer.run(lambda: ops.irradiance(tracer.scene, d_pos, d_nrm, d_shift, SPP, out=out_q, order=order_q), args.steps, 1)
        extras["irt_rgbe_texture"] = {"value": P * SPP * args.steps / (ms_q * 1e-3), "ms_per_step": ms_q / args.steps,
                                      "bit_identical_to_float_texture": same, "texture_bytes": int(rgbe.numel()),
                                      "what": "the radiance texture as Radiance RGBE texels (what hdr_texture.hdr holds), decoded per tap"}
        tracer.scene.set_texture(t_cpu)

    # ---- traversal work per ray (report nodes visited / triangles tested beside the roofline) ----
    trav = None
    if rank == 0:
        from texir_code_b200 import sample_util
        q = min(P, 512)
        l = sample_util.generate_dir(d_nrm[:q], 256, None, "uniform", shift=d_shift[:q])
        st = tracer.scene.cast_rays(torch.cat([d_pos[:q].unsqueeze(1).expand_as(l), l], -1), stats=True)
        node_b = 80 if info["bvh_kind"] == 8 else 64
        nv, nt = st["nodes_visited"].float().mean().item(), st["tris_tested"].float().mean().item()
        trav = {"nodes_per_ray": nv, "tris_per_ray": nt, "touched_bytes_per_ray": nv * node_b + nt * 52 + 48 + 64,
                "hit_rate": torch.isfinite(st["t_hit"]).float().mean().item()}

    # ---- single-GPU extras: Mat stage (fwd+bwd), training steps, cold start, CPU baseline ------------
    if world == 1 and not args.no_mat:
        extras["mat_stage"] = bench_material_stage(tracer.scene, sc, dev, args, tracer.texture, timer)
    if world == 1 and not args.no_cold:
        extras["e2e_cold"] = bench_cold(sc, tex, dev, h_pos, h_nrm, h_shift, h_out, P)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, sec = cpu_reference_rate(sc, tex, 4096, SPP, 3, 1, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "3 steps of 4096 texels (8 reference batches) x 2048 spp (%.1f s each); reference Python restated in torch (1 torch "
                         "thread) + C BVH closest hit on %d threads; NOT Embree" % (sec, threads)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        balg = b_alg(NUM_TRIS)
        k_ms = statistics.mean(kernel_ms)
        rays_per_launch = (P // world) * SPP
        achieved = balg * rays_per_launch / (k_ms * 1e-3) / 1e9
        tr = ncu_traffic("dram_bytes_per_ray_sample")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "step_ms": [round(x, 3) for x in step_ms], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_text(P),
                "points": P, "points_per_gpu": P // world, "spp": SPP, "triangles": NUM_TRIS, "texture": [TEX_RES, TEX_RES],
                "bvh": "BVH%d" % info["bvh_kind"], "scene_bytes": info["device_bytes"],
                "l2": "no flush: BVH + triangles + UVs + texture = %.0f MB resident, larger than the 126 MB L2"
                      % (info["device_bytes"] / 1e6),
                "scene_create_s": round(t_build, 3), "bvh_builder": {1: "host", 2: "device"}.get(info.get("builder"), "?"),
                "sharding": "none" if world == 1 else "texels dealt out in %d-blocks along a Morton curve, block-cyclic over %d ranks" % (args.block, world),
                "collective": {"none": "none (one GPU)", "peer": "all-gather fused into the kernel epilogue (NVLink peer stores) + flag barrier, inside the timed step",
                               "nccl": "one NCCL all-reduce(sum) of [P,3] per step, inside the timed step"}[collective],
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if tr is None else tr * rays_per_launch, "peak_source": peak_src,
                         "kernel": "k_irradiance_lean<4,FRONT> (QBVH2, shared first descent)" if info["bvh_kind"] == 2 else "k_irradiance<8>",
                         "kernel_ms": k_ms, "alg_bytes_per_ray_sample": balg, "ray_samples_per_launch": rays_per_launch,
                         **issue_frac(rays_per_launch / (k_ms * 1e-3), "warp_inst_per_ray_sample", clocks),
                         "issue_frac_note": "see DESIGN.md section 3: the kernel is bound by instruction issue, not by HBM"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "checksum": checksum},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if trav is not None:
            line["traversal"] = trav
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if nccl_cmp is not None:
            line["irt_nccl_formulation"] = nccl_cmp
        line.update(extras)
        print_line(json.dumps(line))
    job.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_cold(sc, tex, dev, h_pos, h_nrm, h_shift, h_out, P):
    """one IrT from nothing, as the reference's runner does it once (trainer/generate_ir_texture.py:64-82): scene
    creation (BVH build + upload), texture upload, H2D of the texels, the fused launch, D2H of the result"""
    import torch
    from texir_code_b200 import models
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    tr = models.TracerO3d(sc.vertices, sc.triangles, sc.triangle_uvs, torch.from_numpy(tex), sample_l=[SPP, 64], device=dev)
    tr.scene.set_texture(tr.texture)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    d_in = [h.to(dev, non_blocking=True) for h in (h_pos, h_nrm, h_shift)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    e = tr.irradiance(*d_in)
    ev[1].record()
    h_out.copy_(e, non_blocking=True)
    torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    inf = tr.scene.info()
    tr.scene.close()
    return {"total_s": t2 - t0, "scene_create_and_texture_s": t1 - t0, "irt_s": t2 - t1, "scene_create_s": inf["build_seconds"],
            "irradiance_device_s": ev[0].elapsed_time(ev[1]) * 1e-3,
            "bvh_builder": {1: "host", 2: "device"}.get(inf.get("builder"), "?"),
            "ray_samples_per_s": P * SPP / (t2 - t0), "what": "scene_create + set_texture + H2D + irradiance + D2H, wall clock"}


def mat_inputs(scene, sc, dev):
    import numpy as np
    import torch
    from texir_code_b200 import synth
    rays = synth.cube_map_rays(128)
    hit = scene.cast_rays(torch.from_numpy(rays).to(dev))
    pos, nrm = synth.gbuffer_from_hits(sc, rays, hit["t_hit"].cpu().numpy(),
                                       hit["primitive_ids"].cpu().numpy().view(np.uint32))
    P = pos.shape[0]
    rough, albedo, irr, g_rgb = synth.make_materials(P, seed=8)
    return pos, nrm, rough, albedo, irr, g_rgb


def bench_material_stage(scene, sc, dev, args, texture, timer):
    """BASELINE configs[2] shape: 6x128^2 cube-map pixels, importance samples, forward + backward -- first-class:
    >= 20 timed steps per leg, per-leg roofline on the algorithmic-bytes model (2 B_alg when the backward re-casts)
    and the issue-slot fraction; then whole optimisation steps through MaterialModel.forward"""
    import torch
    from texir_code_b200 import launch_count, models, ops
    pos, nrm, rough, albedo, irr, g_rgb = mat_inputs(scene, sc, dev)
    P = pos.shape[0]
    torch.manual_seed(666)
    sh = torch.rand(P, 1, 2).to(dev)
    t = lambda x: torch.from_numpy(x).to(dev)
    nrm_d, pos_d, irr_d, g_d, cam = t(nrm), t(pos), t(irr), t(g_rgb), torch.zeros(3, device=dev)
    peak, _ = measured_peak_gbs()
    balg = b_alg(NUM_TRIS)
    steps = max(20, args.steps)
    out = {"pixels": P, "timed_steps": steps}
    for mode, N in (("stash", 256), ("recompute", 256), ("stash", 16), ("recompute", 16)):  # 16 spp = configs/syn.conf training
        a = t(albedo).requires_grad_(True)
        r = t(rough).requires_grad_(True)

        def step():
            a.grad = None
            r.grad = None
            rgb = ops.shade(scene, nrm_d, a, r, pos_d, cam, irr_d, sh, num_spec=N, backward=mode)
            rgb.backward(g_d)

        sampler = ClockSampler(dev.index)
        sampler.start()
        ms, _ = timer.run(step, steps, 3)
        clocks = sampler.stop()
        ms /= steps
        rate = P * N / (ms * 1e-3)
        casts = 2 if mode == "recompute" else 1   # ray casts per ray-sample: forward, + the backward's re-trace
        leg = {"ms": ms, "ray_samples_per_s": rate, "spp": N,
               "roofline": {"bound": "hbm", "achieved": rate * casts * balg / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": rate * casts * balg / 1e9 / peak, "alg_bytes_per_ray_sample": casts * balg,
                            "traffic": ncu_traffic("shade_%s_dram_bytes_per_ray_sample" % mode),
                            **issue_frac(rate, "shade_%s_warp_inst_per_ray_sample" % mode, clocks)},
               "clocks": clocks}
        out["fwd_bwd_%s_%dspp" % (mode, N)] = leg

    # BASELINE configs[4] shape (relighting, forward only): 3072 uniform Monte-Carlo diffuse + 1024 importance
    # specular samples per pixel = 4096 spp; and the reference's own default 1024 + 256 (configs/test.conf:10)
    sh2 = torch.rand(P, 1, 2).to(dev)
    a0, r0 = t(albedo), t(rough)
    for tag, nd, ns in (("relight_3072_1024", 3072, 1024), ("relight_1024_256", 1024, 256)):
        ms, _ = timer.run(lambda: ops.shade(scene, nrm_d, a0, r0, pos_d, cam, None, sh, num_spec=ns, shift_diff=sh2,
                                            num_diff=nd, mode_diff="uniform", eps=1e-6), 5, 2)
        ms /= 5
        out[tag + "_ms"] = ms
        out[tag + "_ray_samples_per_s"] = P * (nd + ns) / (ms * 1e-3)

    # whole optimisation steps of the Mat stage through the reference-signature call: G-buffer from the cube-face mvps
    # -> fetches of materials_a (2048^2 x 3), materials_r (4096^2 x 1) and the IrT (1024^2 x 3) -> fused render ->
    # backward into the two textures, at the training sample count (configs/syn.conf:24: 16) and the final-eval one (256)
    mvp = cube_mvps(dev)
    g_img = g_d.reshape(6, 128, 128, 3)
    for stage in (1, 2):
        for nspec in (16, 256):
            mm = models.MaterialModel(None, None, sc.triangle_uvs, texture, sample_l=[32, nspec], scene=scene,
                                      materials_a_res=2048, materials_r_res=4096, cube_res=128,
                                      irrt=torch.rand(1024, 1024, 3) * 3.0, max_mip_level=13)
            mm._warned_normals = True
            if stage == 1:
                mm.materials_a.requires_grad_(False)   # trainer/train_material.py:477-479

            def train_step():
                mm.zero_grad(set_to_none=True)
                res = mm(mvp, 0, cam, stage)
                res["rgb"].backward(g_img)
                # (the optimiser would run here: the next step rebuilds the pyramids of the trainable textures)

            train_step()
            n0 = launch_count()
            train_step()
            per_step_launches = launch_count() - n0
            reps = 50
            torch.cuda.synchronize(dev)
            w0 = time.perf_counter()
            ms, _ = timer.run(train_step, reps, 3)
            wall = (time.perf_counter() - w0) / (reps + 3) * 1e3
            key = "train_step_stage%d_%dspp" % (stage, nspec)
            out[key] = {"ms": ms / reps, "host_ms_per_step": wall, "launches": per_step_launches, "timed_steps": reps,
                        "ray_samples_per_s": P * nspec / (ms / reps * 1e-3)}
            # the same step captured once as a CUDA graph and replayed (texir_code_b200.graphed.GraphedMatStep): one launch
            # call per step, the shift still drawn from torch's CPU generator stream
            try:
                from texir_code_b200.graphed import GraphedMatStep
                gs = GraphedMatStep(mm, stage, lambda o, gt: (o["rgb"] * gt).sum(), example=(mvp, cam, g_img))
                w0 = time.perf_counter()
                ms_g, _ = timer.run(gs.replay, reps, 3)
                out[key]["ms_graphed"] = ms_g / reps
                out[key]["host_ms_per_step_graphed"] = (time.perf_counter() - w0) / (reps + 3) * 1e3
                del gs
            except Exception as exc:
                out[key]["ms_graphed"] = None
                out[key]["graphed_unavailable"] = repr(exc)[:300]
            if stage == 2 and nspec == 16:
                out[key]["breakdown"] = train_step_breakdown(train_step, dev)
                # the same step with the view's G-buffer kept between steps (opt-in: MaterialModel(cache_gbuffers=True),
                # valid because geometry and cameras are constant during the reference's optimisation)
                mm.cache_gbuffers = True
                train_step()
                ms_c, _ = timer.run(train_step, reps, 3)
                out[key]["ms_with_cached_gbuffer"] = ms_c / reps
                # ... and with the Cranley-Patterson shift drawn on the device (opt-in shift_rng='device': not the
                # reference's CPU stream any more; torch.rand(98304,1,2) on the host is what bounds the step otherwise)
                mm.shift_rng = "device"
                train_step()
                ms_d, _ = timer.run(train_step, reps, 3)
                out[key]["ms_with_cached_gbuffer_and_device_rng"] = ms_d / reps
                mm.shift_rng = "reference"
                mm.cache_gbuffers = False
            del mm
    out["train_step_what"] = ("MaterialModel.forward(mvp, id, cam, stage) + backward: G-buffer of 6x128^2 primary rays (1 launch), "
                              "pyramids of the two material textures (2 launches each; the IrT's is cached), the four "
                              "fetches (1 launch), fused render (1), ONE backward kernel into texture space, 1 fold per "
                              "texture; 'launches' counts this library's kernels per step")
    return out


def train_step_breakdown(step, dev, reps=20):
    """device time per kernel of one training step, microseconds (torch profiler / CUPTI over `reps` steps); kernels on
    different streams overlap, so the parts add up to more than the step"""
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                step()
            torch.cuda.synchronize(dev)
        parts = {}
        for e in prof.key_averages():
            t = getattr(e, "self_device_time_total", None)
            if t is None:
                t = getattr(e, "self_cuda_time_total", 0.0)
            if t <= 0:
                continue
            name = e.key
            for cut in ("void ", "(anonymous namespace)::", "texir::"):
                name = name.replace(cut, "")
            name = name.split("(")[0].strip()
            parts[name] = parts.get(name, 0.0) + t / reps
        top = dict(sorted(parts.items(), key=lambda kv: -kv[1])[:14])
        return {"us_per_step": {k: round(v, 1) for k, v in top.items()}, "sum_us": round(sum(parts.values()), 1),
                "note": "device time per kernel from the torch profiler (CUPTI), %d steps; the mt19937 shift kernel and "
                        "the pyramid builds run on side streams beside the G-buffer walk" % reps}
    except Exception as exc:   # a box without CUPTI access still gets the step time above
        return {"unavailable": repr(exc)[:200]}


def bench_mat_sharded(scene, sc, dev, args, texture, timer, rank, world):
    """ONE stage-2 optimisation step of the Mat stage over all ranks: pixels dealt out in 128-blocks, every rank renders
    and back-propagates its share, then the single all-reduce over the flat texture-gradient buffer -- all timed"""
    import torch
    import torch.distributed as dist
    from texir_code_b200 import models, parallel
    torch.manual_seed(4)
    g_full = torch.randn(6, 128, 128, 3)
    g_loc = parallel.local_pixels(g_full, rank, world, 6).to(dev)
    mvp, cam = cube_mvps(dev), torch.zeros(3, device=dev)
    out = {"pixels": 6 * 128 * 128, "pixels_per_gpu": 6 * 128 * 128 // world}
    for nspec in (16, 256):
        mm = models.MaterialModel(None, None, sc.triangle_uvs, texture, sample_l=[32, nspec], scene=scene,
                                  materials_a_res=2048, materials_r_res=4096, cube_res=128,
                                  irrt=torch.rand(1024, 1024, 3) * 3.0, max_mip_level=13, shard=(rank, world))
        mm._warned_normals = True
        params = [mm.materials_a, mm.materials_r]
        parts = {"compute": 0.0, "allreduce": 0.0}
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

        def step(record=False):
            mm.zero_grad(set_to_none=True)
            if record:
                ev[0].record()
            res = mm(mvp, 0, cam, 2)
            res["rgb"].backward(g_loc)
            if record:
                ev[1].record()
            nbytes = parallel.allreduce_gradients(params)
            if record:
                ev[2].record()
            return nbytes

        nbytes = step()
        reps = 30
        ms, _ = timer.run(step, reps, 3)
        for _ in range(5):
            step(True)
            torch.cuda.synchronize(dev)
            parts["compute"] += ev[0].elapsed_time(ev[1]) / 5
            parts["allreduce"] += ev[1].elapsed_time(ev[2]) / 5
        t = torch.tensor([parts["compute"], parts["allreduce"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["stage2_%dspp" % nspec] = {"ms_per_step": ms / reps, "ray_samples_per_s": 6 * 128 * 128 * nspec / (ms / reps * 1e-3),
                                       "allreduce_bytes": nbytes, "compute_ms_max_rank": float(t[0]), "allreduce_ms_max_rank": float(t[1]),
                                       "timed_steps": reps}
        # the other synchronisation: all-gather of the per-pixel gradient rows (11 floats per pixel) inside the backward,
        # every rank scatters all rows -- no all-reduce of the 112 MiB textures
        mm.grad_sync = "gather"

        def step_gather():
            mm.zero_grad(set_to_none=True)
            res = mm(mvp, 0, cam, 2)
            res["rgb"].backward(g_loc)

        step_gather()
        ms, _ = timer.run(step_gather, reps, 3)
        out["stage2_%dspp" % nspec]["ms_per_step_gather_sync"] = ms / reps
        out["stage2_%dspp" % nspec]["gather_bytes"] = 6 * 128 * 128 * 11 * 4
        # both synchronisations with the rank's whole step (and the collective) captured as ONE CUDA graph
        try:
            from texir_code_b200.graphed import GraphedMatStep
            lf = lambda o, gt: (o["rgb"] * gt).sum()
            gs = GraphedMatStep(mm, 2, lf, example=(mvp, cam, g_loc))
            ms, _ = timer.run(gs.replay, reps, 3)
            out["stage2_%dspp" % nspec]["ms_per_step_graphed_gather_sync"] = ms / reps
            del gs
            mm.grad_sync = "allreduce"
            gs = GraphedMatStep(mm, 2, lf, example=(mvp, cam, g_loc), sync=lambda: parallel.allreduce_gradients(params))
            ms, _ = timer.run(gs.replay, reps, 3)
            out["stage2_%dspp" % nspec]["ms_per_step_graphed"] = ms / reps
            del gs
        except Exception as exc:
            out["stage2_%dspp" % nspec]["graphed_unavailable"] = repr(exc)[:300]
        mm.grad_sync = "gather"
        mm.shift_rng, mm.cache_gbuffers = "device", True      # opt-ins: shift from the CUDA generator, G-buffer kept per view
        torch.cuda.manual_seed(1234)
        step_gather()
        ms, _ = timer.run(step_gather, reps, 3)
        out["stage2_%dspp" % nspec]["ms_per_step_gather_sync_device_rng_cached_gbuffer"] = ms / reps
        mm.grad_sync = "allreduce"
        step()
        ms, _ = timer.run(step, reps, 3)
        out["stage2_%dspp" % nspec]["ms_per_step_device_rng_cached_gbuffer"] = ms / reps
        del mm
    out["what"] = ("MaterialModel(shard=(rank, world)).forward(mvp, id, cam, 2) + backward on this rank's pixels, then ONE "
                   "all-reduce(sum) over the flat buffer that holds both texture gradients (2048^2x3 + 4096^2x1 fp32), "
                   "inside the timed step; the limiter at small sample counts is that dense all-reduce")
    return out


def bench_relight_sharded(scene, sc, dev, args, timer, rank, world):
    """BASELINE configs[4] shape over all ranks: relighting forward (3072 uniform diffuse + 1024 importance specular
    samples per pixel, models/test_nvdiffrast.py:275-295), pixels dealt out in 128-blocks (block-cyclic: every rank gets
    the same mix of cube faces), one all_gather of the colours + the permutation back inside the timed step"""
    import torch
    import torch.distributed as dist
    from texir_code_b200 import ops, parallel
    pos, nrm, rough, albedo, irr, g_rgb = mat_inputs(scene, sc, dev)
    P = pos.shape[0]
    sel = parallel.block_cyclic_index(P, rank, world, 128)
    order = torch.cat([parallel.block_cyclic_index(P, r, world, 128) for r in range(world)]).to(dev)   # rank-major -> pixel
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(P, device=dev)
    t = lambda x: torch.from_numpy(x)[sel].to(dev)
    torch.manual_seed(666)
    sh = torch.rand(P, 1, 2)[sel].to(dev)
    sh2 = torch.rand(P, 1, 2)[sel].to(dev)
    nrm_d, pos_d, a0, r0, cam = t(nrm), t(pos), t(albedo), t(rough), torch.zeros(3, device=dev)
    full = torch.empty(P, 3, device=dev)
    out = {"pixels": P}
    for tag, nd, ns in (("3072_1024", 3072, 1024), ("1024_256", 1024, 256)):
        def step():
            rgb = ops.shade(scene, nrm_d, a0, r0, pos_d, cam, None, sh, num_spec=ns, shift_diff=sh2, num_diff=nd,
                            mode_diff="uniform", eps=1e-6)
            dist.all_gather_into_tensor(full, rgb)
            return full[inverse]
        ms, _ = timer.run(step, 10, 2)
        out[tag] = {"ms_per_step": ms / 10, "ray_samples_per_s": P * (nd + ns) / (ms / 10 * 1e-3)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1 << 20, help="texels of the IrT job (default: the full 1024^2 IrT)")
    ap.add_argument("--bvh", type=int, default=2, choices=[2, 8], help="BVH layout the kernels traverse")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU IrT: results stored to every rank by the kernel (peer) or one NCCL all-reduce")
    ap.add_argument("--block", type=int, default=512, help="multi-GPU IrT: texels per dealt-out block of the Morton order")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-mat", action="store_true", help="skip the material-stage legs")
    ap.add_argument("--no-cold", action="store_true", help="skip the cold-start (scene creation included) leg")
    ap.add_argument("--tris", type=int, default=NUM_TRIS, help="triangles of the synthetic scene (configs[3]: 4194304)")
    ap.add_argument("--tex", type=int, default=TEX_RES, help="radiance texture resolution (configs[3]: 8192)")
    args = ap.parse_args()
    globals().update(NUM_TRIS=args.tris, TEX_RES=args.tex)
    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 meanwhile (NCCL announces its version
    # there during init) goes to stderr instead
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    globals()["print_line"] = lambda text: (real_stdout.write(text + "\n"), real_stdout.flush())
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
