"""A/B of the G-buffer kernel's walks (lock step / shared first descent from the camera / free-running), of 8x4 pixel
tiles per warp and of the pyramid prefetch, on the training step's own shape (6 x 128^2 primary rays, T = 2^20):
    python profiles/r02_gbuffer_ab.py            runs every variant in its own process (the switches are read once)
Each variant prints the G-buffer kernel time (CUDA events, median of 200), a checksum of the G-buffer and the 16-spp
stage-2 step time (median of 5 x 50 steps)."""
import hashlib, os, statistics, subprocess, sys, time

VARIANTS = [("default (free-running walk, 8x4 tiles, prefetch)", {}),
            ("lock-step walk, rows (round-2 first kernel)", {"TEXIR_GB_WALK": "0", "TEXIR_GB_TILE": "0"}),
            ("lock step + shared first descent, rows", {"TEXIR_GB_WALK": "1", "TEXIR_GB_TILE": "0"}),
            ("lock step, 8x4 tiles", {"TEXIR_GB_WALK": "0"}),
            ("free-running, rows", {"TEXIR_GB_TILE": "0"}),
            ("default without the pyramid prefetch", {"TEXIR_MAT_PREFETCH": "0"}),
            ("default, shift drawn on the host", {"TEXIR_SHIFT_DEVICE": "0"})]


def child():
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from texir_code_b200 import models, ops, synth, Scene
    dev = torch.device("cuda:0")
    sc = synth.make_room_scene(1 << 20, seed=666)
    tex = torch.from_numpy(synth.make_radiance_texture(4096, seed=666))
    scene = Scene(sc.vertices, sc.triangles, sc.triangle_uvs, device=dev)
    mvp, cam = bench.cube_mvps(dev), torch.zeros(3, device=dev)
    for _ in range(5):
        gb = ops.gbuffer_mvp(scene, mvp, cam, 128, flip_v=True)
    ts = []
    for _ in range(200):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gb = ops.gbuffer_mvp(scene, mvp, cam, 128, flip_v=True)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    h = hashlib.sha1()
    for k in sorted(gb):
        if torch.is_tensor(gb[k]):
            h.update(gb[k].contiguous().cpu().numpy().tobytes())
    mm = models.MaterialModel(None, None, sc.triangle_uvs, tex, sample_l=[32, 16], scene=scene, materials_a_res=2048,
                              materials_r_res=4096, cube_res=128, irrt=torch.rand(1024, 1024, 3) * 3.0, max_mip_level=13)
    mm._warned_normals = True
    g = torch.randn(6, 128, 128, 3, device=dev)

    def step():
        mm.zero_grad(set_to_none=True)
        mm(mvp, 0, cam, 2)["rgb"].backward(g)
    for _ in range(10):
        step()
    reps = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            step()
        torch.cuda.synchronize()
        reps.append((time.perf_counter() - t0) / 50 * 1e3)
    torch.manual_seed(3)
    mm.zero_grad(set_to_none=True)
    mm(mvp, 0, cam, 2)["rgb"].backward(g)
    gsum = float(mm.materials_a.grad.double().abs().sum() + mm.materials_r.grad.double().abs().sum())
    print("gbuffer %.1f us (min %.1f) | sha1 %s | step 16 spp %.4f ms (min %.4f) | grad-abs-sum %.9e"
          % (statistics.median(ts), min(ts), h.hexdigest()[:16], statistics.median(reps), min(reps), gsum))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for name, env in VARIANTS:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, **env),
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout.strip().splitlines()
            print("%-45s %s" % (name, out[-1] if out else "(no output)"), flush=True)
