"""Runs bench.py (IrT leg only) once per build variant of the library under texir_code_b200/variants/ (TEXIR_LIB selects
the .so) and prints rate + checksum per variant:   python profiles/sweep_variants.py [--mat] name1 name2 ...
Variants are built by hand with extra -D flags ("measured on the way")."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mat = "--mat" in sys.argv
names = [a for a in sys.argv[1:] if not a.startswith("--")]
for name in names:
    if "=" in name:     # KEY=VALUE[,KEY=VALUE]: environment knobs of the default build (TEXIR_COHERENT_BITS=4 ...)
        env = dict(os.environ, **dict(kv.split("=", 1) for kv in name.split(",")))
    else:
        so = os.path.join(ROOT, "texir_code_b200", "variants", "libtexir_%s.so" % name)
        env = dict(os.environ, TEXIR_LIB=so)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--steps", "3", "--warmup", "2"] + ([] if mat else ["--no-mat"])
    out = subprocess.run(cmd, env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        extra = ""
        if mat:
            m = d["mat_stage"]
            extra = "  mat stash %.3f ms recompute %.3f ms relight %.1f ms step16 %.4f ms step256 %.3f ms" % (
                m["fwd_bwd_stash_256spp"]["ms"], m["fwd_bwd_recompute_256spp"]["ms"], m["relight_3072_1024_ms"],
                m["train_step_stage2_16spp"]["ms"], m["train_step_stage2_256spp"]["ms"])
        print("%-26s %.4e ray-samples/s  kernel %.2f ms  checksum %.6f  clocks %s%s" % (
            name, d["value"], d["roofline"]["kernel_ms"], d["e2e"]["checksum"], d["clocks"]["sm_mhz"], extra), flush=True)
    except Exception as exc:
        print(name, "FAILED", exc, out.stderr[-500:], flush=True)
