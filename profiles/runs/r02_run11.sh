set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 600 python -m pytest tests/test_gpu_rgbe_conf.py -m gpu -q -x -p no:cacheprovider -k "device_shift" > gpurun_out/r02_gpu_tests11.log 2>&1; tail -8 gpurun_out/r02_gpu_tests11.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace11.log 2>&1; head -26 gpurun_out/r02_train_trace11.log | cut -c1-75,150-215
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n1e.json 2> gpurun_out/r02_bench_n1e.err; tail -5 gpurun_out/r02_bench_n1e.err; cat gpurun_out/r02_bench_n1e.json
timeout 600 ncu --set full --clock-control none -k regex:k_gbuffer_mvp -c 1 -o gpurun_out/r02_ncu_gbuffer -f python profiles/r02_train_step_trace.py > gpurun_out/r02_ncu_gbuffer.log 2>&1
ncu -i gpurun_out/r02_ncu_gbuffer.ncu-rep --page raw --csv > gpurun_out/r02_ncu_gbuffer_raw.csv 2>/dev/null; rm -f gpurun_out/r02_ncu_gbuffer.ncu-rep
python - <<'PY'
import csv, io
raw = open("gpurun_out/r02_ncu_gbuffer_raw.csv").read()
rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
h, v = rows[0], rows[2]
for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"):
    if k in h: print(k, v[h.index(k)], rows[1][h.index(k)])
for i, k in enumerate(h):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and float(v[i] or 0) > 0.5:
        print(" stall", k[34:-23], v[i])
PY
