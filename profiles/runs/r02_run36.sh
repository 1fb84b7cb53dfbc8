set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
( time timeout 1200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2> gpurun_out/r02_bench_default.time; tail -3 gpurun_out/r02_bench_default.time; tail -3 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'kernel',d['roofline']['kernel_ms'],'e2e',d['e2e']['value'],'cold',d['e2e_cold']['total_s'],d['e2e_cold']['irradiance_device_s'])
PY
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cat gpurun_out/r02_bench_reference.json | cut -c1-300
