set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_matstage.py tests/test_gpu_rgbe_conf.py tests/test_gpu_texture.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests21.log 2>&1; tail -30 gpurun_out/r02_gpu_tests20.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-cold > gpurun_out/r02_bench_n1g.json 2> gpurun_out/r02_bench_n1g.err; tail -5 gpurun_out/r02_bench_n1g.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1g.json').read().strip().splitlines()[-1])
for k,v in d['mat_stage'].items():
    if k.startswith('train_step') and isinstance(v,dict): print(k,{kk:vv for kk,vv in v.items() if kk!='breakdown'})
print(d['mat_stage']['train_step_stage2_16spp'].get('breakdown'))
PY
