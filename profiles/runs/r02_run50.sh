set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --tris 4194304 --tex 8192 --no-mat > gpurun_out/r02_bench_n8_configs3.json 2> gpurun_out/r02_bench_n8_configs3.err; tail -2 gpurun_out/r02_bench_n8_configs3.err | cut -c1-200
timeout 900 $TR --master-port 29562 bench.py --gpus 8 --steps 5 --warmup 3 --tris 8388608 --tex 8192 > gpurun_out/r02_bench_n8_configs4.json 2> gpurun_out/r02_bench_n8_configs4.err; tail -2 gpurun_out/r02_bench_n8_configs4.err | cut -c1-200
python - <<'PY'
import json
for k in (3,4):
    d=json.loads(open('gpurun_out/r02_bench_n8_configs%d.json'%k).read().strip().splitlines()[-1])
    print(k, d['value'], d['ms_per_step'], d.get('step_ms'), d['e2e']['value'], d.get('relight_sharded'))
PY
