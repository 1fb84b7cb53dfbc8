set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; tail -5 gpurun_out/r02_bench_n2b.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
print(json.dumps(d['mat_stage_sharded'], indent=0)[:3000])
PY
