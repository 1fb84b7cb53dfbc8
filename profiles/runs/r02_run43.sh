set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
for k in a b; do
( time timeout 1200 python bench.py > gpurun_out/r02_bench_default_$k.json 2> gpurun_out/r02_bench_default_$k.err ) 2> gpurun_out/r02_bench_default_$k.time; tail -3 gpurun_out/r02_bench_default_$k.time | head -1
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_default_$k.json').read().strip().splitlines()[-1])
m=d['mat_stage']
print('value',d['value'],d.get('step_ms'),'e2e',d['e2e']['value'],'cold',round(d['e2e_cold']['total_s'],3),'create',d['config']['scene_create_s'],'step16',round(m['train_step_stage2_16spp']['ms'],4),round(m['train_step_stage1_16spp']['ms'],4),'graphed',round(m['train_step_stage2_16spp']['ms_graphed'],4))
PY
done
