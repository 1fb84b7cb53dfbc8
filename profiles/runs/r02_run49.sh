set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
( time timeout 900 python bench.py --tris 4194304 --tex 8192 --steps 3 --warmup 2 --no-cpu --no-mat > gpurun_out/r02_bench_n1_configs3.json 2> gpurun_out/r02_bench_n1_configs3.err ) 2>&1 | grep real
( time timeout 900 python bench.py --tris 8388608 --tex 8192 --steps 3 --warmup 2 --no-cpu --no-mat > gpurun_out/r02_bench_n1_configs4.json 2> gpurun_out/r02_bench_n1_configs4.err ) 2>&1 | grep real
python - <<'PY'
import json
for k in (3,4):
    d=json.loads(open('gpurun_out/r02_bench_n1_configs%d.json'%k).read().strip().splitlines()[-1])
    print(k, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['alg_bytes_per_ray_sample'], d['config']['scene_create_s'], d.get('e2e_cold',{}).get('total_s'), d['config']['triangles'])
PY
