set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
for a in 3 2 4 5 3; do
  TEXIR_COHERENT_BITS=$a timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-mat --no-cold 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('abits=$a value %.4e ms %.2f checksum %s' % (d['value'], d['ms_per_step'], d['e2e'].get('checksum')))"
done > gpurun_out/r02_abits_sweep.txt 2>&1
cat gpurun_out/r02_abits_sweep.txt
