set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
for k in 1 2; do
timeout 600 python bench.py --no-cpu --no-mat --no-cold 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default steps/warmup: value %.4e ms %.2f step_ms %s kernel %.2f e2e_ms %.2f' % (d['value'], d['ms_per_step'], d.get('step_ms'), d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
timeout 600 python bench.py --no-cpu --no-mat --no-cold --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('5 steps: value %.4e ms %.2f step_ms %s kernel %.2f e2e_ms %.2f' % (d['value'], d['ms_per_step'], d.get('step_ms'), d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
