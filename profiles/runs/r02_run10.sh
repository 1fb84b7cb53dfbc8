set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_matstage.py -m gpu -q -x -p no:cacheprovider -k "two_ranks or sharded" > gpurun_out/r02_gpu_tests10.log 2>&1; tail -6 gpurun_out/r02_gpu_tests10.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for B in 512 4096 32768; do
  timeout 600 $TR --master-port 2953$((B % 7)) bench.py --gpus 4 --steps 5 --warmup 3 --no-mat --block $B > gpurun_out/r02_bench_n4_block$B.json 2> gpurun_out/r02_bench_n4_block$B.err
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_n4_block$B.json')); print('BLOCK $B', '%.4e' % d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])" || tail -5 gpurun_out/r02_bench_n4_block$B.err
done
timeout 900 $TR --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; tail -3 gpurun_out/r02_bench_n4.err | cut -c1-300; cat gpurun_out/r02_bench_n4.json
