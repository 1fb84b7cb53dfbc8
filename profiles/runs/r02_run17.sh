set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_matstage.py -m gpu -q -x -p no:cacheprovider -k "two_ranks or sharded" > gpurun_out/r02_gpu_tests17.log 2>&1; tail -3 gpurun_out/r02_gpu_tests17.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -3 gpurun_out/r02_bench_n2.err | cut -c1-300; cat gpurun_out/r02_bench_n2.json | cut -c1-5000
timeout 600 $TR --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_n2_reference.json 2> gpurun_out/r02_bench_n2_reference.err; cat gpurun_out/r02_bench_n2_reference.json | cut -c1-600
