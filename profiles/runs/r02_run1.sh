set -x
cd $GRAFT_REPO_ROOT   # (these scripts ran as profiles/r02_runNN.sh; paths inside are repository-relative)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -2 gpurun_out/r02_build.log
rm -f gpurun_out/parity_large.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_gpu_tests.log 2>&1; tail -30 gpurun_out/r02_gpu_tests.log
timeout 300 python profiles/discover_torch_cuda_rounding.py > gpurun_out/r02_discover.log 2>&1; cat gpurun_out/r02_discover.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -5 gpurun_out/r02_bench_n1.err; cat gpurun_out/r02_bench_n1.json
