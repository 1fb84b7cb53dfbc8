set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -2 gpurun_out/r02_build.log
timeout 300 python profiles/discover_torch_cuda_rounding.py > gpurun_out/r02_discover3.log 2>&1; grep -E "torch.sum|generate_dir" gpurun_out/r02_discover3.log
rm -f gpurun_out/parity_large.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_gpu_tests3.log 2>&1; tail -25 gpurun_out/r02_gpu_tests3.log
cat gpurun_out/parity_large.jsonl
timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace3.log 2>&1; head -30 gpurun_out/r02_train_trace3.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; tail -5 gpurun_out/r02_bench_n1b.err; cat gpurun_out/r02_bench_n1b.json
