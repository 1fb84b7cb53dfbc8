set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -2 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r02_gpu_tests5a.log 2>&1; tail -15 gpurun_out/r02_gpu_tests5a.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_large.py > gpurun_out/r02_gpu_tests5.log 2>&1; tail -8 gpurun_out/r02_gpu_tests5.log
timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace5.log 2>&1; head -26 gpurun_out/r02_train_trace5.log | cut -c1-75,150-215
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n1c.json 2> gpurun_out/r02_bench_n1c.err; tail -5 gpurun_out/r02_bench_n1c.err; cat gpurun_out/r02_bench_n1c.json
