set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests35.log 2>&1; tail -4 gpurun_out/r02_gpu_tests35.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
( time timeout 1200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2> gpurun_out/r02_bench_default.time; tail -3 gpurun_out/r02_bench_default.time; tail -3 gpurun_out/r02_bench_default.err
( time timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err ) 2> gpurun_out/r02_bench_reference.time; tail -3 gpurun_out/r02_bench_reference.time; cat gpurun_out/r02_bench_reference.json | cut -c1-400
timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace35.log 2>&1; head -22 gpurun_out/r02_train_trace35.log | cut -c1-75,150-215
timeout 600 python examples/pipeline_synthetic.py > gpurun_out/r02_example.log 2>&1; tail -6 gpurun_out/r02_example.log
