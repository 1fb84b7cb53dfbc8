set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -2 gpurun_out/r02_build.log
timeout 300 python profiles/discover_torch_cuda_rounding.py > gpurun_out/r02_discover2.log 2>&1; head -120 gpurun_out/r02_discover2.log
timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace.log 2>&1; tail -80 gpurun_out/r02_train_trace.log
timeout 600 python -m pytest tests/test_gpu_matstage.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_gpu_tests2.log 2>&1; tail -5 gpurun_out/r02_gpu_tests2.log
