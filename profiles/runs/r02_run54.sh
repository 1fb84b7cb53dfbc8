set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests54.log 2>&1; tail -3 gpurun_out/r02_gpu_tests54.log | cut -c1-300
for k in 1 2 3; do timeout 300 python profiles/r02_train_step_trace.py 2>/dev/null | grep "^== 16 spp"; done
