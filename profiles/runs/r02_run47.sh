set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_matstage.py tests/test_gpu_texture.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests47.log 2>&1; tail -3 gpurun_out/r02_gpu_tests47.log | cut -c1-300
timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace47.log 2>&1; head -24 gpurun_out/r02_train_trace47.log | cut -c1-75,150-215
TEXIR_FOLD_WIDE=0 timeout 300 python profiles/r02_train_step_trace.py 2>/dev/null | head -1
