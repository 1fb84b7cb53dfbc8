set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -2 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_matstage.py tests/test_gpu_texture.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_gpu_tests4.log 2>&1; tail -12 gpurun_out/r02_gpu_tests4.log
timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace4.log 2>&1; head -28 gpurun_out/r02_train_trace4.log | cut -c1-230
TEXIR_FRONT=0 timeout 300 python profiles/r02_train_step_trace.py > gpurun_out/r02_train_trace4_nofront.log 2>&1; head -12 gpurun_out/r02_train_trace4_nofront.log | cut -c1-230
