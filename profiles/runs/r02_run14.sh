set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python profiles/r02_gbuffer_ab.py > gpurun_out/r02_gbuffer_ab.txt 2>&1; cat gpurun_out/r02_gbuffer_ab.txt
timeout 900 python -m pytest tests/test_gpu_matstage.py tests/test_gpu_texture.py tests/test_gpu_rgbe_conf.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests14.log 2>&1; tail -5 gpurun_out/r02_gpu_tests14.log | cut -c1-300
