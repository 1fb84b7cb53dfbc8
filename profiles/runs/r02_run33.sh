set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 1500 python profiles/sweep_variants.py --mat TEXIR_FRONT=1 shtnear0 shtnear2 TEXIR_FRONT=1 shtnear0 shtnear2 > gpurun_out/r02_shade_tnear_sweep.txt 2>&1; cat gpurun_out/r02_shade_tnear_sweep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_matstage.py tests/test_gpu_texture.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests33.log 2>&1; tail -3 gpurun_out/r02_gpu_tests33.log | cut -c1-300
