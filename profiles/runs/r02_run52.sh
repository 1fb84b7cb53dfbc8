set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_matstage.py -m gpu -q -x -p no:cacheprovider -k graphed > gpurun_out/r02_gpu_tests52.log 2>&1; tail -12 gpurun_out/r02_gpu_tests52.log | cut -c1-250
timeout 600 python examples/pipeline_synthetic.py --graphed > gpurun_out/r02_example_graphed.log 2>&1; tail -4 gpurun_out/r02_example_graphed.log
