set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 600 python examples/pipeline_synthetic.py > gpurun_out/r02_example.log 2>&1; tail -4 gpurun_out/r02_example.log
timeout 600 python examples/pipeline_synthetic.py --graphed > gpurun_out/r02_example_graphed.log 2>&1; tail -4 gpurun_out/r02_example_graphed.log
