set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python profiles/sweep_variants.py TEXIR_FRONT=1 TEXIR_L2_PERSIST=2 TEXIR_FRONT=1 TEXIR_L2_PERSIST=1 > gpurun_out/r02_l2persist_sweep.txt 2>&1; cat gpurun_out/r02_l2persist_sweep.txt
