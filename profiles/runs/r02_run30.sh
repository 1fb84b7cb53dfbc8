set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 1200 python profiles/sweep_variants.py --mat TEXIR_FRONT=1 pfpush pfpush2 pfleaf pfpush2leaf TEXIR_FRONT=1 > gpurun_out/r02_prefetch_sweep.txt 2>&1; cat gpurun_out/r02_prefetch_sweep.txt
