set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 1500 python profiles/sweep_variants.py --mat TEXIR_FRONT=1 nopf pfuv pffront minb10 TEXIR_COHERENT_BITS=4 TEXIR_FRONT=1 > gpurun_out/r02_prefetch_sweep2.txt 2>&1; cat gpurun_out/r02_prefetch_sweep2.txt
