set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 600 python profiles/r02_sampler_ab.py > gpurun_out/r02_sampler_ab.txt 2>&1; tail -8 gpurun_out/r02_sampler_ab.txt
