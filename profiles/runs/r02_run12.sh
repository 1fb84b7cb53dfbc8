set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests12.log 2>&1; tail -8 gpurun_out/r02_gpu_tests12.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches.csv "-c 4000 python bench.py --steps 2 --warmup 1 --no-cpu" > gpurun_out/r02_launches_summary.txt; head -30 gpurun_out/r02_launches_summary.txt
