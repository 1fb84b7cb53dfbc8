set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_rgbe_conf.py tests/test_gpu_build.py tests/test_gpu_matstage.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests8.log 2>&1; tail -15 gpurun_out/r02_gpu_tests8.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n1d.json 2> gpurun_out/r02_bench_n1d.err; tail -5 gpurun_out/r02_bench_n1d.err; cat gpurun_out/r02_bench_n1d.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_shade_lean -c 3 -o gpurun_out/r02_ncu_shade -f python profiles/r02_shade_step.py > gpurun_out/r02_ncu_shade.log 2>&1; tail -3 gpurun_out/r02_ncu_shade.log
python profiles/summarize_ncu.py gpurun_out/r02_ncu_shade.ncu-rep gpurun_out/r02_ncu_shade 25165824 > gpurun_out/r02_ncu_shade.txt 2>&1; grep -E "^void|gpu__time|dram__bytes|thread_inst_executed_per|issue_active|pipe_alu|derived" gpurun_out/r02_ncu_shade.txt | cut -c1-260
rm -f gpurun_out/r02_ncu_shade.ncu-rep
