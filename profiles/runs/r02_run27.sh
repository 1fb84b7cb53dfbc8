set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python -m pytest tests/test_gpu_matstage.py -m gpu -q -x -p no:cacheprovider -k "graphed" > gpurun_out/r02_gpu_tests27.log 2>&1; tail -5 gpurun_out/r02_gpu_tests27.log | cut -c1-300
timeout 600 python profiles/r02_sanitizer_step.py > gpurun_out/r02_sanitizer_plain.log 2>&1; tail -2 gpurun_out/r02_sanitizer_plain.log
timeout 1500 compute-sanitizer --tool memcheck python profiles/r02_sanitizer_step.py > gpurun_out/r02_compute_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_compute_sanitizer_memcheck.txt; tail -4 gpurun_out/r02_compute_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python profiles/r02_sanitizer_step.py > gpurun_out/r02_compute_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_compute_sanitizer_racecheck.txt; tail -4 gpurun_out/r02_compute_sanitizer_racecheck.txt
