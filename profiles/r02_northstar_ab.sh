# A/B of the north-star traversal structure against the default kernel (one B200):
#   default            k_irradiance_lean<4,FRONT>: thread per ray, QBVH2 nodes through L1
#   TEXIR_TOP_SMEM=n   the first n nodes (top levels, breadth first) staged in shared memory by TMA (cp.async.bulk + mbarrier)
#   TEXIR_COOP8=1      warp-cooperative walk over 8-wide nodes, eight lanes per ray
# bench lines (full headline workload, 3 timed steps) + one ncu --set full capture each at P = 32768 texels.
set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-mat --no-cold > gpurun_out/r02_ab_$name.json 2> gpurun_out/r02_ab_$name.err || tail -5 gpurun_out/r02_ab_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_ab_$name.json"))
    print("AB %-14s %.4e ray-samples/s  kernel %.2f ms  checksum %.6f  sm %s MHz" % ("$name", d["value"], d["roofline"]["kernel_ms"], d["e2e"]["checksum"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("AB $name FAILED", e)
PY
}
run default TEXIR_TOP_SMEM=0
run top127 TEXIR_TOP_SMEM=127
run top255 TEXIR_TOP_SMEM=255
run top511 TEXIR_TOP_SMEM=511
run top1023 TEXIR_TOP_SMEM=1023
run coop8 TEXIR_COOP8=1
cap() { # name, kernel regex, env...
  name=$1; rx=$2; shift; shift
  env "$@" timeout 900 ncu --set full --import-source on --clock-control none -k regex:$rx -c 1 -o gpurun_out/r02_ncu_$name -f python bench.py --points 32768 --steps 1 --warmup 0 --no-cpu --no-mat --no-cold > gpurun_out/r02_ncu_$name.log 2>&1
  python profiles/summarize_ncu.py gpurun_out/r02_ncu_$name.ncu-rep gpurun_out/r02_ncu_$name 67108864 > gpurun_out/r02_ncu_$name.txt 2>&1; head -40 gpurun_out/r02_ncu_$name.txt
}
cap default k_irradiance_lean TEXIR_TOP_SMEM=0
cap top255 k_irradiance_top TEXIR_TOP_SMEM=255
cap coop8 k_irradiance_coop8 TEXIR_COOP8=1
ncu -i gpurun_out/r02_ncu_default.ncu-rep --page source --csv > gpurun_out/r02_ncu_default_source.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_top255.ncu-rep gpurun_out/r02_ncu_coop8.ncu-rep
ls -la gpurun_out | tail -20
