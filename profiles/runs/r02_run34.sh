set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
# final-kernel ncu captures (IrT kernel; the three render kernels of a 256-spp stash + recompute step)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_irradiance_lean -c 1 -o gpurun_out/r02_ncu_final -f python bench.py --points 32768 --steps 1 --warmup 0 --no-cpu --no-mat --no-cold > gpurun_out/r02_ncu_final.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02_ncu_final.ncu-rep gpurun_out/r02_ncu_final 67108864 > gpurun_out/r02_ncu_final.txt 2>&1; head -45 gpurun_out/r02_ncu_final.txt | cut -c1-200
ncu -i gpurun_out/r02_ncu_final.ncu-rep --page source --csv > gpurun_out/r02_ncu_final_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:k_shade_lean -c 3 -o gpurun_out/r02_ncu_shade_final -f python profiles/r02_shade_step.py > gpurun_out/r02_ncu_shade_final.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02_ncu_shade_final.ncu-rep gpurun_out/r02_ncu_shade_final 25165824 > gpurun_out/r02_ncu_shade_final.txt 2>&1; grep -E "^void|gpu__time|dram__bytes|thread_inst_executed_per|issue_active|derived" gpurun_out/r02_ncu_shade_final.txt | cut -c1-260
rm -f gpurun_out/r02_ncu_shade_final.ncu-rep
