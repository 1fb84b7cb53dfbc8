set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
timeout 900 python profiles/r02_gbuffer_ab.py > gpurun_out/r02_gbuffer_ab.txt 2>&1; cat gpurun_out/r02_gbuffer_ab.txt
timeout 600 python profiles/r02_cold_breakdown.py > gpurun_out/r02_cold_breakdown.txt 2>&1; tail -4 gpurun_out/r02_cold_breakdown.txt
timeout 600 ncu --set full --clock-control none -k regex:"k_shade_lean|k_gbuffer_mvp" -c 2 -o gpurun_out/r02_ncu_step16 -f python profiles/r02_train_step_trace.py > gpurun_out/r02_ncu_step16.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02_ncu_step16.ncu-rep gpurun_out/r02_ncu_step16 1572864 > gpurun_out/r02_ncu_step16.txt 2>&1; grep -E "^void|^==|gpu__time|dram__bytes|thread_inst_executed_per|issue_active|pipe_alu|pipe_fma|pipe_lsu|warps_active|registers|derived|stall|long_sc|l1tex__t_sector_hit|lts__t_sector_hit" gpurun_out/r02_ncu_step16.txt | cut -c1-200
rm -f gpurun_out/r02_ncu_step16.ncu-rep
