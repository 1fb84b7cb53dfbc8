set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1; tail -1 gpurun_out/r02_build.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; grep -E "NVLS|nranks|Connected all" gpurun_out/r02_bench_n8.err | head -5; tail -3 gpurun_out/r02_bench_n8.err | cut -c1-300; cat gpurun_out/r02_bench_n8.json
timeout 1200 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --tris 4194304 --tex 8192 > gpurun_out/r02_bench_n8_configs3.json 2> gpurun_out/r02_bench_n8_configs3.err; tail -3 gpurun_out/r02_bench_n8_configs3.err | cut -c1-300; cat gpurun_out/r02_bench_n8_configs3.json
timeout 1500 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --tris 8388608 --tex 8192 > gpurun_out/r02_bench_n8_configs4.json 2> gpurun_out/r02_bench_n8_configs4.err; tail -3 gpurun_out/r02_bench_n8_configs4.err | cut -c1-300; cat gpurun_out/r02_bench_n8_configs4.json
grep -E "AllReduce|AllGather" gpurun_out/r02_bench_n8.err | awk '{print $0}' | cut -c1-200 | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02_n8_nccl_collectives.txt; cat gpurun_out/r02_n8_nccl_collectives.txt
