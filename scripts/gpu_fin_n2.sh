#!/bin/bash
# 2-GPU sanity of the final tree: the bench line at N = 2 (NCCL, bucketed gradient reduction inside the captured graph)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
timeout 240 $RUN --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/fin_n2_bench.json 2> gpurun_out/fin_n2_bench.err
echo "bench N=2 rc=$?"; tail -c 900 gpurun_out/fin_n2_bench.json; tail -3 gpurun_out/fin_n2_bench.err
