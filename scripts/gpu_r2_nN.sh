#!/bin/bash
# N-GPU box (N = $1): weak-scaling bench lines of the headline config and configs 3 / 5 (Swin-S / Swin-B W14, 2+10 crops)
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N"
if [ "$N" = "1" ]; then RUN="python bench.py --gpus 1 --no-cpu-baseline"; fi
timeout 600 $RUN --steps 30 --warmup 3 > gpurun_out/n${N}_bench_tiny.json 2> gpurun_out/n${N}_bench_tiny.err
echo "bench tiny N=$N rc=$?"; tail -c 700 gpurun_out/n${N}_bench_tiny.json
timeout 600 $RUN --steps 20 --warmup 3 --arch swin_small_w14 --local-crops 10 --batch 32 ${REF3:---no-gpu-reference} > gpurun_out/n${N}_bench_small14.json 2> gpurun_out/n${N}_bench_small14.err
echo "bench small_w14 N=$N rc=$?"; tail -c 400 gpurun_out/n${N}_bench_small14.json
timeout 600 $RUN --steps 20 --warmup 3 --arch swin_base_w14 --local-crops 10 --batch 32 ${REF5:---no-gpu-reference} > gpurun_out/n${N}_bench_base14.json 2> gpurun_out/n${N}_bench_base14.err
echo "bench base_w14 N=$N rc=$?"; tail -c 400 gpurun_out/n${N}_bench_base14.json
