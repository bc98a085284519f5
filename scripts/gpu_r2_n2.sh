#!/bin/bash
# 2-GPU box: NCCL parity test of the multi-rank path + N=2 bench lines (headline config, and configs 3 / 5)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu > gpurun_out/n2_pytest.log 2>&1
echo "dist pytest rc=$?"; tail -5 gpurun_out/n2_pytest.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
timeout 600 $RUN --steps 30 --warmup 3 > gpurun_out/n2_bench_tiny.json 2> gpurun_out/n2_bench_tiny.err
echo "bench tiny N=2 rc=$?"; tail -c 700 gpurun_out/n2_bench_tiny.json
timeout 600 $RUN --steps 20 --warmup 3 --arch swin_small_w14 --local-crops 10 --batch 32 --no-gpu-reference > gpurun_out/n2_bench_small14.json 2> gpurun_out/n2_bench_small14.err
echo "bench small_w14 N=2 rc=$?"; tail -c 400 gpurun_out/n2_bench_small14.json
timeout 600 $RUN --steps 20 --warmup 3 --arch swin_base_w14 --local-crops 10 --batch 32 --no-gpu-reference > gpurun_out/n2_bench_base14.json 2> gpurun_out/n2_bench_base14.err
echo "bench base_w14 N=2 rc=$?"; tail -c 400 gpurun_out/n2_bench_base14.json
