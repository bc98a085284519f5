#!/bin/bash
# GPU box: W=14 attention fast path - parity (block tests), micro-benchmark, Swin-B W14 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "swin_block" 2>&1 | tail -60 > gpurun_out/test_blocks.log
echo "== blocks: $(tail -1 gpurun_out/test_blocks.log)"; grep -E "^FAILED|Error|assert" gpurun_out/test_blocks.log | head -20
timeout 300 python scripts/bench_attn14.py > gpurun_out/attn14.txt 2>&1; tail -12 gpurun_out/attn14.txt
timeout 600 python bench.py --arch swin_base_w14 --batch 32 --steps 5 --warmup 3 --no-cpu-baseline --profile gpurun_out/prof_b14.txt > gpurun_out/bench_b14.json 2> gpurun_out/bench_b14.err
echo "== bench B w14: $(head -c 600 gpurun_out/bench_b14.json)"; tail -3 gpurun_out/bench_b14.err
