#!/bin/bash
# round-2 check F: tcgen05 attention with pipelined gathers: parity + micro-bench + step
mkdir -p gpurun_out
ESVIT_ATTN_TC=2 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "swin_block and not w14" > gpurun_out/r2f_pytest_attn_tc2.log 2>&1
echo "pytest ATTN_TC=2 rc=$?"; tail -6 gpurun_out/r2f_pytest_attn_tc2.log
ESVIT_ATTN_TC=2 timeout 300 python scripts/bench_attn.py > gpurun_out/r2f_attn_tc.txt 2>&1
tail -2 gpurun_out/r2f_attn_tc.txt
ESVIT_ATTN_TC=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_tc.json 2> gpurun_out/r2f_bench_tc.err
echo "bench TC rc=$?"; tail -c 300 gpurun_out/r2f_bench_tc.json
