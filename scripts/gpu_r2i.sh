#!/bin/bash
# round-2 check I: tcgen05 attention backward with two threads per row: parity + micro-bench
mkdir -p gpurun_out
ESVIT_ATTN_TC=2 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "swin_block and not w14" > gpurun_out/r2i_pytest_attn_tc2.log 2>&1
echo "pytest ATTN_TC=2 rc=$?"; tail -6 gpurun_out/r2i_pytest_attn_tc2.log
ESVIT_ATTN_TC=2 timeout 300 python scripts/bench_attn.py > gpurun_out/r2i_attn_tc.txt 2>&1
tail -3 gpurun_out/r2i_attn_tc.txt
ESVIT_ATTN_ONLY0=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2i_attn_mma.txt 2>&1
tail -3 gpurun_out/r2i_attn_mma.txt
