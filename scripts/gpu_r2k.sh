#!/bin/bash
mkdir -p gpurun_out
ESVIT_ATTN_ONLY0=1 ESVIT_ATTN_TC=0 timeout 300 python scripts/bench_attn.py > gpurun_out/r2m_attn_mma.txt 2>&1; tail -1 gpurun_out/r2m_attn_mma.txt
for dbg in 0 64 128; do
ESVIT_ATTN_TC=3 ESVIT_ATTN_DBG=$dbg timeout 300 python scripts/bench_attn.py > gpurun_out/r2m_attn_tc3_$dbg.txt 2>&1
echo "dbg $dbg"; tail -1 gpurun_out/r2m_attn_tc3_$dbg.txt
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
ESVIT_ATTN_TC=3 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "swin_block and not w14" 2>&1 | tail -2
