#!/bin/bash
mkdir -p gpurun_out
ESVIT_ATTN_TC=2 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "swin_block and not w14" 2>&1 | tail -8
: > gpurun_out/r2j_prof.txt
for ngw in 2 3; do for dbg in 0 1; do
  echo "=== NGW=$ngw DBG=$dbg" >> gpurun_out/r2j_prof.txt
  ESVIT_ATTN_NGW=$ngw ESVIT_ATTN_DBG=$dbg timeout 120 python scripts/prof_attn_tc.py 2>&1 | grep -A4 "H=56 shift=3 rep 1" >> gpurun_out/r2j_prof.txt
done; done
for ngw in 2 3; do for dbg in 0 1; do
ESVIT_ATTN_TC=2 ESVIT_ATTN_NGW=$ngw ESVIT_ATTN_DBG=$dbg timeout 300 python scripts/bench_attn.py > gpurun_out/r2j_attn_tc_ngw${ngw}_$dbg.txt 2>&1
echo "ngw $ngw dbg $dbg"; tail -1 gpurun_out/r2j_attn_tc_ngw${ngw}_$dbg.txt
done; done
cat gpurun_out/r2j_prof.txt
