#!/bin/bash
# ncu source-level captures of the tcgen05 attention kernels (stage-0 global-crop geometry)
mkdir -p gpurun_out
python scripts/ncu_kernels.py --list > gpurun_out/ncu_cases.json
ESVIT_ATTN_TC=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:window_attn_fwd7_tc -s 2 -c 1 -f -o gpurun_out/k_attn_fwd7_tc python scripts/ncu_kernels.py attn_fwd7_s0 > gpurun_out/k_attn_fwd7_tc.log 2>&1
ESVIT_ATTN_TC=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:window_attn_bwd7_tc -s 2 -c 1 -f -o gpurun_out/k_attn_bwd7_tc python scripts/ncu_kernels.py attn_bwd7_s0 > gpurun_out/k_attn_bwd7_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
