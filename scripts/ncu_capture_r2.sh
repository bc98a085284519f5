#!/bin/bash
# GPU box: `ncu --set full` of ONE launch of every hot kernel at a known shape (scripts/ncu_kernels.py) + the launch list
# of two eager training steps.  usage: ncu_capture_r2.sh [cases...]   (default: all)
mkdir -p gpurun_out
CASES=${@:-"gemm_fwd_qkv0 gemm_gelu_fc1_0 gemm_mul_fc2dgrad_0 gemm_wgrad_qkv0 gemm_fwd_fc2_2 gemm_fwd_lastlayer attn_fwd7_s0 attn_bwd7_s0 add_ln_fwd_96 add_ln_bwd_96 dino_ce_fwd dino_ce_bwd patch_embed_fwd patch_embed_bwd region_match"}
python scripts/ncu_kernels.py --list > gpurun_out/ncu_cases.json
for c in $CASES; do
  k=$(python -c "import json;print(json.load(open('gpurun_out/ncu_cases.json'))['$c']['kernel'])")
  s=2; case $c in dino_ce_fwd) s=5;; dino_ce_bwd) s=4;; esac   # the region-row launch (the cls launch runs beside it)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/k_$c python scripts/ncu_kernels.py $c > gpurun_out/k_$c.log 2>&1
  # gpurun brings back at most 64 MiB: keep the raw-metric CSV of every capture, the .ncu-rep only for KEEP_REPS
  ncu -i gpurun_out/k_$c.ncu-rep --page raw --csv > gpurun_out/k_$c.csv 2>/dev/null
  echo "== $c ($k): $(ls -la gpurun_out/k_$c.ncu-rep 2>/dev/null | awk '{print $5}') bytes, csv $(wc -c < gpurun_out/k_$c.csv)"
  case " ${KEEP_REPS:-attn_bwd7_s0 gemm_gelu_fc1_0} " in *" $c "*) ;; *) rm -f gpurun_out/k_$c.ncu-rep;; esac
done
