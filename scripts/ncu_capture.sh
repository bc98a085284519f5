#!/bin/bash
# GPU box: ncu evidence for the final kernels (launch list of one eager step + full captures of the top kernels)
mkdir -p gpurun_out
NCU_BENCH="python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches.csv $NCU_BENCH > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
echo "== launch list rows: $(wc -l < gpurun_out/launches.csv)"
# -s: skip the small late-stage launches so that the capture hits stage-0/1 sized ones
for spec in "gemm_bias_act:2" "window_attn_bwd7:20" "window_attn_fwd7:2" "dino_ce_bwd:1" "add_ln_bwd:30" "mul_bwd_dbias:10"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/full_$k $NCU_BENCH > /dev/null 2> gpurun_out/full_$k.err
  echo "== full $k: $(ls -la gpurun_out/full_$k.ncu-rep 2>/dev/null | awk '{print $5}')"
done
