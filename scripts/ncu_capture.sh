#!/bin/bash
# GPU box: ncu evidence (launch list of eager steps + optional full captures).  usage: ncu_capture.sh [list|full|all]
mkdir -p gpurun_out
MODE=${1:-all}
NCU_BENCH="python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
if [ "$MODE" != "full" ]; then
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches.csv $NCU_BENCH > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
echo "== launch list rows: $(wc -l < gpurun_out/launches.csv)"
fi
if [ "$MODE" != "list" ]; then
# -s: skip the small late-stage launches so that the capture hits stage-0/1 sized ones.  gemm_bias_act:58 is one of the
# last (stage-0) act=2 launches of the first backward: the fused fc2-dgrad * gelu' + column-sum epilogue
for spec in ${NCU_KERNELS:-"gemm_bias_act:58 window_attn_bwd7:20 window_attn_fwd7:2 dino_ce_bwd:1 add_ln_bwd:30"}; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/full_${k}_s$s $NCU_BENCH > /dev/null 2> gpurun_out/full_$k.err
  echo "== full $k: $(ls -la gpurun_out/full_${k}_s$s.ncu-rep 2>/dev/null | awk '{print $5}')"
done
fi
