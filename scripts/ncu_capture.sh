#!/bin/bash
# GPU box: reference-eager baseline + ncu evidence (launch list of one eager step, full captures of the top kernels)
mkdir -p gpurun_out
NCU_BENCH="python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"
timeout 400 python bench.py --impl reference --device cuda --steps 3 --warmup 1 > gpurun_out/ref_cuda.json 2> gpurun_out/ref_cuda.err
echo "== ref cuda: $(tail -c 600 gpurun_out/ref_cuda.json)"; tail -3 gpurun_out/ref_cuda.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $NCU_BENCH > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
echo "== launch list rows: $(wc -l < gpurun_out/launches.csv)"
for k in window_attn_bwd7 window_attn_fwd7 dino_ce_bwd gelu_bwd_dbias add_ln_bwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -f -o gpurun_out/full_$k $NCU_BENCH > /dev/null 2> gpurun_out/full_$k.err
  echo "== full $k: $(ls -la gpurun_out/full_$k.ncu-rep 2>/dev/null | awk '{print $5}')"
done
