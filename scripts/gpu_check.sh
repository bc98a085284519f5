#!/bin/bash
# Run on the GPU box: parity tests + smoke + bench (+ optional ncu launch list), logs into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$1" != "nobench_tests" ]; then
timeout 240 python -m pytest tests/test_gemm_tcgen05_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/test_gemm.log
echo "== tcgen05 gemm: $(tail -1 gpurun_out/test_gemm.log)"; grep -E "tcgen05 gemm\+" gpurun_out/test_gemm.log
for f in test_ops_gpu test_model_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/$f.log
  echo "== $f: $(tail -1 gpurun_out/$f.log)"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
fi
timeout 900 python bench.py --steps 10 --warmup 3 --profile gpurun_out/prof_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: $(tail -c 2500 gpurun_out/bench.json)"; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
  echo "== ncu rows: $(wc -l < gpurun_out/launches.csv)"
fi
