#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_$i.json 2> gpurun_out/r2q_bench_$i.err
echo "bench rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2q_bench_$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('clocks'))
for r in [d['roofline']]+d['roofline_others']:
    print('   ', r['kernel'][:60].ljust(60), round(r['frac'],3), round(r['ms_per_step'],3))"
done
timeout 300 python scripts/bench_gemm2.py > gpurun_out/r2q_gemm2.txt 2>&1; tail -30 gpurun_out/r2q_gemm2.txt
