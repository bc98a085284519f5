#!/bin/bash
mkdir -p gpurun_out
for occ in 3 2 3 2; do
ESVIT_ADDLN_OCC=$occ timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_occ$occ.json 2> gpurun_out/r2p_bench_occ$occ.err
echo "bench OCC=$occ rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2p_bench_occ$occ.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for r in [d['roofline']]+d['roofline_others']:
    if 'add_ln' in r['kernel']: print('   ', r['kernel'][:40], round(r['frac'],3), round(r['ms_per_step'],3))"
done
