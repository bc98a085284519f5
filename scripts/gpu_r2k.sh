#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -x -q -m gpu -k "dino or loss or k65536" > gpurun_out/r2v_pytest.log 2>&1; tail -2 gpurun_out/r2v_pytest.log
for q in 1 0 1; do
ESVIT_CE_Q=$q timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_q$q.json 2> gpurun_out/r2v_bench_q$q.err
echo "CE_Q=$q rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2v_bench_q$q.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for r in [d['roofline']]+d['roofline_others']:
    if 'dino' in r['kernel'] or 'softmax' in r['kernel']: print('   ', r['kernel'][:60].ljust(60), round(r['frac'],3), round(r['ms_per_step'],3))"
done
