#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
ESVIT_PROFILE_STACKS=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --profile gpurun_out/r2t_prof.txt > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
echo rc=$?; python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('clocks'))"
grep -v "^      " gpurun_out/r2t_prof.txt.stacks.txt | head -30
