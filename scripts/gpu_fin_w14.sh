#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --arch swin_small_w14 --local-crops 10 --batch 32 --steps 8 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/fin_w14_small.json 2> gpurun_out/fin_w14_small.err
echo "rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/fin_w14_small.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'])"; tail -2 gpurun_out/fin_w14_small.err
