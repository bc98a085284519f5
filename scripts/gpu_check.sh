#!/bin/bash
# Run on the GPU box: parity tests + smoke, logs into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_ops_gpu test_model_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/$f.log
  echo "== $f: $(tail -1 gpurun_out/$f.log)"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
