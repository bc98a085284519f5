#!/bin/bash
# round-2 check A: new real-shape parity tests, optimizer test, bench with gpu_reference + parity_check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_real_shapes_gpu.py tests/test_ops_gpu.py::test_fused_adamw_clip_ema_matches_reference_sequence -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
