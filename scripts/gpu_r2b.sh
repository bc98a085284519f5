#!/bin/bash
# round-2 check B: real-shape parity tests, optimizer test, second-generation GEMM tests per tile config, GEMM micro-bench,
# bench with gpu_reference + parity_check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2b_gpu.txt 2>&1
nproc >> gpurun_out/r2b_gpu.txt
for t in 1128 1256 2128 2256; do
  timeout 600 python -m pytest tests/test_gemm2_gpu.py -q -m gpu -k "$t or old_entry" -p no:cacheprovider > gpurun_out/r2b_gemm2_$t.log 2>&1
  echo "gemm2 tile $t rc=$? : $(tail -1 gpurun_out/r2b_gemm2_$t.log)"
done
timeout 900 python scripts/bench_gemm2.py > gpurun_out/r2b_gemm2_bench.txt 2>&1
echo "gemm bench rc=$?"; tail -5 gpurun_out/r2b_gemm2_bench.txt
timeout 900 python -m pytest tests/test_real_shapes_gpu.py tests/test_ops_gpu.py::test_fused_adamw_clip_ema_matches_reference_sequence -q -m gpu > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"
tail -c 2500 gpurun_out/r2b_bench.json
tail -5 gpurun_out/r2b_bench.err
# the same model-level parity tests and the bench with every Linear on the second-generation GEMM family
ESVIT_GEMM2=1 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_real_shapes_gpu.py -q -m gpu > gpurun_out/r2b_pytest_gemm2.log 2>&1
echo "pytest GEMM2 rc=$?"; tail -15 gpurun_out/r2b_pytest_gemm2.log
ESVIT_GEMM2=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --profile gpurun_out/r2b_prof_gemm2.txt > gpurun_out/r2b_bench_gemm2.json 2> gpurun_out/r2b_bench_gemm2.err
echo "bench GEMM2 rc=$?"
tail -c 1500 gpurun_out/r2b_bench_gemm2.json
tail -5 gpurun_out/r2b_bench_gemm2.err
