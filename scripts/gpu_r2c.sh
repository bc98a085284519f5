#!/bin/bash
# round-2 check C: GEMM2 with quad epilogues: tests per tile config, micro-bench, model tests + bench with ESVIT_GEMM2=1
mkdir -p gpurun_out
for t in 1128 1256 2128 2256; do
  timeout 600 python -m pytest tests/test_gemm2_gpu.py -q -m gpu -k "$t or old_entry" -p no:cacheprovider > gpurun_out/r2c_gemm2_$t.log 2>&1
  echo "gemm2 tile $t rc=$? : $(tail -1 gpurun_out/r2c_gemm2_$t.log)"
done
timeout 900 python scripts/bench_gemm2.py > gpurun_out/r2c_gemm2_bench.txt 2>&1
echo "gemm bench rc=$?"; tail -3 gpurun_out/r2c_gemm2_bench.txt
ESVIT_GEMM2=1 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_real_shapes_gpu.py tests/test_ops_gpu.py -q -m gpu > gpurun_out/r2c_pytest_gemm2.log 2>&1
echo "pytest GEMM2 rc=$?"; tail -15 gpurun_out/r2c_pytest_gemm2.log
ESVIT_GEMM2=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --profile gpurun_out/r2c_prof_gemm2.txt > gpurun_out/r2c_bench_gemm2.json 2> gpurun_out/r2c_bench_gemm2.err
echo "bench GEMM2 rc=$?"
tail -c 1200 gpurun_out/r2c_bench_gemm2.json
tail -5 gpurun_out/r2c_bench_gemm2.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --profile gpurun_out/r2c_prof_lib.txt > gpurun_out/r2c_bench_lib.json 2> gpurun_out/r2c_bench_lib.err
echo "bench lib rc=$?"
tail -c 600 gpurun_out/r2c_bench_lib.json
# tcgen05 attention forward core (ESVIT_ATTN_TC=1): block-level parity tests + micro-benchmark next to the mma.sync kernel
ESVIT_ATTN_TC=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "swin_block and not w14" > gpurun_out/r2c_pytest_attn_tc.log 2>&1
echo "pytest ATTN_TC rc=$?"; tail -12 gpurun_out/r2c_pytest_attn_tc.log
ESVIT_ATTN_TC=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2c_attn_tc.txt 2>&1
ESVIT_ATTN_ONLY0=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2c_attn_base.txt 2>&1
tail -3 gpurun_out/r2c_attn_tc.txt; tail -3 gpurun_out/r2c_attn_base.txt
