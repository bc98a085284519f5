#!/bin/bash
# round-2 check E: tcgen05 attention forward v2 + backward (ESVIT_ATTN_TC=2), bench line, ncu captures (CSV), launch list
mkdir -p gpurun_out
ESVIT_ATTN_TC=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "swin_block and not w14" > gpurun_out/r2e_pytest_attn_tc1.log 2>&1
echo "pytest ATTN_TC=1 rc=$?"; tail -4 gpurun_out/r2e_pytest_attn_tc1.log
ESVIT_ATTN_TC=2 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "swin_block and not w14" > gpurun_out/r2e_pytest_attn_tc2.log 2>&1
echo "pytest ATTN_TC=2 rc=$?"; tail -12 gpurun_out/r2e_pytest_attn_tc2.log
ESVIT_ATTN_TC=2 timeout 300 python scripts/bench_attn.py > gpurun_out/r2e_attn_tc.txt 2>&1
ESVIT_ATTN_ONLY0=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2e_attn_base.txt 2>&1
tail -2 gpurun_out/r2e_attn_tc.txt; tail -2 gpurun_out/r2e_attn_base.txt
timeout 900 python bench.py --steps 20 --warmup 3 --profile gpurun_out/r2e_prof.txt > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
ESVIT_ATTN_TC=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r2e_bench_tc.json 2> gpurun_out/r2e_bench_tc.err
echo "bench TC rc=$?"; tail -c 600 gpurun_out/r2e_bench_tc.json
bash scripts/ncu_capture_r2.sh > gpurun_out/r2e_ncu.log 2>&1; tail -16 gpurun_out/r2e_ncu.log
NCU_BENCH="python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline --no-gpu-reference"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $NCU_BENCH > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
echo "== launch list rows: $(wc -l < gpurun_out/launches.csv)"; du -sh gpurun_out
