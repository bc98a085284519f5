#!/bin/bash
# round-2 check N: tcgen05 attention backward as the default: whole GPU suite, smoke, full bench line, ncu of the attention
# kernels, launch list of the eager step, attention micro-benchmarks (mma.sync vs tcgen05 on the same box)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2n_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --profile gpurun_out/r2n_prof.txt > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; tail -c 700 gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
ESVIT_ATTN_TC=0 ESVIT_ATTN_ONLY0=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2n_attn_mma.txt 2>&1
ESVIT_ATTN_TC=2 timeout 300 python scripts/bench_attn.py > gpurun_out/r2n_attn_tc_bwd.txt 2>&1
ESVIT_ATTN_TC=3 timeout 300 python scripts/bench_attn.py > gpurun_out/r2n_attn_tc_both.txt 2>&1
tail -1 gpurun_out/r2n_attn_mma.txt; tail -1 gpurun_out/r2n_attn_tc_bwd.txt; tail -1 gpurun_out/r2n_attn_tc_both.txt
KEEP_REPS="attn_bwd7_s0" bash scripts/ncu_capture_r2.sh attn_fwd7_s0 attn_bwd7_s0 > gpurun_out/r2n_ncu.log 2>&1; tail -3 gpurun_out/r2n_ncu.log
NCU_BENCH="python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline --no-gpu-reference"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $NCU_BENCH > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
echo "== launch list rows: $(wc -l < gpurun_out/launches.csv)"; du -sh gpurun_out
