#!/bin/bash
# round-2 final check: whole GPU suite, smoke, full default bench line (reference arms, cpu baseline), ncu of the kernels
# changed since r02_v3 (CSV only), launch list of the eager step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/fin_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/fin_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/fin_smoke.log
timeout 900 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/fin_bench.json; tail -2 gpurun_out/fin_bench.err
KEEP_REPS="none" bash scripts/ncu_capture_r2.sh attn_fwd7_s0 add_ln_bwd_96 gemm_gelu_fc1_0 dino_ce_fwd dino_ce_bwd > gpurun_out/fin_ncu.log 2>&1; tail -5 gpurun_out/fin_ncu.log
NCU_BENCH="python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline --no-gpu-reference"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $NCU_BENCH > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
echo "== launch list rows: $(wc -l < gpurun_out/launches.csv)"; du -sh gpurun_out
