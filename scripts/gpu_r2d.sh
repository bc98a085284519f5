#!/bin/bash
# round-2 check D: whole GPU suite with every Linear on the tg2 GEMM family (default), full bench line, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --profile gpurun_out/r2d_prof.txt > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2d_smoke.log
bash scripts/ncu_capture_r2.sh > gpurun_out/r2d_ncu.log 2>&1; tail -16 gpurun_out/r2d_ncu.log
NCU_BENCH="python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline --no-gpu-reference"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $NCU_BENCH > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
echo "== launch list rows: $(wc -l < gpurun_out/launches.csv)"
