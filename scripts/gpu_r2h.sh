#!/bin/bash
# round-2 check H: whole GPU suite, smoke, full bench line, ncu of the CE kernels (image-major CTA order)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2h_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --profile gpurun_out/r2h_prof.txt > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
bash scripts/ncu_capture_r2.sh dino_ce_fwd dino_ce_bwd > gpurun_out/r2h_ncu.log 2>&1; tail -3 gpurun_out/r2h_ncu.log
