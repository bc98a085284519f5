#!/bin/bash
# GPU box: the driver's own commands on the final tree (+ the per-group loop as the A/B: ESVIT_FUSE_GROUPS=0)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/final_pytest.log; echo "== pytest -m gpu: $(tail -1 gpurun_out/final_pytest.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
timeout 300 python bench.py --profile gpurun_out/prof_table.txt > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "== bench: $(head -c 420 gpurun_out/bench_final.json)"; tail -2 gpurun_out/bench_final.err
if [ "$1" == "ab" ]; then
ESVIT_FUSE_GROUPS=0 timeout 200 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3; 
ESVIT_FUSE_GROUPS=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_per_group.json 2> gpurun_out/bench_per_group.err; echo "== bench per-group loop: $(head -c 400 gpurun_out/bench_per_group.json)"
fi
