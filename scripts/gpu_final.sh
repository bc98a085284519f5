#!/bin/bash
# GPU box: the driver's own commands on the final tree, then the opt-in fused-groups path (parity + A/B bench)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/final_pytest.log; echo "== pytest -m gpu: $(tail -1 gpurun_out/final_pytest.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "== bench: $(head -c 400 gpurun_out/bench_final.json)"; tail -2 gpurun_out/bench_final.err
ESVIT_FUSE_GROUPS=1 timeout 200 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/fused_groups_pytest.log; echo "== fused groups parity: $(tail -1 gpurun_out/fused_groups_pytest.log)"; grep -E "^FAILED|Error" gpurun_out/fused_groups_pytest.log | head -5
ESVIT_FUSE_GROUPS=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile gpurun_out/prof_fused_groups.txt > gpurun_out/bench_fused_groups.json 2> gpurun_out/bench_fused_groups.err; echo "== bench fused groups: $(head -c 400 gpurun_out/bench_fused_groups.json)"; tail -2 gpurun_out/bench_fused_groups.err
