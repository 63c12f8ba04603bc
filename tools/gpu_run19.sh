#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gen_backward.py -m gpu -q -x -s 2>&1 | grep -E "out conv backward|passed|failed|Error|error" | head -20
timeout 600 python -m pytest tests/test_gpu_train_step.py -m gpu -q -x 2>&1 | tail -3
for m in 1 0; do
DUCOSY_OUTCONV_DGRAD_MMA=$m timeout 300 python tools/train_bench.py --batch 8 --steps 10 --graph --out gpurun_out/r2_train_b8_mma$m.json > gpurun_out/r2_train_b8_mma$m.log 2>&1
grep -H ms_per_step gpurun_out/r2_train_b8_mma$m.json
done
timeout 300 python tools/train_bench.py --batch 8 --steps 3 --profile 2>&1 | grep -A45 "kernel time total" | head -50
