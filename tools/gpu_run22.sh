#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/train_bench.py --batch 8 --steps 10 --graph --out gpurun_out/r2_train_b8_v12.json > gpurun_out/r2_train_b8_v12.log 2>&1
timeout 300 python tools/train_bench.py --batch 1 --steps 20 --graph --out gpurun_out/r2_train_b1_v12.json > gpurun_out/r2_train_b1_v12.log 2>&1
grep -H ms_per_step gpurun_out/r2_train_b8_v12.json gpurun_out/r2_train_b1_v12.json
