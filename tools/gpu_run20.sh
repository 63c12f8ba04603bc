#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gen_backward.py tests/test_gpu_kernels.py tests/test_gpu_train_step.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/train_bench.py --batch 8 --steps 10 --graph --out gpurun_out/r2_train_b8_v11.json > gpurun_out/r2_train_b8_v11.log 2>&1
timeout 300 python tools/train_bench.py --batch 1 --steps 20 --graph --out gpurun_out/r2_train_b1_v11.json > gpurun_out/r2_train_b1_v11.log 2>&1
grep -H ms_per_step gpurun_out/r2_train_b8_v11.json gpurun_out/r2_train_b1_v11.json
timeout 300 python tools/train_bench.py --batch 8 --steps 3 --profile 2>&1 | grep -A24 "kernel time total" | head -28
