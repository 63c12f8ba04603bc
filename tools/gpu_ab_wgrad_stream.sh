#!/bin/bash
# A/B: weight gradients on a side stream (DUCOSY_WGRAD_STREAM=1, default) vs in line (0); graph-replayed CycleGAN step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_gen_backward.py -m gpu -q -x 2>&1 | tail -3
for w in 1 0 1 0; do
  for b in 1 8; do
    DUCOSY_WGRAD_STREAM=$w timeout 300 python tools/train_bench.py --batch $b --steps 20 --graph --out gpurun_out/r2_train_b${b}_ws$w.json > gpurun_out/r2_train_b${b}_ws$w.log 2>&1
    echo "wgrad_stream=$w batch=$b $(grep ms_per_step gpurun_out/r2_train_b${b}_ws$w.json)"
  done
done
