#!/bin/bash
# A/B: conv weight gradients accumulated straight into p.grad (DUCOSY_WGRAD_DIRECT=1, default) vs returned to autograd (0)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_gen_backward.py -m gpu -q -x 2>&1 | tail -3
for w in 1 0 1 0; do
  for b in 1 8; do
    DUCOSY_WGRAD_DIRECT=$w timeout 300 python tools/train_bench.py --batch $b --steps 20 --graph --out gpurun_out/r2_train_b${b}_wd$w.json > gpurun_out/r2_train_b${b}_wd$w.log 2>&1
    echo "wgrad_direct=$w batch=$b $(grep ms_per_step gpurun_out/r2_train_b${b}_wd$w.json) $(grep '"G"' gpurun_out/r2_train_b${b}_wd$w.json)"
  done
done
