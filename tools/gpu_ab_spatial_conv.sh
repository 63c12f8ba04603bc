#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generator.py tests/test_gpu_kernels.py tests/test_gpu_synthesis.py -m gpu -q -x 2>&1 | tail -4
for s in 1 0 1 0; do
  DUCOSY_FUSED_SPATIAL=$s AB_STEPS=3 AB_ONLY_DEFAULT=1 timeout 300 python tools/infer_ab.py 2>&1 | grep slices_per_s | sed "s/^/fused_spatial=$s batch30 /"
done
for s in 1 0; do
  DUCOSY_FUSED_SPATIAL=$s AB_BATCH=1 AB_SLICES=30 AB_STEPS=5 AB_ONLY_DEFAULT=1 timeout 300 python tools/infer_ab.py 2>&1 | grep slices_per_s | sed "s/^/fused_spatial=$s batch1 /"
done
