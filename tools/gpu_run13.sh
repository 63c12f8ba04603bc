#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_train_step.py -m gpu -q -x -k "metrics or validation" -s 2>&1 | tail -25
