#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generator.py -m gpu -q -x -k "split" -s 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
