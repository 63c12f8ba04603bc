#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_generator.py tests/test_gpu_train_step.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/r2_tests2.log
python tools/latency_b1.py > gpurun_out/r2_latency_b1_v2.json 2> gpurun_out/r2_latency_b1_v2.err
tail -15 gpurun_out/r2_tests2.log
