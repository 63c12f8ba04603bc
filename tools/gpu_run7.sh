#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/infer_ab.py > gpurun_out/r2_infer_ab2.json 2>&1
cat gpurun_out/r2_infer_ab2.json
python tools/infer_ab.py AB_BATCH=1 AB_SLICES=30 > gpurun_out/r2_infer_ab2_b1.json 2>&1
cat gpurun_out/r2_infer_ab2_b1.json
python tools/latency_b1.py > gpurun_out/r2_latency_b1_v7.json 2> gpurun_out/r2_latency_b1_v7.err
cat gpurun_out/r2_latency_b1_v7.json
python -m pytest tests/test_gpu_generator.py tests/test_gpu_synthesis.py tests/test_gpu_train_step.py tests/test_gpu_gen_backward.py -m gpu -q -x 2>&1 | tail -5
