#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/infer_ab.py > gpurun_out/r2_infer_ab.json 2>&1
cat gpurun_out/r2_infer_ab.json
python tools/infer_ab.py AB_BATCH=1 AB_SLICES=30 > gpurun_out/r2_infer_ab_b1.json 2>&1
cat gpurun_out/r2_infer_ab_b1.json
python tools/latency_b1.py > gpurun_out/r2_latency_b1_v6.json 2> gpurun_out/r2_latency_b1_v6.err
cat gpurun_out/r2_latency_b1_v6.json
DUCOSY_FUSED_FINALIZE=0 DUCOSY_FUSED_SPATIAL=0 python tools/latency_b1.py > gpurun_out/r2_latency_b1_v6_unfused.json 2>/dev/null
cat gpurun_out/r2_latency_b1_v6_unfused.json
python -m pytest tests/test_gpu_train_step.py tests/test_gpu_gen_backward.py -m gpu -q -x 2>&1 | tail -5
