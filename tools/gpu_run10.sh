#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train_step.py tests/test_gpu_gen_backward.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4
python tools/train_bench.py --batch 1 --steps 20 --graph --out gpurun_out/r2_train_b1_graph.json > gpurun_out/r2_train_b1_graph.log 2>&1
DUCOSY_TRAIN_STREAMS=1 python tools/train_bench.py --batch 1 --steps 20 --graph --out gpurun_out/r2_train_b1_graph_s1.json > gpurun_out/r2_train_b1_graph_s1.log 2>&1
python tools/train_bench.py --batch 8 --steps 10 --graph --out gpurun_out/r2_train_b8_graph.json > gpurun_out/r2_train_b8_graph.log 2>&1
python tools/train_bench.py --batch 1 --steps 5 --profile > gpurun_out/r2_train_b1_v10.log 2>&1
grep ms_per_step gpurun_out/r2_train_b1_graph.json gpurun_out/r2_train_b1_graph_s1.json gpurun_out/r2_train_b8_graph.json
grep -A30 "kernel time total" gpurun_out/r2_train_b1_v10.log | head -32
