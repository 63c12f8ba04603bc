#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2_tests3.log
python tools/latency_b1.py > gpurun_out/r2_latency_b1_v3.json 2> gpurun_out/r2_latency_b1_v3.err
python tools/train_bench.py --batch 1 --steps 5 --profile --out gpurun_out/r2_train_b1_v3.json > gpurun_out/r2_train_b1_v3.log 2>&1
python tools/train_bench.py --batch 8 --steps 5 --profile --out gpurun_out/r2_train_b8_v3.json > gpurun_out/r2_train_b8_v3.log 2>&1
tail -8 gpurun_out/r2_tests3.log
