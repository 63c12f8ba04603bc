#!/bin/bash
# round-2 GPU pass 1: full GPU test suite, 1-GPU bench, batch-1 profiles
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/r2_tests.log
python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
python tools/train_bench.py --batch 1 --steps 5 --profile --out gpurun_out/r2_train_b1.json > gpurun_out/r2_train_b1.log 2>&1
python tools/latency_b1.py > gpurun_out/r2_latency_b1.json 2> gpurun_out/r2_latency_b1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_b1.csv python tools/latency_b1.py > gpurun_out/r2_ncu_b1.log 2>&1
tail -5 gpurun_out/r2_tests.log
