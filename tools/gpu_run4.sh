#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2_tests4.log
tail -5 gpurun_out/r2_tests4.log
for s in 1 2; do
  DUCOSY_TRAIN_STREAMS=$s python tools/train_bench.py --batch 1 --steps 10 --out gpurun_out/r2_train_b1_s$s.json > gpurun_out/r2_train_b1_s$s.log 2>&1
  DUCOSY_TRAIN_STREAMS=$s python tools/train_bench.py --batch 8 --steps 5 --out gpurun_out/r2_train_b8_s$s.json > gpurun_out/r2_train_b8_s$s.log 2>&1
done
python tools/train_bench.py --batch 1 --steps 5 --profile > gpurun_out/r2_train_b1_v4.log 2>&1
grep ms_per_step gpurun_out/r2_train_b*_s*.json
