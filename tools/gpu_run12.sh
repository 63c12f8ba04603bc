#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/accuracy_report.py > gpurun_out/r2_accuracy.log 2>&1; tail -3 gpurun_out/r2_accuracy.log | cut -c1-600
python tools/resblock_sweep.py > gpurun_out/r2_resblock_sweep.log 2>&1; grep fp16x2 gpurun_out/r2_resblock_sweep.log
DUCOSY_PRECISION=fp16x2 python bench.py --steps 2 --warmup 3 --train-steps 0 --skip-cpu-baseline > gpurun_out/r2_bench_fp16x2.json 2> gpurun_out/r2_bench_fp16x2.err; cut -c1-400 gpurun_out/r2_bench_fp16x2.json; tail -3 gpurun_out/r2_bench_fp16x2.err
