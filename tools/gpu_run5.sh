#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -x 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2_tests5.log
tail -5 gpurun_out/r2_tests5.log
python tools/latency_b1.py > gpurun_out/r2_latency_b1_v5.json 2> gpurun_out/r2_latency_b1_v5.err
python bench.py > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
cat gpurun_out/r2_latency_b1_v5.json
