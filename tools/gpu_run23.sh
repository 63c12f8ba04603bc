#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err; cut -c1-200 gpurun_out/r2_final_bench.json
timeout 300 python tools/train_bench.py --batch 8 --steps 3 --profile 2>&1 | grep -A12 "kernel time total" | head -14
python -c "import __graft_entry__ as g; g.smoke()"
