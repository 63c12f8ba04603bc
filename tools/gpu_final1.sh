#!/bin/bash
# final single-GPU pass: parity suite, smoke, the bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err; cut -c1-200 gpurun_out/r2_final_bench.json
