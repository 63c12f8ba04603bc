#!/bin/bash
# final single-GPU pass: parity suite, the bench line, launch list of the bench command (ncu), a compute-sanitizer attempt
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err; cut -c1-300 gpurun_out/r2_final_bench.json
CMD="python bench.py --steps 1 --warmup 3 --train-steps 0 --skip-cpu-baseline --skip-checks"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6040 -c 2000 --csv --log-file gpurun_out/r2_final_launches.csv $CMD > gpurun_out/r2_ncu_final.log 2>&1
tail -2 gpurun_out/r2_ncu_final.log
timeout 400 compute-sanitizer --tool memcheck --launch-timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/r2_sanitizer_memcheck.log
du -sh gpurun_out
