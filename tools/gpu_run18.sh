#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/outconv_bench.py 30 > gpurun_out/r2_outconv_variants.json 2>&1; cat gpurun_out/r2_outconv_variants.json
for v in 0 2; do
DUCOSY_OUTCONV_LEAN=$v DUCOSY_SINGLE_STREAM=1 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:"out_conv7x7|dewindow" -c 2 -o /tmp/r2_prof_out$v python tools/prof_infer.py > gpurun_out/r2_ncu_out$v.log 2>&1
ncu -i /tmp/r2_prof_out$v.ncu-rep --page raw --csv > gpurun_out/r2_prof_out${v}_raw.csv 2>/dev/null
done
ls -la gpurun_out
