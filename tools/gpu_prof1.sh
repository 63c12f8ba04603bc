#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --train-steps 0 --skip-cpu-baseline --skip-checks"
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 5500 -c 1800 --csv --log-file gpurun_out/r2_launches_infer.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"out_conv7x7|stem_fused|stem_input|dewindow|in_apply_pad|residual_apply|cbam_pool|in_finalize|cbam_channel_mlp" -s 600 -c 150 -o gpurun_out/r2_prof_bw $CMD > gpurun_out/r2_ncu_bw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel" -s 240 -c 50 -o gpurun_out/r2_prof_conv $CMD > gpurun_out/r2_ncu_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/r2_ncu_bw.log
