#!/bin/bash
# launch list + ncu --set full summaries of one 30-slice chunk; the .ncu-rep files stay on the box (64 MiB pull limit):
# their raw pages are exported as CSV into gpurun_out/
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export DUCOSY_SINGLE_STREAM=1
python tools/prof_infer.py > gpurun_out/r2_prof_plain.log 2>&1 || { tail -5 gpurun_out/r2_prof_plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_infer.csv python tools/prof_infer.py > gpurun_out/r2_ncu_l.log 2>&1
timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:"out_conv7x7|stem_fused|stem_input|dewindow|in_apply_pad|residual_apply|cbam_pool|in_finalize|cbam_channel_mlp" -c 40 -o /tmp/r2_prof_bw python tools/prof_infer.py > gpurun_out/r2_ncu_bw.log 2>&1
ncu -i /tmp/r2_prof_bw.ncu-rep --page raw --csv > gpurun_out/r2_prof_bw_raw.csv 2>/dev/null
timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:"conv_gemm_kernel" -s 17 -c 6 -o /tmp/r2_prof_conv python tools/prof_infer.py > gpurun_out/r2_ncu_conv.log 2>&1
ncu -i /tmp/r2_prof_conv.ncu-rep --page raw --csv > gpurun_out/r2_prof_conv_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/*.ncu-rep
du -sh gpurun_out
