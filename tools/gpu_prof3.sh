#!/bin/bash
# ncu --set full of the training-tail kernels of round 2 (eager batch-8 step), raw pages exported as CSV; launch list of one step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/train_bench.py --batch 8 --steps 1 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 2700 --csv --log-file gpurun_out/r2_launches_train.csv $CMD > gpurun_out/r2_ncu_tl.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"out_conv_dgrad|out_conv_wgrad_mma|dgrad_s1_edge_cols|in_bwd_reduce|in_bwd_apply|pad_fold|wgrad_reduce_kernel|cbam_bwd_dv" -s 200 -c 36 -o /tmp/r2_prof_train $CMD > gpurun_out/r2_ncu_train.log 2>&1
ncu -i /tmp/r2_prof_train.ncu-rep --page raw --csv > gpurun_out/r2_prof_train_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"conv_wgrad_kernel" -s 40 -c 4 -o /tmp/r2_prof_wg $CMD > gpurun_out/r2_ncu_wg.log 2>&1
ncu -i /tmp/r2_prof_wg.ncu-rep --page raw --csv > gpurun_out/r2_prof_wg_raw.csv 2>/dev/null
ls -la gpurun_out; tail -2 gpurun_out/r2_ncu_train.log
