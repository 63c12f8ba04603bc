#!/bin/bash
# device-resident synthesis throughput vs slices per generator call (300-slice volume)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for b in 30 50 60 75 100 30; do
  AB_BATCH=$b AB_STEPS=3 AB_ONLY_DEFAULT=1 timeout 300 python tools/infer_ab.py 2>&1 | grep slices_per_s | sed "s/^/batch_slices=$b /"
done
