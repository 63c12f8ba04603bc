#!/bin/bash
# A/B: in_apply_pad compiled for 4 CTAs per SM (64 registers, DUCOSY_APPLY_CAP=1) vs uncapped (66 registers, 3 CTAs per SM)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for c in 1 0 1 0; do
  DUCOSY_APPLY_CAP=$c AB_STEPS=3 AB_ONLY_DEFAULT=1 timeout 300 python tools/infer_ab.py 2>&1 | grep slices_per_s | sed "s/^/apply_cap=$c batch30 /"
done
