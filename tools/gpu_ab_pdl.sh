#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for p in 1 0; do
  DUCOSY_PDL=$p timeout 300 python tools/latency_b1.py > gpurun_out/r2_latency_pdl$p.json 2> gpurun_out/r2_latency_pdl$p.err
  DUCOSY_PDL=$p timeout 300 python tools/train_bench.py --batch 1 --steps 20 --graph --out gpurun_out/r2_train_b1_pdl$p.json > gpurun_out/r2_train_b1_pdl$p.log 2>&1
  DUCOSY_PDL=$p timeout 300 python tools/train_bench.py --batch 8 --steps 10 --graph --out gpurun_out/r2_train_b8_pdl$p.json > gpurun_out/r2_train_b8_pdl$p.log 2>&1
  DUCOSY_PDL=$p AB_STEPS=3 AB_ONLY_DEFAULT=1 timeout 300 python tools/infer_ab.py > gpurun_out/r2_infer_pdl$p.json 2>&1
done
grep -H "back_to_back\|generate_py" gpurun_out/r2_latency_pdl*.json
grep -H ms_per_step gpurun_out/r2_train_b*_pdl*.json
grep -H -A2 '"default"' gpurun_out/r2_infer_pdl*.json
