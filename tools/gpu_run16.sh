#!/bin/bash
# 2-GPU bench: DP all-reduce overlap on (default) and off
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for o in 1 0; do
DUCOSY_DP_OVERLAP=$o python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$o bench.py --gpus 2 --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_n2_overlap$o.json 2> gpurun_out/r2_bench_n2_overlap$o.err
tail -2 gpurun_out/r2_bench_n2_overlap$o.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_n2_overlap$o.json').read().strip().splitlines()[-1])
print('overlap=$o', 'value', d['value'], 'train ms', d['train']['ms_per_step'], 'weak ms', (d.get('train_weak') or {}).get('ms_per_step'), 'checks', d['checks'].get('train'))
P
done
