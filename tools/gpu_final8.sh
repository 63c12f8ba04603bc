#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_final_bench_n8.json 2> gpurun_out/r2_final_bench_n8.err
tail -3 gpurun_out/r2_final_bench_n8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n8.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'strong', d.get('strong'))
print('train', d['train']['value'], d['train']['ms_per_step'], d['train'].get('breakdown'))
print('weak', (d.get('train_weak') or {}).get('value'), (d.get('train_weak') or {}).get('ms_per_step'))
print('checks', d['checks'].get('train'))
P
