#!/bin/bash
# 2-GPU bench with the output / gradient checks and the step-time breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err
tail -2 gpurun_out/r2_final_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'strong', d['strong']['value'], d['strong']['efficiency_vs_n1'], d['strong']['sharded_output_equals_single_gpu_output'])
print('train', d['train']['value'], d['train']['ms_per_step'], d['train'].get('breakdown'))
print('weak', (d.get('train_weak') or {}).get('value'))
print('checks', d['checks'])
P
