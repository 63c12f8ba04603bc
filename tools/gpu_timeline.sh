#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/train_timeline.py --out gpurun_out/r2_timeline_dp$N.json > gpurun_out/r2_timeline_dp$N.log 2>&1
tail -25 gpurun_out/r2_timeline_dp$N.log
