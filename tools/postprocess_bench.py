"""Post-composite smoothing (N1) throughput: GPU kernels vs the scipy pipeline of the reference on a bounded sample."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ducosy_gan_b200.postprocess import postprocess_volume  # noqa: E402
from oracle import ducosy_oracle as orc  # noqa: E402

S, H, W = 300, 512, 512
vol_h = orc.synthetic_volume(S, H, W, seed=0)
vol = torch.from_numpy(vol_h).cuda()
out = torch.empty_like(vol)
for _ in range(3):
    postprocess_volume(vol, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    postprocess_volume(vol, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
n = 40
t0 = time.perf_counter()
ref = orc.postprocess_volume(vol_h[:n])
cpu_s = time.perf_counter() - t0
# algorithmic traffic: int16 in (2) + v1 write/read (4+4+4) + pp write/read (4+4) + int16 out (2) = 28 B/voxel
res = {"volume": [S, H, W], "gpu_ms": ms, "gpu_slices_per_s": S / ms * 1e3, "algorithmic_GBps": S * H * W * 28 / ms / 1e6,
       "scipy_slices_per_s": n / cpu_s, "scipy_sample_slices": n, "cpu_cores": os.cpu_count()}
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/postprocess_bench.json", "w"), indent=1)
