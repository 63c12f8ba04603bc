"""Times the fused stem (both passes) alone: python tools/stem_bench.py [B]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ducosy_gan_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10
px = torch.from_numpy(np.random.Generator(np.random.PCG64(0)).integers(0, 2500, size=(B, 512, 512), dtype=np.int16)).cuda()
wp = ops.pack_stem_weight(torch.randn(64, 1, 7, 7, device="cuda") * 0.02, torch.float16)
for _ in range(3):
    ops.stem_fused(wp, px=px, window=(1.0, -1024.0, -150.0, 250.0))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.stem_fused(wp, px=px, window=(1.0, -1024.0, -150.0, 250.0))
e1.record(); torch.cuda.synchronize()
print(json.dumps({"B": B, "us_per_stem_two_passes_plus_finalize": e0.elapsed_time(e1) * 100}))
