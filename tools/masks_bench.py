"""Times ducosy_gan_b200.mask_generator.generate_anatomical_masks (all four masks of modules/mask_generator.py) on batches of
512 x 512 HU slices resident on the GPU, next to the oracle (scipy restatement pinned to the reference) on the same slices."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ducosy_gan_b200 import mask_generator as mg
from oracle import ducosy_oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hu = orc.mask_test_slices_bone(8, 512, 512, seed=3)
dev = torch.from_numpy(np.concatenate([hu] * ((B + 7) // 8))[:B]).cuda()
for _ in range(2):
    mg.generate_anatomical_masks(dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    out = mg.generate_anatomical_masks(dev)
torch.cuda.synchronize()
gpu = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
lung = orc.mask_detect_lung(hu)
orc.mask_detect_mediastinum(hu, lung); orc.mask_detect_bone(hu, lung); orc.mask_detect_lung_vessels(hu, lung)
cpu = time.perf_counter() - t0
print(json.dumps({"batch": B, "gpu_ms_per_batch": gpu * 1e3, "gpu_slices_per_s": B / gpu, "cpu_oracle_slices_per_s_one_core": 8 / cpu,
                  "pixels": {k: int(v.sum()) for k, v in out.items()}}))
