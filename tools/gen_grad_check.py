"""Per-tensor gradient error of the generator autograd bridge against fp32 torch autograd (prints a JSON report).
usage: python tools/gen_grad_check.py [Cin blocks use_cbam B H W]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_gpu_gen_backward as T  # noqa: E402

a = [int(v) for v in sys.argv[1:]] or [1, 2, 0, 1, 64, 512]
for aware in (True, False):
    rep = T._check_generator_grads(a[0], a[1], bool(a[2]), a[3], a[4], a[5], seed=5, rounding_aware=aware)
    print("rounding-aware reference" if aware else "fp32 reference")
    print(json.dumps({k: round(v, 5) for k, v in rep.items()}))

dt = torch.bfloat16 if os.environ.get("DUCOSY_PRECISION", "fp16").lower() == "bf16" else torch.float16
print("torch model with 16-bit stored activations and gradient maps vs fp32 autograd (noise floor)")
print(json.dumps({k: round(v, 5) for k, v in T._noise_floor(a[0], a[1], bool(a[2]), a[3], a[4], a[5], 5, dt).items()}))
