"""BASELINE.json config 3: PatchGAN discriminator forward/backward, batch 8 at 512x512 (32x32 patch output) with the
MSE adversarial loss -- CUDA path through the drop-in module + autograd bridge.  Writes gpurun_out/disc_bench.json."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ducosy_gan_b200.modules.model import Discriminator, weights_init_normal

torch.manual_seed(2)
D = Discriminator(1).apply(weights_init_normal).cuda().train()
x = torch.rand(8, 1, 512, 512, device="cuda") * 2 - 1
valid = torch.ones(8, 1, 32, 32, device="cuda")
def step():
    for p in D.parameters():
        p.grad = None
    loss = torch.nn.functional.mse_loss(D(x), valid)
    loss.backward()
    return loss
for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
with torch.no_grad():
    for _ in range(3): D(x)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): D(x)
    e1.record(); torch.cuda.synchronize()
fwd_ms = e0.elapsed_time(e1) / n
rec = {"config": "Discriminator fwd+bwd, batch 8, 512x512, MSE-GAN loss", "ms_fwd_bwd": ms, "ms_fwd_only": fwd_ms,
       "steps_per_s": 1e3 / ms, "nominal_gflop_fwd_bwd": 312.9, "tflops_nominal": 312.9 / ms}
print(json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "disc_bench.json"), "w"), indent=1)
