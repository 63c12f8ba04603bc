"""nn.DataParallel wrapping as the reference does it on a multi-GPU box (modules/trainer.py:335-338): forward and backward of
the wrapped Generator / Discriminator against the unwrapped modules on one GPU (same weights, same batch).  Needs >= 2 GPUs."""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ducosy_gan_b200.modules.model import Discriminator, Generator, weights_init_normal  # noqa: E402

assert torch.cuda.device_count() >= 2
torch.manual_seed(0)
res = {"devices": torch.cuda.device_count()}
G = Generator(2, 2, True).cuda().apply(weights_init_normal)
D = Discriminator(1).cuda().apply(weights_init_normal)
x = (torch.rand(4, 2, 256, 512, device="cuda") * 2 - 1)
img = (torch.rand(4, 1, 256, 256, device="cuda") * 2 - 1)
with torch.no_grad():
    ref_g, ref_d = G(x), D(img)
Gp, Dp = nn.DataParallel(G), nn.DataParallel(D)
with torch.no_grad():
    out_g, out_d = Gp(x), Dp(img)
res["forward_G_equal"] = bool(torch.equal(out_g, ref_g))
res["forward_D_equal"] = bool(torch.equal(out_d, ref_d))
# backward: gradients reduced onto the master copy by DataParallel's Broadcast
target = torch.rand_like(ref_g)
G.zero_grad(); (G(x) - target).abs().mean().backward()
g_single = [p.grad.clone() for p in G.parameters()]
G.zero_grad(); (Gp(x) - target).abs().mean().backward()
g_dp = [p.grad.clone() for p in G.parameters()]
num = sum(((a - b) ** 2).sum() for a, b in zip(g_dp, g_single)).sqrt()
den = sum((b ** 2).sum() for b in g_single).sqrt()
res["backward_G_rel_l2"] = float(num / den)
D.zero_grad(); (D(img) ** 2).mean().backward()
d_single = [p.grad.clone() for p in D.parameters()]
D.zero_grad(); (Dp(img) ** 2).mean().backward()
d_dp = [p.grad.clone() for p in D.parameters()]
num = sum(((a - b) ** 2).sum() for a, b in zip(d_dp, d_single)).sqrt()
den = sum((b ** 2).sum() for b in d_single).sqrt()
res["backward_D_rel_l2"] = float(num / den)
res["module_access"] = type(Gp.module).__name__            # trainer.py:193-196 unwraps with .module
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dataparallel_check.json", "w"), indent=1)
