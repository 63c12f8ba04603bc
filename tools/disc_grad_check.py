"""Gradient parity of the CUDA discriminator backward vs torch autograd through the fp32 oracle (diagnostic)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ducosy_oracle as orc
from ducosy_gan_b200.modules.model import Discriminator

def x_(seed, shape):
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).uniform(-1, 1, size=shape).astype(np.float32))

for prec in ("fp16", "bf16"):
    os.environ["DUCOSY_PRECISION"] = prec
    for wstd in (0.02, 0.1):
        B, H, W = 2, 256, 256
        sd = orc.make_state_dict(orc.discriminator_param_shapes(1), 21, weight_std=wstd)
        D = Discriminator(1); D.load_state_dict(sd); D = D.cuda().train()
        x = x_(41, (B, 1, H, W))
        xg = x.clone().cuda().requires_grad_(True)
        out = D(xg)
        loss = torch.nn.functional.mse_loss(out, torch.ones_like(out)); loss.backward()
        ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        ro = orc.discriminator_forward(ref_sd, xr)
        rl = orc.mse_gan_loss(ro, True); rl.backward()
        rec = {"precision": prec, "weight_std": wstd, "out_rel_err": ((out.detach().cpu() - ro.detach()).abs().max() / ro.detach().abs().max()).item()}
        for name, p in D.named_parameters():
            g, r = p.grad.cpu(), ref_sd[name].grad
            if r.abs().max() < 1e-6: continue
            rec[name] = [round(((g - r).abs().max() / r.abs().max()).item(), 4),
                         round(1 - torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item(), 6)]
        gx, rx = xg.grad.cpu(), xr.grad
        rec["input"] = [round(((gx - rx).abs().max() / rx.abs().max()).item(), 4),
                        round(1 - torch.nn.functional.cosine_similarity(gx.flatten(), rx.flatten(), dim=0).item(), 6)]
        print(json.dumps(rec), flush=True)
