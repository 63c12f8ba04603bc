"""``Adam`` with the update fused into one kernel per tensor (drop-in for ``torch.optim.Adam`` as the reference
constructs it at modules/trainer.py:360-362: ``Adam(params, lr=..., betas=(0.5, 0.999))``).  Subclasses
``torch.optim.Optimizer``, so ``zero_grad``, ``param_groups``, ``state_dict`` and ``LambdaLR`` (trainer.py:364-366) work."""
from __future__ import annotations

import torch

from ._lib import call, ptr, stream_ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("the fused kernel implements weight_decay=0, amsgrad=False (what the reference uses)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("ducosy_gan_b200.optim.Adam needs contiguous fp32 CUDA parameters (no CPU path exists)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                g = p.grad.contiguous()
                with torch.cuda.device(p.device):
                    call("ducosy_adam_step", ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]),
                         float(b1), float(b2), float(group["eps"]), int(st["step"]), stream_ptr())
                # the kernel wrote through the raw pointer: tell autograd / the packed-weight caches (keyed by _version)
                torch.autograd.graph.increment_version(p)
        return loss
