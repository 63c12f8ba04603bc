"""``Adam`` with the update fused into one kernel per tensor (drop-in for ``torch.optim.Adam`` as the reference
constructs it at modules/trainer.py:360-362: ``Adam(params, lr=..., betas=(0.5, 0.999))``).  Subclasses
``torch.optim.Optimizer``, so ``zero_grad``, ``param_groups``, ``state_dict`` and ``LambdaLR`` (trainer.py:364-366) work.

``capturable=True`` keeps the learning rate and the step count in device memory, so a whole optimisation step
(forward, backward, all-reduce, update) can be captured in a CUDA graph and replayed: nothing that changes between steps
is baked into a launch.  ``group["lr"]`` is pushed to the device whenever it changed (outside the graph).
"""
from __future__ import annotations

import torch

from ._lib import call, ptr, stream_ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, capturable=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("the fused kernel implements weight_decay=0, amsgrad=False (what the reference uses)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = bool(capturable)
        self._dev_state = {}     # group index -> float32 tensor [lr, step] (capturable mode)
        self._dev_lr = {}

    def push_lr(self):
        """Copy ``group['lr']`` into the device state when it changed (call between graph replays after a scheduler step)."""
        for gi, group in enumerate(self.param_groups):
            st = self._dev_state.get(gi)
            if st is not None and self._dev_lr.get(gi) != float(group["lr"]):
                st[0:1].fill_(float(group["lr"]))
                self._dev_lr[gi] = float(group["lr"])

    # -- checkpoint / resume (reference modules/trainer.py:394-396,586-588 save and restore the optimiser state_dicts) -----
    def state_dict(self):
        if self.capturable:          # graph replays advance only the device-side count: bring the host copies up to date
            for gi, group in enumerate(self.param_groups):
                st = self._dev_state.get(gi)
                if st is not None:
                    step = int(round(float(st[1].item())))
                    for p in group["params"]:
                        if p in self.state and self.state[p]:
                            self.state[p]["step"] = step
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._dev_state, self._dev_lr = {}, {}
        if self.capturable:
            for gi, group in enumerate(self.param_groups):
                steps = [int(self.state[p]["step"]) for p in group["params"] if p in self.state and self.state[p]]
                if steps:
                    dev = group["params"][0].device
                    self._dev_state[gi] = torch.tensor([float(group["lr"]), float(max(steps))], dtype=torch.float32, device=dev)
                    self._dev_lr[gi] = float(group["lr"])

    def _state_for(self, gi, group, device):
        st = self._dev_state.get(gi)
        if st is None:
            st = torch.tensor([float(group["lr"]), 0.0], dtype=torch.float32, device=device)
            self._dev_state[gi], self._dev_lr[gi] = st, float(group["lr"])
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            dev_state = None
            if self.capturable:
                dev_state = self._state_for(gi, group, live[0].device)
                if not capturing:
                    self.push_lr()
                with torch.cuda.device(live[0].device):
                    call("ducosy_adam_advance", ptr(dev_state), stream_ptr())
            for p in live:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("ducosy_gan_b200.optim.Adam needs contiguous fp32 CUDA parameters (no CPU path exists)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1          # host copy (state_dict / inspection); the capturable kernels read the device count
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                with torch.cuda.device(p.device):
                    if dev_state is not None:
                        call("ducosy_adam_step_dev", ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), ptr(dev_state),
                             float(b1), float(b2), float(group["eps"]), stream_ptr())
                    else:
                        call("ducosy_adam_step", ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]),
                             float(b1), float(b2), float(group["eps"]), int(st["step"]), stream_ptr())
                # the kernel wrote through the raw pointer: tell autograd / the packed-weight caches (keyed by _version)
                torch.autograd.graph.increment_version(p)
        return loss
