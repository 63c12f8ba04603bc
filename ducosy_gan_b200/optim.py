"""``Adam`` with the update fused into three launches per parameter group (drop-in for ``torch.optim.Adam`` as the reference
constructs it at modules/trainer.py:360-362: ``Adam(params, lr=..., betas=(0.5, 0.999))``).  Subclasses
``torch.optim.Optimizer``, so ``zero_grad``, ``param_groups``, ``state_dict`` and ``LambdaLR`` (trainer.py:364-366) work.

The learning rate and the step count live in device memory, so a whole optimisation step (forward, backward, all-reduce,
update) can be captured in a CUDA graph and replayed: nothing that changes between steps is baked into a launch.
``group["lr"]`` is pushed to the device whenever it changed (outside the graph).

Non-finite guard (``check_finite=True``, the default): the training path stores gradient maps in 16 bit, which can overflow
where the reference's fp32 autograd cannot.  A step whose gradients contain Inf/NaN is skipped on the device, the way
``torch.amp.GradScaler.step`` skips it -- parameters, moments and the step count stay untouched; ``skipped_steps()`` reports
how often that happened.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import call, ptr, stream_ptr

ADAM_CHUNK = 16384      # include/ducosy.h DUCOSY_ADAM_CHUNK
_STATE_FLOATS = 8       # lr, step, flag, skipped, skipped-now, reserved x3


class _AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("n", C.c_longlong)]


class _AdamChunk(C.Structure):
    _fields_ = [("tensor", C.c_int), ("reserved", C.c_int), ("start", C.c_longlong)]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, capturable=False,
                 check_finite=True):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("the fused kernel implements weight_decay=0, amsgrad=False (what the reference uses)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = bool(capturable)   # kept for API compatibility: the device-side {lr, step} is always used
        self.check_finite = bool(check_finite)
        self._dev_state = {}     # group index -> float32 tensor [lr, step, flag, skipped, skipped-now, ...]
        self._dev_lr = {}
        self._tables = {}        # group index -> (signature, tensor table, chunk table, number of chunks)

    # -- device-side state ----------------------------------------------------------------------------------------
    def push_lr(self):
        """Copy ``group['lr']`` into the device state when it changed (call between graph replays after a scheduler step)."""
        for gi, group in enumerate(self.param_groups):
            st = self._dev_state.get(gi)
            if st is not None and self._dev_lr.get(gi) != float(group["lr"]):
                st[0:1].fill_(float(group["lr"]))
                self._dev_lr[gi] = float(group["lr"])

    def skipped_steps(self) -> int:
        """Steps the non-finite guard skipped so far (host sync)."""
        return sum(int(round(float(st[3].item()))) for st in self._dev_state.values())

    def _sync_host_steps(self):
        for gi, group in enumerate(self.param_groups):
            st = self._dev_state.get(gi)
            if st is not None:
                step = int(round(float(st[1].item())))
                for p in group["params"]:
                    if p in self.state and self.state[p]:
                        self.state[p]["step"] = step

    # -- checkpoint / resume (reference modules/trainer.py:394-396,586-588 save and restore the optimiser state_dicts) -----
    def state_dict(self):
        self._sync_host_steps()      # graph replays and skipped steps advance only the device-side count
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._dev_state, self._dev_lr, self._tables = {}, {}, {}
        for gi, group in enumerate(self.param_groups):
            steps = [int(self.state[p]["step"]) for p in group["params"] if p in self.state and self.state[p]]
            if steps:
                self._state_for(gi, group, group["params"][0].device, step=max(steps))

    def _state_for(self, gi, group, device, step=0):
        st = self._dev_state.get(gi)
        if st is None:
            st = torch.zeros(_STATE_FLOATS, dtype=torch.float32, device=device)
            st[0], st[1] = float(group["lr"]), float(step)
            self._dev_state[gi], self._dev_lr[gi] = st, float(group["lr"])
        return st

    def _table_for(self, gi, live, capturing):
        """Device tables of (param, grad, exp_avg, exp_avg_sq, n) and of the 16384-element chunks, rebuilt when any pointer
        changed (eager autograd allocates fresh ``.grad`` tensors; flat gradient buckets keep them fixed)."""
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                     p.numel()) for p in live)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == sig:
            return cached
        if capturing:
            raise RuntimeError("ducosy_gan_b200.optim.Adam: parameter / gradient storages changed during CUDA-graph capture; "
                               "run warm-up steps first and keep gradients in static buffers (data_parallel.GradBucket)")
        tensors = (_AdamTensor * len(sig))(*[_AdamTensor(*s) for s in sig])
        chunks = [(i, 0, start) for i, s in enumerate(sig) for start in range(0, s[4], ADAM_CHUNK)]
        ctab = (_AdamChunk * len(chunks))(*[_AdamChunk(*c) for c in chunks])
        dev = live[0].device
        to_dev = lambda arr: torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        cached = (sig, to_dev(tensors), to_dev(ctab), len(chunks))
        self._tables[gi] = cached
        return cached

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
            dev = live[0].device
            for p in live:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                    raise RuntimeError("ducosy_gan_b200.optim.Adam needs contiguous fp32 CUDA parameters on one device "
                                       "(no CPU path exists)")
                if not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                    p.grad = p.grad.to(torch.float32).contiguous()
            fresh = [p for p in live if not self.state[p]]
            if fresh:
                if capturing:
                    raise RuntimeError("ducosy_gan_b200.optim.Adam: optimiser state must exist before CUDA-graph capture (run a warm-up step)")
                # moments of the group in two flat buffers; state[p] holds views, so state_dict() keeps torch's layout
                n = sum((p.numel() + 3) // 4 * 4 for p in fresh)          # every view 16-byte aligned (float4 path)
                flat_m, flat_v = torch.zeros(n, dtype=torch.float32, device=dev), torch.zeros(n, dtype=torch.float32, device=dev)
                off = 0
                for p in fresh:
                    st = self.state[p]
                    st["step"] = 0
                    st["exp_avg"] = flat_m[off:off + p.numel()].view_as(p)
                    st["exp_avg_sq"] = flat_v[off:off + p.numel()].view_as(p)
                    off += (p.numel() + 3) // 4 * 4
            dev_state = self._state_for(gi, group, dev)
            if not capturing:
                self.push_lr()
            _, ttab, ctab, nchunks = self._table_for(gi, live, capturing)
            with torch.cuda.device(dev):
                call("ducosy_adam_multi_step", ptr(ttab), ptr(ctab), nchunks, ptr(dev_state), float(b1), float(b2), float(group["eps"]),
                     int(self.check_finite), stream_ptr())
            for p in live:
                self.state[p]["step"] += 1   # host copy (inspection); state_dict() re-reads the device count
                # the kernel wrote through the raw pointer: tell autograd / the packed-weight caches (keyed by _version)
                torch.autograd.graph.increment_version(p)
        return loss
