"""Drop-ins for the loss modules of the reference's ``modules/trainer.py`` (file:line in each docstring).

Same class names, constructor arguments and ``forward`` signatures; forward AND backward run as fused kernels of
libducosy_sm100.so (csrc/loss.cu) through ``torch.autograd.Function`` bridges.  Gradients flow to the first argument
(the generated image) only -- which is how the reference uses them (targets / sources are real images).
Inputs: fp32 CUDA tensors ``[B,1,H,W]``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr, stream_ptr


def _scratch(dev):
    return torch.empty(_lib.load().ducosy_loss_scratch_bytes() // 4 + 4, dtype=torch.float32, device=dev)


def _prep(*ts):
    out = []
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("ducosy_gan_b200 losses need CUDA tensors (no CPU path exists)")
        out.append(t.detach().to(torch.float32).contiguous())
    return out


def _img_dims(t):
    if t.dim() != 4 or t.shape[1] != 1:
        raise RuntimeError(f"expected [B,1,H,W], got {tuple(t.shape)}")
    return t.shape[0], t.shape[2], t.shape[3]


class _L1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _prep(a, b)
        with torch.cuda.device(a.device):
            out = torch.empty((), dtype=torch.float32, device=a.device)
            call("ducosy_loss_l1_forward", ptr(a), ptr(b), a.numel(), ptr(out), ptr(_scratch(a.device)), stream_ptr())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(a.device):
            da = torch.empty_like(a)
            call("ducosy_loss_l1_backward", ptr(a), ptr(b), a.numel(), ptr(g), ptr(da), stream_ptr())
        return da, (-da if ctx.needs_input_grad[1] else None)


class _MSEConst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, target):
        (a,) = _prep(a)
        with torch.cuda.device(a.device):
            out = torch.empty((), dtype=torch.float32, device=a.device)
            call("ducosy_loss_mse_const_forward", ptr(a), float(target), a.numel(), ptr(out), ptr(_scratch(a.device)), stream_ptr())
        ctx.save_for_backward(a)
        ctx.target = float(target)
        return out

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        g = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(a.device):
            da = torch.empty_like(a)
            call("ducosy_loss_mse_const_backward", ptr(a), ctx.target, a.numel(), ptr(g), ptr(da), stream_ptr())
        return da, None


def l1_loss(a, b):
    """nn.L1Loss() as used for the cycle / identity terms (reference modules/trainer.py:348-349,469,482)."""
    return _L1.apply(a, b)


def mse_gan_loss(d_out, is_real: bool):
    """nn.MSELoss()(D(x), ones|zeros) (reference modules/trainer.py:347,459-460,470,518,523)."""
    return _MSEConst.apply(d_out, 1.0 if is_real else 0.0)


class _Gradient(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, t):
        p, t = _prep(p, t)
        B, H, W = _img_dims(p)
        with torch.cuda.device(p.device):
            out = torch.empty((), dtype=torch.float32, device=p.device)
            call("ducosy_loss_gradient_forward", ptr(p), ptr(t), B, H, W, ptr(out), ptr(_scratch(p.device)), stream_ptr())
        ctx.save_for_backward(p, t)
        return out

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        B, H, W = _img_dims(p)
        g = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(p.device):
            dp = torch.empty_like(p)
            call("ducosy_loss_gradient_backward", ptr(p), ptr(t), B, H, W, ptr(g), ptr(dp), stream_ptr())
        return dp, None


class GradientLoss(nn.Module):
    """reference modules/trainer.py:22-40."""

    def forward(self, pred, target):
        return _Gradient.apply(pred, target)


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, t, s, sigma, wmin, wmax):
        p, t, s = _prep(p, t, s)
        B, H, W = _img_dims(p)
        with torch.cuda.device(p.device):
            out = torch.empty((), dtype=torch.float32, device=p.device)
            umap = torch.empty_like(p)
            call("ducosy_loss_contrast_attention_forward", ptr(p), ptr(t), ptr(s), B, H, W, float(sigma), float(wmin), float(wmax),
                 ptr(out), ptr(umap), ptr(_scratch(p.device)), stream_ptr())
        ctx.save_for_backward(umap)
        return out

    @staticmethod
    def backward(ctx, g):
        (umap,) = ctx.saved_tensors
        B, H, W = _img_dims(umap)
        g = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(umap.device):
            dp = torch.empty_like(umap)
            call("ducosy_loss_contrast_attention_backward", ptr(umap), B, H, W, ptr(g), ptr(dp), stream_ptr())
        return dp, None, None, None, None, None


class ContrastAttentionLoss(nn.Module):
    """reference modules/trainer.py:43-86 (blur_kernel 7 as constructed at trainer.py:356)."""

    def __init__(self, sigma=0.1, min_weight=1.0, max_weight=3.0, blur_kernel=5):
        super().__init__()
        if blur_kernel != 7:
            raise NotImplementedError("the fused kernel implements blur_kernel=7 (reference modules/trainer.py:356)")
        self.sigma, self.min_weight, self.max_weight, self.blur_kernel = sigma, min_weight, max_weight, blur_kernel

    def forward(self, pred, target, source):
        return _Attention.apply(pred, target, source, self.sigma, self.min_weight, self.max_weight)


class _Region(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, t, s, threshold, weight):
        p, t, s = _prep(p, t, s)
        B, H, W = _img_dims(p)
        with torch.cuda.device(p.device):
            out = torch.empty((), dtype=torch.float32, device=p.device)
            state = torch.empty(8, dtype=torch.float32, device=p.device)
            call("ducosy_loss_contrast_region_forward", ptr(p), ptr(t), ptr(s), B, H, W, float(threshold), float(weight), ptr(out),
                 ptr(state), ptr(_scratch(p.device)), stream_ptr())
        ctx.save_for_backward(p, t, s, state)
        ctx.cfg = (float(threshold), float(weight))
        return out

    @staticmethod
    def backward(ctx, g):
        p, t, s, state = ctx.saved_tensors
        B, H, W = _img_dims(p)
        g = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(p.device):
            dp = torch.empty_like(p)
            call("ducosy_loss_contrast_region_backward", ptr(p), ptr(t), ptr(s), B, H, W, ctx.cfg[0], ctx.cfg[1], ptr(state), ptr(g),
                 ptr(dp), stream_ptr())
        return dp, None, None, None, None


class ContrastRegionLoss(nn.Module):
    """reference modules/trainer.py:89-130."""

    def __init__(self, threshold=0.3, weight=2.0):
        super().__init__()
        self.threshold, self.weight = threshold, weight

    def forward(self, pred, target, source):
        return _Region.apply(pred, target, source, self.threshold, self.weight)


class _Edge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, t):
        p, t = _prep(p, t)
        B, H, W = _img_dims(p)
        with torch.cuda.device(p.device):
            out = torch.empty((), dtype=torch.float32, device=p.device)
            ep, et = torch.empty_like(p), torch.empty_like(p)
            state = torch.empty(8, dtype=torch.float32, device=p.device)
            call("ducosy_loss_contrast_edge_forward", ptr(p), ptr(t), B, H, W, ptr(out), ptr(ep), ptr(et), ptr(state),
                 ptr(_scratch(p.device)), stream_ptr())
        ctx.save_for_backward(p, ep, state)
        return out

    @staticmethod
    def backward(ctx, g):
        p, ep, state = ctx.saved_tensors
        B, H, W = _img_dims(p)
        g = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(p.device):
            dp = torch.empty_like(p)
            call("ducosy_loss_contrast_edge_backward", ptr(p), ptr(ep), B, H, W, ptr(state), ptr(g), ptr(dp), stream_ptr())
        return dp, None


class ContrastEdgeLoss(nn.Module):
    """reference modules/trainer.py:133-184 (the top-10 % mean uses an exact radix selection instead of torch.topk)."""

    def forward(self, pred, target, source=None):
        return _Edge.apply(pred, target)


class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, data_range):
        x, y = _prep(x, y)
        B, H, W = _img_dims(x)
        with torch.cuda.device(x.device):
            out = torch.empty((), dtype=torch.float32, device=x.device)
            tmp = torch.empty(5 * B * H * (W - 10), dtype=torch.float32, device=x.device)
            dmaps = torch.empty(3 * B * (H - 10) * (W - 10), dtype=torch.float32, device=x.device)
            call("ducosy_loss_ssim_forward", ptr(x), ptr(y), B, H, W, float(data_range), ptr(out), ptr(tmp), ptr(dmaps),
                 ptr(_scratch(x.device)), stream_ptr())
        ctx.save_for_backward(x, y, dmaps)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, dmaps = ctx.saved_tensors
        B, H, W = _img_dims(x)
        g = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(x.device):
            tmp = torch.empty(3 * B * H * (W - 10), dtype=torch.float32, device=x.device)
            dx = torch.empty_like(x)
            call("ducosy_loss_ssim_backward", ptr(x), ptr(y), ptr(dmaps), B, H, W, ptr(g), ptr(tmp), ptr(dx), stream_ptr())
        return dx, None, None


class SSIM(nn.Module):
    """``pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=1)`` as constructed at reference
    modules/trainer.py:351 and called at :485 (``1 - ssim(rec, real)``).  PARITY UNPINNED: the package is not part of the
    reference tree; this follows its published algorithm (gaussian 11 / 1.5, valid convolution, K = (0.01, 0.03))."""

    def __init__(self, data_range=1.0, size_average=True, channel=1, **unused):
        super().__init__()
        if not size_average or channel != 1:
            raise NotImplementedError("only size_average=True, channel=1 (reference modules/trainer.py:351)")
        self.data_range = data_range

    def forward(self, x, y):
        return _SSIM.apply(x, y, self.data_range)
