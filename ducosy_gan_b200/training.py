"""Generator forward-with-saved-activations and backward for the CycleGAN step (reference modules/trainer.py:455-500
through modules/model.py:56-115).  Host orchestration only: every tensor op below is a kernel of libducosy_sm100.so
reached through the C ABI (ops.py); nothing here computes with torch.

Layout: activations are NHWC 16-bit, raw conv outputs are kept next to the InstanceNorm (scale, shift) pairs so the
normalised values are re-derived instead of stored; 16-bit gradient maps carry the power-of-two scale gs[0]
(ops.grad_scale), fp32 parameter / image gradients leave with the true scale.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .ops import ACT_NONE, ACT_RELU, PAD_REFLECT, PAD_ZERO


_WGRAD_STREAMS = {}


class _SideWgrad:
    """Weight gradients on a side stream.  In the backward of a convolution the weight gradient (dy, x -> dW) has no consumer
    before the optimiser, while the input gradient (dy, W -> dx) is on the critical path of the whole chain; at small per-rank
    batches neither fills the GPU (128-256 tiles on 148 SMs, and the InstanceNorm / padding kernels between the convolutions
    use a fraction of it), so the weight gradients of a generator pass CAN run on a second stream next to the chain, joined at
    the end of the backward.  Inputs are kept alive until the join (no allocator reuse while the side stream may still read
    them); captured in a CUDA graph this becomes a fork / join like the two generator streams.
    OPT-IN (DUCOSY_WGRAD_STREAM=1): measured neutral -- 17.84 vs 17.80 ms per step at batch 1, 101.4 vs 101.0 ms at batch 8
    (tools/gpu_ab_wgrad_stream.sh, profiles/r02_wgrad_stream_ab.json): the two generator streams already fill what the chain
    leaves idle."""

    def __init__(self):
        self.cur = torch.cuda.current_stream()
        self.side = None
        if os.environ.get("DUCOSY_WGRAD_STREAM", "0") == "1":
            key = (self.cur.device.index, self.cur.cuda_stream)
            if key not in _WGRAD_STREAMS:
                _WGRAD_STREAMS[key] = torch.cuda.Stream(device=self.cur.device)
            self.side = _WGRAD_STREAMS[key]
        self.keep = []

    def __call__(self, fn, *inputs):
        if self.side is None:
            return fn()
        self.keep.extend(inputs)
        self.side.wait_stream(self.cur)          # dy was just produced on the main stream
        with torch.cuda.stream(self.side):
            return fn()

    def join(self):
        if self.side is not None:
            self.cur.wait_stream(self.side)
        self.keep.clear()


def _split_params(params, num_blocks, use_cbam):
    """named_parameters() order of modules/model.py:90-115 -> (stem, d1, d2, blocks, u1, u2, out)."""
    it = iter(params)
    take = lambda n: [next(it) for _ in range(n)]
    stem, d1, d2 = take(2), take(2), take(2)
    blocks = [take(7 if use_cbam else 4) for _ in range(num_blocks)]
    u1, u2, out = take(2), take(2), take(2)
    return stem, d1, d2, blocks, u1, u2, out


def generator_forward_train(params, cfg, x, dtype):
    """x fp32 [B,Cin,H,W] -> (out fp32 [B,1,H,W], saved activations)."""
    _, num_blocks, use_cbam = cfg
    stem, d1, d2, blocks, u1, u2, outp = _split_params(params, num_blocks, use_cbam)
    B, _, H, W = x.shape
    if H % 32 or (W % 512 and W != 256):
        raise RuntimeError(f"training path needs H % 32 == 0 and W = 256 or a multiple of 512 (got {H}x{W}); 512x512 is what train.py uses")
    S = {"shape": (B, H, W)}
    # ops.conv2d_nhwc_in = conv + InstanceNorm finalize (one launch with DUCOSY_FUSED_FINALIZE=1, two by default)
    S["cols"] = ops.stem_im2col(x, dtype)
    S["y0"], S["n0"] = ops.conv2d_nhwc_in(S["cols"], ops.pack_stem_weight(stem[0], dtype), 1, 1, 1)
    S["p0"] = ops.in_apply_pad(S["y0"], *S["n0"], 1, PAD_ZERO, ACT_RELU)
    S["y1"], S["n1"] = ops.conv2d_nhwc_in(S["p0"], ops.pack_conv_weight(d1[0], dtype), 3, 3, 2)
    S["p1"] = ops.in_apply_pad(S["y1"], *S["n1"], 1, PAD_ZERO, ACT_RELU)
    S["y2"], S["n2"] = ops.conv2d_nhwc_in(S["p1"], ops.pack_conv_weight(d2[0], dtype), 3, 3, 2)
    r = ops.in_apply_pad(S["y2"], *S["n2"], 1, PAD_REFLECT if num_blocks else PAD_ZERO, ACT_RELU)
    S["blocks"] = []
    for i, bp in enumerate(blocks):
        last = i == num_blocks - 1
        sv = {"r": r}
        sv["ya"], sv["na"] = ops.conv2d_nhwc_in(r, ops.pack_conv_weight(bp[0], dtype), 3, 3, 1)
        sv["pa"] = ops.in_apply_pad(sv["ya"], *sv["na"], 1, PAD_REFLECT, ACT_RELU)
        out_mode = PAD_ZERO if last else PAD_REFLECT
        if use_cbam:
            sv["yb"], stats = ops.conv2d_nhwc_in(sv["pa"], ops.pack_conv_weight(bp[2], dtype), 3, 3, 1, want_chmax=True)
            r = ops.cbam_forward_train(sv, stats, bp[4], bp[5], bp[6], out_mode)
        else:
            sv["yb"], sv["nb"] = ops.conv2d_nhwc_in(sv["pa"], ops.pack_conv_weight(bp[2], dtype), 3, 3, 1)
            r = ops.residual_apply_pad(sv["yb"], *sv["nb"], None, sv["r"], 1, 1, out_mode)
        S["blocks"].append(sv)
    S["r_last"] = r
    S["yu1"], S["nu1"] = ops.upconv2x_nhwc_in(r, ops.pack_upconv_weight(u1[0], dtype), merged=False)
    S["pu1"] = ops.in_apply_pad(S["yu1"], *S["nu1"], 1, PAD_ZERO, ACT_RELU)
    S["yu2"], S["nu2"] = ops.upconv2x_nhwc_in(S["pu1"], ops.pack_upconv_merged_weight(u2[0], dtype), merged=True)
    S["pout"] = ops.in_apply_pad(S["yu2"], *S["nu2"], 3, PAD_REFLECT, ACT_RELU)
    out = ops.out_conv7x7_tanh(S["pout"], ops.pack_out_weight(outp[0], dtype), outp[1])
    S["out"] = out
    return out, S


def generator_backward(params, cfg, S, dout, want_dx=True, grad_out=None):
    """dout fp32 [B,1,H,W] -> (list of fp32 parameter gradients in named_parameters() order, dx fp32 [B,1,H,W] | None).
    Biases in front of an InstanceNorm receive an exact zero (the norm removes any per-channel constant), returned as None.
    ``grad_out``: per parameter an existing gradient tensor (``p.grad``) or None; the 3x3 convolution weights of the residual
    blocks and the down convolutions and the CBAM parameters accumulate straight into theirs (``+=``, what autograd's AccumulateGrad would do with a
    returned tensor) and come back as None."""
    _, num_blocks, use_cbam = cfg
    stem, d1, d2, blocks, u1, u2, outp = _split_params(params, num_blocks, use_cbam)
    if grad_out is None:
        grad_out = [None] * len(params)
    _, g_d1, g_d2, g_blocks, _, _, _ = _split_params(grad_out, num_blocks, use_cbam)
    dout = dout.to(torch.float32).contiguous()
    gs = ops.grad_scale(dout)
    zero = lambda p: None      # dead bias (in front of a non-affine InstanceNorm): exact zero gradient, materialised by the caller if needed

    side = _SideWgrad()
    da, dw_out, db_out = ops.out_conv_backward(dout, S["out"], S["pout"], outp[0], gs)
    dyu2 = ops.in_backward_pad(da, S["yu2"], *S["nu2"], 2, ACT_RELU)
    dpu1, dw_u2 = ops.upconv2x_backward(S["pu1"], dyu2, u2[0], gs)
    dyu1 = ops.in_backward_pad(dpu1, S["yu1"], *S["nu1"], 2, ACT_RELU)
    dr, dw_u1 = ops.upconv2x_backward(S["r_last"], dyu1, u1[0], gs)

    block_grads = []
    for bp, sv, gb in zip(reversed(blocks), reversed(S["blocks"]), reversed(g_blocks)):
        if use_cbam:
            dn, cbam_grads = ops.cbam_backward(sv, dr, bp[4], bp[5], bp[6], gs, accumulate_into=gb[4:7])
            dyb = ops.in_backward_pad(dn, sv["yb"], *sv["nb"], 2, ACT_NONE)
        else:
            cbam_grads = []
            dyb = ops.in_backward_pad(dr, sv["yb"], *sv["nb"], 2, ACT_NONE)
        C = bp[0].shape[0]
        dw_b = side(lambda: ops.conv2d_wgrad_oihw(sv["pa"], dyb, 3, 3, 1, dy_pad=2, gs=gs, accumulate_into=gb[2]), sv["pa"], dyb, gs)
        Wb = sv["ya"].shape[2]
        if Wb >= 8 and Wb & (Wb - 1) == 0:
            # the reflection adjoint of conv b's input padding is folded into the loads of the InstanceNorm backward
            _, dpa_pad = ops.conv3x3s1_dgrad(dyb, bp[2], PAD_REFLECT, fold=False)
            dya = ops.in_backward_pad_folded(dpa_pad, PAD_REFLECT, sv["ya"], *sv["na"], 2, ACT_RELU)
        else:
            dpa, _ = ops.conv3x3s1_dgrad(dyb, bp[2], PAD_REFLECT)
            dya = ops.in_backward_pad(dpa, sv["ya"], *sv["na"], 2, ACT_RELU)
        dw_a = side(lambda: ops.conv2d_wgrad_oihw(sv["r"], dya, 3, 3, 1, dy_pad=2, gs=gs, accumulate_into=gb[0]), sv["r"], dya, gs)
        dr, _ = ops.conv3x3s1_dgrad(dya, bp[0], PAD_REFLECT, add=dr)     # conv path + skip connection
        block_grads.append([dw_a, zero(bp[1]), dw_b, zero(bp[3])] + cbam_grads)
    block_grads.reverse()

    dy2 = ops.in_backward_pad(dr, S["y2"], *S["n2"], 1, ACT_RELU)
    dw_d2 = side(lambda: ops.conv2d_wgrad_oihw(S["p1"], dy2, 3, 3, 2, dy_pad=1, gs=gs, accumulate_into=g_d2[0]), S["p1"], dy2, gs)
    dp1 = ops.convs2_dgrad_nhwc(dy2, d2[0])
    dy1 = ops.in_backward_pad(dp1, S["y1"], *S["n1"], 1, ACT_RELU)
    dw_d1 = side(lambda: ops.conv2d_wgrad_oihw(S["p0"], dy1, 3, 3, 2, dy_pad=1, gs=gs, accumulate_into=g_d1[0]), S["p0"], dy1, gs)
    dp0 = ops.convs2_dgrad_nhwc(dy1, d1[0])
    dy0 = ops.in_backward_pad(dp0, S["y0"], *S["n0"], 0, ACT_RELU)
    dw_stem, dx = ops.stem_backward(dy0, S["cols"], stem[0], gs, want_dx)
    side.join()

    grads = [dw_stem, zero(stem[1]), dw_d1, zero(d1[1]), dw_d2, zero(d2[1])]
    for g in block_grads:
        grads += g
    grads += [dw_u1, zero(u1[1]), dw_u2, zero(u2[1]), dw_out, db_out]
    return grads, dx
