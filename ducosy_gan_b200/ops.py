"""Tensor-level wrappers over the layer entry points of libducosy_sm100.so.

Activations are NHWC 16-bit torch tensors (fp16 or bf16), statistics fp32.  These wrappers only allocate
outputs and pass raw pointers; every FLOP happens in the CUDA library.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ACT_LRELU02, ACT_NONE, ACT_RELU, PAD_REFLECT, PAD_ZERO, call, dtype_code, ptr, stream_ptr  # noqa: F401


def _dev(t):
    return torch.cuda.device(t.device)


# Packed 16-bit operand forms of a weight (forward K-major, the three dgrad packings, ...) are functions of the fp32 master
# weight alone.  One optimisation step uses every generator weight in several passes (translation + identity batch, cycle
# pass, and the backward of each), so inside ``with pack_cache(d):`` the packings are kept in ``d`` -- a dict OWNED BY THE
# MODULE whose parameters they belong to, so it dies with them and a recycled address can never alias a dead tensor -- and
# rebuilt when the tensor's ``_version`` moved (optimiser step, load_state_dict, any in-place op).  Inside a CUDA-graph
# capture the first use of a step records the pack kernel and the later uses record nothing, which is exactly the order a
# replay needs.  Outside such a context every call packs afresh.
import contextlib
import threading

_tls = threading.local()


@contextlib.contextmanager
def pack_cache(cache):
    prev = getattr(_tls, "cache", None)
    _tls.cache = cache
    try:
        yield
    finally:
        _tls.cache = prev


def _cached_pack(kind, w, dtype, build):
    cache = getattr(_tls, "cache", None)
    if cache is None:
        return build()
    key = (kind, w.data_ptr(), dtype)
    ent = cache.get(key)
    if ent is not None and ent[0] == w._version and ent[2] == tuple(w.shape):
        return ent[1]
    out = build()
    cache[key] = (w._version, out, tuple(w.shape))
    return out


# ------------------------------------------------------------------ HU kernels
def hu_window(px: torch.Tensor, slope, intercept, soft=(-150.0, 250.0), lung=(-1000.0, -150.0)):
    """int16 stored values -> (soft-window, lung-window) fp32 in [-1,1]   (reference preprocess.py:72-84)."""
    assert px.dtype == torch.int16 and px.is_cuda and px.is_contiguous()
    with _dev(px):
        o_s = torch.empty(px.shape, dtype=torch.float32, device=px.device)
        o_l = torch.empty(px.shape, dtype=torch.float32, device=px.device)
        call("ducosy_hu_window", ptr(px), ptr(o_s), ptr(o_l), px.numel(), float(slope), float(intercept),
             float(soft[0]), float(soft[1]), float(lung[0]), float(lung[1]), stream_ptr())
    return o_s, o_l


def hu_thresholds(px: torch.Tensor, slope, intercept):
    """(body, lung, bone) uint8 candidates   (reference mask_generator.py:14-20,179-183)."""
    assert px.dtype == torch.int16 and px.is_cuda and px.is_contiguous()
    with _dev(px):
        outs = [torch.empty(px.shape, dtype=torch.uint8, device=px.device) for _ in range(3)]
        call("ducosy_hu_thresholds", ptr(px), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), px.numel(), float(slope),
             float(intercept), stream_ptr())
    return tuple(outs)


def dewindow_composite(raw_px, y_soft, y_lung, slope, intercept, soft=(-150.0, 250.0), lung=(-1000.0, -150.0),
                       want_parts=False, out=None):
    """De-window both generator outputs and merge by HU range of the raw NCCT (reference preprocess.py:96-111,
    generate.py:218-237).  Returns merged int16 (and soft_px, lung_px, masks when want_parts)."""
    assert raw_px.dtype == torch.int16 and raw_px.is_contiguous()
    assert y_soft.dtype == torch.float32 and y_lung.dtype == torch.float32
    assert y_soft.numel() == raw_px.numel() == y_lung.numel()
    y_soft, y_lung = y_soft.contiguous(), y_lung.contiguous()
    with _dev(raw_px):
        merged = torch.empty_like(raw_px) if out is None else out
        sp = torch.empty_like(raw_px) if want_parts else None
        lp = torch.empty_like(raw_px) if want_parts else None
        mk = torch.empty(raw_px.shape, dtype=torch.uint8, device=raw_px.device) if want_parts else None
        call("ducosy_dewindow_composite", ptr(raw_px), ptr(y_soft), ptr(y_lung), ptr(merged), ptr(sp), ptr(lp), ptr(mk),
             raw_px.numel(), float(slope), float(intercept), float(soft[0]), float(soft[1]), float(lung[0]),
             float(lung[1]), stream_ptr())
    return (merged, sp, lp, mk) if want_parts else merged


# ------------------------------------------------------------------ weight packing
def _pack_conv_weight_build(w, dtype):
    Cout, Cin, kh, kw = w.shape
    w = w.detach().to(torch.float32).contiguous()
    with _dev(w):
        out = torch.empty((Cout, kh * kw * Cin), dtype=dtype, device=w.device)
        call("ducosy_pack_conv_weight", ptr(w), ptr(out), Cout, Cin, kh, kw, dtype_code(dtype), stream_ptr())
    return out


def pack_conv_weight(w: torch.Tensor, dtype=torch.float16):
    return _cached_pack("conv", w, dtype, lambda: _pack_conv_weight_build(w, dtype))


def _pack_upconv_weight_build(w, dtype):
    Cout, Cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    w = w.detach().to(torch.float32).contiguous()
    with _dev(w):
        out = torch.empty((4 * Cout, 4 * Cin), dtype=dtype, device=w.device)
        call("ducosy_pack_upconv_weight", ptr(w), ptr(out), Cout, Cin, dtype_code(dtype), stream_ptr())
    return out


def pack_upconv_weight(w: torch.Tensor, dtype=torch.float16):
    return _cached_pack("upconv", w, dtype, lambda: _pack_upconv_weight_build(w, dtype))


def _pack_stem_weight_build(w, dtype):
    Cout, Cin, kh, kw = w.shape
    assert Cout == 64 and kh == 7 and kw == 7
    w = w.detach().to(torch.float32).contiguous()
    kpad = (49 * Cin + 63) // 64 * 64
    with _dev(w):
        out = torch.empty((64, kpad), dtype=dtype, device=w.device)
        call("ducosy_pack_stem_weight", ptr(w), ptr(out), Cin, dtype_code(dtype), stream_ptr())
    return out


def pack_stem_weight(w: torch.Tensor, dtype=torch.float16):
    return _cached_pack("stem", w, dtype, lambda: _pack_stem_weight_build(w, dtype))


def _pack_out_weight_build(w, dtype):
    assert tuple(w.shape) == (1, 64, 7, 7)
    w = w.detach().to(torch.float32).contiguous()
    with _dev(w):
        out = torch.empty((7, 8, 64), dtype=dtype, device=w.device)
        call("ducosy_pack_out_weight", ptr(w), ptr(out), dtype_code(dtype), stream_ptr())
    return out


def pack_out_weight(w: torch.Tensor, dtype=torch.float16):
    return _cached_pack("out", w, dtype, lambda: _pack_out_weight_build(w, dtype))


# ------------------------------------------------------------------ convolutions
def conv2d_nhwc(x_pad: torch.Tensor, w_packed: torch.Tensor, kh, kw, stride, want_stats=True, bias=None, act=ACT_NONE):
    """x_pad [B,Hp,Wp,Cin] 16-bit (already padded) -> raw y [B,Ho,Wo,Cout], partials [B,Ho*Wo/128,3,Cout] | None."""
    B, Hp, Wp, Cin = x_pad.shape
    Cout = w_packed.shape[0]
    assert x_pad.is_contiguous() and w_packed.is_contiguous() and w_packed.dtype == x_pad.dtype
    Ho, Wo = (Hp - kh) // stride + 1, (Wp - kw) // stride + 1
    with _dev(x_pad):
        y = torch.empty((B, Ho, Wo, Cout), dtype=x_pad.dtype, device=x_pad.device)
        partials = torch.empty((B, Ho * Wo // 128, 3, Cout), dtype=torch.float32, device=x_pad.device) if want_stats else None
        call("ducosy_conv2d_nhwc", ptr(x_pad), ptr(w_packed), ptr(y), ptr(partials), ptr(bias), int(act), B, Hp, Wp, Cin,
             Cout, kh, kw, stride, dtype_code(x_pad.dtype), stream_ptr())
    return y, partials


_TICKETS = {}


def _tickets(device, B):
    """Zeroed int32 ticket array of the fused conv + InstanceNorm-finalize launches: one per (device, stream) -- launches on
    different streams may overlap -- kept zero by the kernels themselves."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _TICKETS.get(key)
    if t is None or t.numel() < B:
        t = _TICKETS[key] = torch.zeros(max(1024, B), dtype=torch.int32, device=device)
    return t


def _fused_finalize():
    """DUCOSY_FUSED_FINALIZE=1: finalize the InstanceNorm statistics inside the conv launch (ducosy_*_in entry points).  Measured
    slower than the separate finalize kernel at every batch size (profiles/r02_fusion_ab.json), hence off by default."""
    import os
    return os.environ.get("DUCOSY_FUSED_FINALIZE", "0") == "1"


def conv2d_nhwc_in(x_pad: torch.Tensor, w_packed: torch.Tensor, kh, kw, stride, want_chmax=False):
    """conv2d_nhwc + in_finalize: raw y [B,Ho,Wo,Cout] and the InstanceNorm pair (scale, shift) [B,Cout] (+ the normalised
    per-channel max when want_chmax).  One launch with DUCOSY_FUSED_FINALIZE=1, conv + finalize kernel otherwise."""
    B, Hp, Wp, Cin = x_pad.shape
    Cout = w_packed.shape[0]
    Ho, Wo = (Hp - kh) // stride + 1, (Wp - kw) // stride + 1
    if not _fused_finalize():
        y, partials = conv2d_nhwc(x_pad, w_packed, kh, kw, stride)
        return y, in_finalize(partials, Ho * Wo, want_chmax=want_chmax)
    assert x_pad.is_contiguous() and w_packed.is_contiguous() and w_packed.dtype == x_pad.dtype
    with _dev(x_pad):
        dev = x_pad.device
        y = torch.empty((B, Ho, Wo, Cout), dtype=x_pad.dtype, device=dev)
        partials = torch.empty((B, Ho * Wo // 128, 3, Cout), dtype=torch.float32, device=dev)
        scale = torch.empty((B, Cout), dtype=torch.float32, device=dev)
        shift = torch.empty((B, Cout), dtype=torch.float32, device=dev)
        chmax = torch.empty((B, Cout), dtype=torch.float32, device=dev) if want_chmax else None
        call("ducosy_conv2d_nhwc_in", ptr(x_pad), ptr(w_packed), ptr(y), ptr(partials), ptr(scale), ptr(shift), ptr(chmax),
             ptr(_tickets(dev, B)), B, Hp, Wp, Cin, Cout, kh, kw, stride, dtype_code(x_pad.dtype), stream_ptr())
    return (y, (scale, shift, chmax)) if want_chmax else (y, (scale, shift))


def upconv2x_nhwc_in(x_pad: torch.Tensor, w_packed, merged: bool):
    """upconv2x_nhwc / upconv2x_merged_nhwc + in_finalize: raw y [B,2Hs,2Ws,Cout], (scale, shift)."""
    B, Hp, Wp, Cin = x_pad.shape
    Hs, Ws = Hp - 2, Wp - 2
    Cout = w_packed.shape[0] // 4
    if not _fused_finalize():
        y, partials = (upconv2x_merged_nhwc if merged else upconv2x_nhwc)(x_pad, w_packed)
        return y, in_finalize(partials, 4 * Hs * Ws)
    with _dev(x_pad):
        dev = x_pad.device
        y = torch.empty((B, 2 * Hs, 2 * Ws, Cout), dtype=x_pad.dtype, device=dev)
        tiles = (Hs * Ws // 128) if merged else (4 * Hs * Ws // 128)
        partials = torch.empty((B, tiles, 3, Cout), dtype=torch.float32, device=dev)
        scale = torch.empty((B, Cout), dtype=torch.float32, device=dev)
        shift = torch.empty((B, Cout), dtype=torch.float32, device=dev)
        call("ducosy_upconv2x_merged_nhwc_in" if merged else "ducosy_upconv2x_nhwc_in", ptr(x_pad), ptr(w_packed), ptr(y), ptr(partials),
             ptr(scale), ptr(shift), ptr(_tickets(dev, B)), B, Hs, Ws, Cin, Cout, dtype_code(x_pad.dtype), stream_ptr())
    return y, (scale, shift)


def upconv2x_nhwc(x_pad: torch.Tensor, w_packed4: torch.Tensor):
    """Upsample(x2 nearest)+Conv3x3(pad 1) from the zero-padded source [B,Hs+2,Ws+2,Cin] -> raw y [B,2Hs,2Ws,Cout]."""
    B, Hp, Wp, Cin = x_pad.shape
    Hs, Ws = Hp - 2, Wp - 2
    Cout = w_packed4.shape[0] // 4
    with _dev(x_pad):
        y = torch.empty((B, 2 * Hs, 2 * Ws, Cout), dtype=x_pad.dtype, device=x_pad.device)
        partials = torch.empty((B, 4 * Hs * Ws // 128, 3, Cout), dtype=torch.float32, device=x_pad.device)
        call("ducosy_upconv2x_nhwc", ptr(x_pad), ptr(w_packed4), ptr(y), ptr(partials), B, Hs, Ws, Cin, Cout,
             dtype_code(x_pad.dtype), stream_ptr())
    return y, partials


def _pack_upconv_merged_weight_build(w, dtype):
    Cout, Cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    w = w.detach().to(torch.float32).contiguous()
    with _dev(w):
        out = torch.empty((4 * Cout, 9 * Cin), dtype=dtype, device=w.device)
        call("ducosy_pack_upconv_merged_weight", ptr(w), ptr(out), Cout, Cin, dtype_code(dtype), stream_ptr())
    return out


def pack_upconv_merged_weight(w: torch.Tensor, dtype=torch.float16):
    return _cached_pack("upconv_merged", w, dtype, lambda: _pack_upconv_merged_weight_build(w, dtype))


def upconv2x_merged_nhwc(x_pad: torch.Tensor, w_merged: torch.Tensor):
    """Merged-phase Upsample(x2)+Conv3x3 (Cout = 64): raw y [B,2Hs,2Ws,Cout], partials [B,Hs*Ws/128,3,Cout]."""
    B, Hp, Wp, Cin = x_pad.shape
    Hs, Ws = Hp - 2, Wp - 2
    Cout = w_merged.shape[0] // 4
    with _dev(x_pad):
        y = torch.empty((B, 2 * Hs, 2 * Ws, Cout), dtype=x_pad.dtype, device=x_pad.device)
        partials = torch.empty((B, Hs * Ws // 128, 3, Cout), dtype=torch.float32, device=x_pad.device)
        call("ducosy_upconv2x_merged_nhwc", ptr(x_pad), ptr(w_merged), ptr(y), ptr(partials), B, Hs, Ws, Cin, Cout,
             dtype_code(x_pad.dtype), stream_ptr())
    return y, partials


def stem_im2col(x: torch.Tensor, dtype=torch.float16):
    B, Cin, H, W = x.shape
    x = x.to(torch.float32).contiguous()
    kpad = (49 * Cin + 63) // 64 * 64
    with _dev(x):
        a = torch.empty((B, H, W, kpad), dtype=dtype, device=x.device)
        call("ducosy_stem_im2col", ptr(x), ptr(a), B, Cin, H, W, dtype_code(dtype), stream_ptr())
    return a


def stem_im2col_hu(px: torch.Tensor, slope, intercept, lo, hi, dtype=torch.float16):
    B, H, W = px.shape
    with _dev(px):
        a = torch.empty((B, H, W, 64), dtype=dtype, device=px.device)
        call("ducosy_stem_im2col_hu", ptr(px), ptr(a), B, H, W, float(slope), float(intercept), float(lo), float(hi),
             dtype_code(dtype), stream_ptr())
    return a


# ------------------------------------------------------------------ normalisation / attention
def in_finalize(partials: torch.Tensor, npix: int, fc0=None, fc2=None, want_chmax=False):
    B, tiles, _, Cn = partials.shape
    with _dev(partials):
        scale = torch.empty((B, Cn), dtype=torch.float32, device=partials.device)
        shift = torch.empty((B, Cn), dtype=torch.float32, device=partials.device)
        chmax = torch.empty((B, Cn), dtype=torch.float32, device=partials.device) if (fc0 is not None or want_chmax) else None
        call("ducosy_in_finalize", ptr(partials), tiles, int(npix), ptr(scale), ptr(shift), ptr(fc0), ptr(fc2), ptr(chmax),
             B, Cn, stream_ptr())
    if want_chmax:
        return scale, shift, chmax
    return scale, shift


def in_apply_pad(y, scale, shift, pad, pad_mode, act):
    B, H, W, Cn = y.shape
    with _dev(y):
        out = torch.empty((B, H + 2 * pad, W + 2 * pad, Cn), dtype=y.dtype, device=y.device)
        call("ducosy_in_apply_pad", ptr(y), ptr(scale), ptr(shift), ptr(out), B, H, W, Cn, pad, pad_mode, act,
             dtype_code(y.dtype), stream_ptr())
    return out


def cbam_pool(y, scale, shift):
    B, H, W, Cn = y.shape
    with _dev(y):
        pooled = torch.empty((B, H, W, 2), dtype=torch.float32, device=y.device)
        call("ducosy_cbam_pool", ptr(y), ptr(scale), ptr(shift), ptr(pooled), B, H, W, Cn, dtype_code(y.dtype), stream_ptr())
    return pooled


def cbam_spatial_conv(pooled, w_sa):
    B, H, W, _ = pooled.shape
    w_sa = w_sa.detach().to(torch.float32).contiguous()
    with _dev(pooled):
        sa = torch.empty((B, H, W), dtype=torch.float32, device=pooled.device)
        call("ducosy_cbam_spatial_conv", ptr(pooled), ptr(w_sa), ptr(sa), B, H, W, stream_ptr())
    return sa


def residual_apply_pad(y, scale, shift, sa, res_pad, res_pad_width, pad, pad_mode):
    B, H, W, Cn = y.shape
    with _dev(y):
        out = torch.empty((B, H + 2 * pad, W + 2 * pad, Cn), dtype=y.dtype, device=y.device)
        call("ducosy_residual_apply_pad", ptr(y), ptr(scale), ptr(shift), ptr(sa), ptr(res_pad), res_pad_width, ptr(out),
             B, H, W, Cn, pad, pad_mode, dtype_code(y.dtype), stream_ptr())
    return out


def out_conv7x7_tanh(x_pad, w_packed, bias):
    B, Hp, Wp, Cn = x_pad.shape
    assert Cn == 64
    H, W = Hp - 6, Wp - 6
    bias = bias.detach().to(torch.float32).contiguous()
    with _dev(x_pad):
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x_pad.device)
        call("ducosy_out_conv7x7_tanh", ptr(x_pad), ptr(w_packed), ptr(bias), ptr(out), B, H, W,
             dtype_code(x_pad.dtype), stream_ptr())
    return out


def stem_fused(w_packed, x=None, px=None, window=None):
    """Fused stem for Cin = 1: (x fp32 [B,1,H,W] | px int16 [B,H,W] + window=(slope, intercept, lo, hi)) ->
    zero-padded ReLU(IN(conv7x7)) [B,H+2,W+2,64] 16-bit."""
    src = x if x is not None else px
    B, H, W = (x.shape[0], x.shape[2], x.shape[3]) if x is not None else px.shape
    slope, intercept, lo, hi = window if window is not None else (0.0, 0.0, 0.0, 0.0)
    dt = w_packed.dtype
    with _dev(src):
        xw = torch.empty((B, H, W), dtype=dt, device=src.device)
        partials = torch.empty((B, H, 3, 64), dtype=torch.float32, device=src.device)
        out = torch.empty((B, H + 2, W + 2, 64), dtype=dt, device=src.device)
        xin = x.to(torch.float32).contiguous() if x is not None else None
        call("ducosy_stem_prepare", ptr(xin), ptr(px) if px is not None else None, float(slope), float(intercept),
             float(lo), float(hi), ptr(xw), B, H, W, dtype_code(dt), stream_ptr())
        call("ducosy_stem_fused", ptr(xw), ptr(w_packed), ptr(partials), None, None, None, B, H, W, 0, dtype_code(dt), stream_ptr())
        scale, shift = in_finalize(partials, H * W)
        call("ducosy_stem_fused", ptr(xw), ptr(w_packed), None, ptr(scale), ptr(shift), ptr(out), B, H, W, 1, dtype_code(dt), stream_ptr())
    return out


def out_conv7x7_tanh_fused(y_raw, scale, shift, w_packed, bias):
    """IN apply + ReLU + reflect pad 3 + conv7x7(64->1) + tanh from the raw NHWC map [B,H,W,64]."""
    B, H, W, Cn = y_raw.shape
    assert Cn == 64
    bias = bias.detach().to(torch.float32).contiguous()
    with _dev(y_raw):
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=y_raw.device)
        call("ducosy_out_conv7x7_tanh_fused", ptr(y_raw), ptr(scale), ptr(shift), ptr(w_packed), ptr(bias), ptr(out), B, H, W,
             dtype_code(y_raw.dtype), stream_ptr())
    return out


def conv2d_wgrad_nhwc(x_pad, dy, kh, kw, stride, dy_pad=0):
    """Weight gradient dW fp32 [Cout, kh*kw*Cin] (packed forward layout) from the padded input and the output gradient
    (dy may be stored inside a buffer with a border of dy_pad pixels)."""
    B, Hp, Wp, Cin = x_pad.shape
    _, Ho, Wo, Cout = dy.shape
    Ho, Wo = Ho - 2 * dy_pad, Wo - 2 * dy_pad
    assert x_pad.is_contiguous() and dy.is_contiguous() and x_pad.dtype == dy.dtype
    lib = _lib.load()
    with _dev(x_pad):
        need = lib.ducosy_conv2d_wgrad_workspace_bytes(B, Ho, Wo, Cin, Cout, kh, kw)
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=x_pad.device)
        dw = torch.empty((Cout, kh * kw * Cin), dtype=torch.float32, device=x_pad.device)
        call("ducosy_conv2d_wgrad_nhwc", ptr(x_pad), ptr(dy), int(dy_pad), ptr(dw), B, Hp, Wp, Cin, Cout, kh, kw, stride, ptr(ws),
             ws.numel(), dtype_code(x_pad.dtype), stream_ptr())
    return dw


def conv2d_wgrad_oihw(x_pad, dy, kh, kw, stride, dy_pad=0, gs=None, accumulate_into=None):
    """Weight gradient in the parameter's own layout, fp32 [Cout,Cin,kh,kw], true scale (times gs[1]): conv2d_wgrad_nhwc and
    unpack_wgrad in one reduction pass.  ``accumulate_into``: an existing fp32 gradient of that shape (the parameter's ``.grad``)
    that receives ``+= dW`` instead of a new tensor being returned (returns None)."""
    B, Hp, Wp, Cin = x_pad.shape
    _, Ho, Wo, Cout = dy.shape
    Ho, Wo = Ho - 2 * dy_pad, Wo - 2 * dy_pad
    assert x_pad.is_contiguous() and dy.is_contiguous() and x_pad.dtype == dy.dtype
    lib = _lib.load()
    with _dev(x_pad):
        need = lib.ducosy_conv2d_wgrad_workspace_bytes(B, Ho, Wo, Cin, Cout, kh, kw)
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=x_pad.device)
        if accumulate_into is not None:
            assert accumulate_into.shape == (Cout, Cin, kh, kw) and accumulate_into.dtype == torch.float32 and accumulate_into.is_contiguous()
            call("ducosy_conv2d_wgrad_nhwc_oihw_acc", ptr(x_pad), ptr(dy), int(dy_pad), ptr(accumulate_into), ptr(gs), B, Hp, Wp, Cin,
                 Cout, kh, kw, stride, ptr(ws), ws.numel(), dtype_code(x_pad.dtype), stream_ptr())
            return None
        dw = torch.empty((Cout, Cin, kh, kw), dtype=torch.float32, device=x_pad.device)
        call("ducosy_conv2d_wgrad_nhwc_oihw", ptr(x_pad), ptr(dy), int(dy_pad), ptr(dw), ptr(gs), B, Hp, Wp, Cin, Cout, kh, kw, stride,
             ptr(ws), ws.numel(), dtype_code(x_pad.dtype), stream_ptr())
    return dw


def in_backward_pad(da, y, scale, shift, pad, act):
    """InstanceNorm(+activation) backward: da, y NHWC 16-bit -> dy zero-padded [B,H+2p,W+2p,C]."""
    B, H, W, Cn = y.shape
    lib = _lib.load()
    with _dev(y):
        scratch = torch.empty(max(lib.ducosy_in_backward_scratch_bytes(B, H, W, Cn) // 4, 4), dtype=torch.float32, device=y.device)
        out = torch.empty((B, H + 2 * pad, W + 2 * pad, Cn), dtype=y.dtype, device=y.device)
        call("ducosy_in_backward_pad", ptr(da), ptr(y), ptr(scale), ptr(shift), ptr(out), ptr(scratch), B, H, W, Cn, pad, act,
             dtype_code(y.dtype), stream_ptr())
    return out


def in_backward_pad_folded(da_pad1, fold_mode, y, scale, shift, pad, act):
    """``in_backward_pad`` of the map whose gradient w.r.t. its pad-1 form is ``da_pad1`` [B,H+2,W+2,C] (``conv3x3s1_dgrad(...,
    fold=False)``): the streaming kernels read the padded map's interior, the folded map is never written.  ``da_pad1`` is
    CONSUMED: for reflection padding the mirrored border terms are first added into its interior cells, in place."""
    B, H, W, Cn = y.shape
    assert da_pad1.shape == (B, H + 2, W + 2, Cn) and da_pad1.is_contiguous()
    lib = _lib.load()
    with _dev(y):
        scratch = torch.empty(max(lib.ducosy_in_backward_scratch_bytes(B, H, W, Cn) // 4, 4), dtype=torch.float32, device=y.device)
        out = torch.empty((B, H + 2 * pad, W + 2 * pad, Cn), dtype=y.dtype, device=y.device)
        call("ducosy_in_backward_pad_folded", ptr(da_pad1), int(fold_mode), ptr(y), ptr(scale), ptr(shift), ptr(out), ptr(scratch), B, H, W,
             Cn, pad, act, dtype_code(y.dtype), stream_ptr())
    return out


def convs2_dgrad_nhwc(dy_pad, w_oihw):
    """Input gradient of Conv2d(Cin,Cout,k,stride 2,padding 1), k in {3,4}: dy_pad [B,Ho+2,Wo+2,Cout] -> dx [B,2Ho,2Wo,Cin]."""
    Cout, Cin, ksz = w_oihw.shape[:3]
    B, Hp, Wp, _ = dy_pad.shape
    w = w_oihw.detach().to(torch.float32).contiguous()
    with _dev(dy_pad):
        def build():
            wd = torch.empty((4 * Cin, 4 * Cout), dtype=dy_pad.dtype, device=dy_pad.device)
            call("ducosy_pack_dgrad_s2_weight", ptr(w), ptr(wd), Cout, Cin, int(ksz), dtype_code(dy_pad.dtype), stream_ptr())
            return wd
        wd = _cached_pack("dgrad_s2", w_oihw, dy_pad.dtype, build)
        dx = torch.empty((B, 2 * (Hp - 2), 2 * (Wp - 2), Cin), dtype=dy_pad.dtype, device=dy_pad.device)
        call("ducosy_convs2_dgrad_nhwc", ptr(dy_pad), ptr(wd), ptr(dx), B, Hp - 2, Wp - 2, Cin, Cout, dtype_code(dy_pad.dtype),
             stream_ptr())
    return dx


def hu_window_soft(px, slope, intercept, hu_lo, hu_hi, sigma=50.0):
    """Soft-squeezed training windowing (reference preprocess.py:6-55): int16 stored values -> fp32 in [-1,1]."""
    assert px.dtype == torch.int16 and px.is_cuda and px.is_contiguous()
    with _dev(px):
        out = torch.empty(px.shape, dtype=torch.float32, device=px.device)
        call("ducosy_hu_window_soft", ptr(px), ptr(out), px.numel(), float(slope), float(intercept), float(hu_lo), float(hu_hi),
             float(sigma), stream_ptr())
    return out


def apply_windowing(y, hu_lo, hu_hi, window_center, window_width):
    """Display windowing of a tanh-range tensor (reference preprocess.py:58-65)."""
    y = y.to(torch.float32).contiguous()
    with _dev(y):
        out = torch.empty_like(y)
        call("ducosy_apply_windowing", ptr(y), ptr(out), y.numel(), float(hu_lo), float(hu_hi), float(window_center),
             float(window_width), stream_ptr())
    return out


def conv3x3s1_dgrad(dy_pad2, w_oihw, pad_mode, add=None, fold=True):
    """Input gradient of pad(1) + Conv2d(3x3): dy_pad2 [B,H+4,W+4,Cout] (zero border 2) -> dx [B,H,W,Cin] after folding
    the padding adjoint (reflect or zero); ``add`` [B,H,W,Cin] is summed in (the skip connection's gradient).
    ``fold=False`` returns (None, dxpad): the consumer folds while loading (``in_backward_pad_folded``)."""
    Cout, Cin = w_oihw.shape[:2]
    B, Hp, Wp, _ = dy_pad2.shape
    H, W = Hp - 4, Wp - 4
    w = w_oihw.detach().to(torch.float32).contiguous()
    dc = dtype_code(dy_pad2.dtype)
    with _dev(dy_pad2):
        def build():
            wd = torch.empty((Cin, 9 * Cout), dtype=dy_pad2.dtype, device=dy_pad2.device)
            call("ducosy_pack_dgrad_s1_weight", ptr(w), ptr(wd), Cout, Cin, dc, stream_ptr())
            return wd
        wd = _cached_pack("dgrad_s1", w_oihw, dy_pad2.dtype, build)
        dxpad = torch.empty((B, H + 2, W + 2, Cin), dtype=dy_pad2.dtype, device=dy_pad2.device)
        call("ducosy_conv3x3s1_dgrad_nhwc", ptr(dy_pad2), ptr(wd), ptr(dxpad), B, H, W, Cin, Cout, dc, stream_ptr())
        if not fold:
            assert add is None
            return None, dxpad
        dx = torch.empty((B, H, W, Cin), dtype=dy_pad2.dtype, device=dy_pad2.device)
        call("ducosy_pad_fold_add", ptr(dxpad), ptr(add), ptr(dx), B, H, W, Cin, 1, pad_mode, dc, stream_ptr())
    return dx, dxpad


def pad_fold(dxpad, pad, pad_mode, add=None):
    """Adjoint of ReflectionPad2d(pad) / zero padding: dxpad [B,H+2p,W+2p,C] -> dx [B,H,W,C] (+ ``add``)."""
    B, Hp, Wp, Cn = dxpad.shape
    H, W = Hp - 2 * pad, Wp - 2 * pad
    with _dev(dxpad):
        dx = torch.empty((B, H, W, Cn), dtype=dxpad.dtype, device=dxpad.device)
        call("ducosy_pad_fold_add", ptr(dxpad), ptr(add), ptr(dx), B, H, W, Cn, pad, pad_mode, dtype_code(dxpad.dtype), stream_ptr())
    return dx


def unpack_wgrad(dwp, Cout, Cin, taps, gs=None):
    """Packed fp32 weight gradient [Cout, taps*Cin] -> OIHW [Cout, Cin, k, k] (times gs[1] when a gradient scale is given)."""
    k = int(round(taps ** 0.5))
    with _dev(dwp):
        dw = torch.empty((Cout, Cin, k, k), dtype=torch.float32, device=dwp.device)
        call("ducosy_unpack_wgrad", ptr(dwp), ptr(dw), Cout, Cin, taps, ptr(gs), stream_ptr())
    return dw


def add_inplace(a, b):
    """a += b on 16-bit maps."""
    assert a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()
    with _dev(a):
        call("ducosy_add_inplace", ptr(a), ptr(b), a.numel(), dtype_code(a.dtype), stream_ptr())
    return a


def upconv2x_backward(src_pad, dy_pad2, w_oihw, gs=None):
    """Backward of Upsample(x2)+Conv3x3(pad 1): src_pad [B,Hs+2,Ws+2,Cin] (zero border), dy_pad2 [B,2Hs+4,2Ws+4,Cout]
    (zero border 2) -> (dsrc [B,Hs,Ws,Cin], dW fp32 OIHW [Cout,Cin,3,3])."""
    Cout, Cin = w_oihw.shape[:2]
    B, Hp, Wp, _ = src_pad.shape
    Hs, Ws = Hp - 2, Wp - 2
    dt, dc = src_pad.dtype, dtype_code(src_pad.dtype)
    w = w_oihw.detach().to(torch.float32).contiguous()
    with _dev(src_pad):
        def build():
            wd = torch.empty((Cin, 16 * Cout), dtype=dt, device=src_pad.device)
            call("ducosy_pack_upconv_dgrad_weight", ptr(w), ptr(wd), Cout, Cin, dc, stream_ptr())
            return wd
        wd = _cached_pack("dgrad_upconv", w_oihw, dt, build)
        dsrc = torch.empty((B, Hs, Ws, Cin), dtype=dt, device=src_pad.device)
        call("ducosy_upconv2x_dgrad_nhwc", ptr(dy_pad2), ptr(wd), ptr(dsrc), B, Hs, Ws, Cin, Cout, dc, stream_ptr())
        lib = _lib.load()
        ws = torch.empty(max(lib.ducosy_upconv2x_wgrad_workspace_bytes(B, Hs, Ws, Cin, Cout), 16), dtype=torch.uint8, device=src_pad.device)
        dwph = torch.empty((Cout, 16 * Cin), dtype=torch.float32, device=src_pad.device)
        call("ducosy_upconv2x_wgrad_nhwc", ptr(src_pad), ptr(dy_pad2), 2, ptr(dwph), B, Hs, Ws, Cin, Cout, ptr(ws), ws.numel(), dc,
             stream_ptr())
        dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=src_pad.device)
        call("ducosy_unpack_upconv_wgrad", ptr(dwph), ptr(dw), Cout, Cin, ptr(gs), stream_ptr())
    return dsrc, dw


def grad_scale(g):
    """Power-of-two scale for the 16-bit gradient maps: gs[0] brings max|g| into [1,2), gs[1] = 1/gs[0]."""
    g = g.contiguous()
    with _dev(g):
        gs = torch.empty(2, dtype=torch.float32, device=g.device)
        call("ducosy_grad_scale", ptr(g), g.numel(), ptr(gs), stream_ptr())
    return gs


def out_conv_backward(dout, out, in_pad, w, gs):
    """Backward of ReflectionPad(3)+Conv7x7(64->1)+Tanh (reference modules/model.py:112-113): returns
    (da 16-bit [B,H,W,64] scaled by gs[0], dw fp32 [1,64,7,7], db fp32 [1])."""
    B, Hp, Wp, Cn = in_pad.shape
    H, W = Hp - 6, Wp - 6
    assert Cn == 64 and in_pad.is_contiguous()
    dout = dout.to(torch.float32).contiguous()
    out = out.to(torch.float32).contiguous()
    w = w.detach().to(torch.float32).contiguous()
    lib = _lib.load()
    with _dev(in_pad):
        scratch = torch.empty(lib.ducosy_out_conv_backward_scratch_bytes(B, H, W) // 4, dtype=torch.float32, device=in_pad.device)
        da = torch.empty((B, H, W, 64), dtype=in_pad.dtype, device=in_pad.device)
        dw = torch.empty((1, 64, 7, 7), dtype=torch.float32, device=in_pad.device)
        db = torch.empty(1, dtype=torch.float32, device=in_pad.device)
        call("ducosy_out_conv_backward", ptr(dout), ptr(out), ptr(in_pad), ptr(w), ptr(da), ptr(dw), ptr(db), ptr(scratch), ptr(gs),
             B, H, W, dtype_code(in_pad.dtype), stream_ptr())
    return da, dw, db


def stem_backward(dy, cols, w, gs, want_dx=True):
    """Backward of ReflectionPad(3)+Conv7x7(Cin->64) given the gradient dy [B,H,W,64] of the raw conv output and the
    saved im2col matrix cols [B,H,W,Kpad] (reference modules/model.py:90-91): returns (dw fp32 [64,Cin,7,7] true scale,
    dx fp32 [B,1,H,W] for image channel 0 -- the mask channels are data and get no gradient -- or None)."""
    B, H, W, Kpad = cols.shape
    Cin = w.shape[1]
    dt = dy.dtype
    with _dev(dy):
        dwp = conv2d_wgrad_nhwc(cols, dy, 1, 1, 1)                       # [64, Kpad]
        dw = torch.empty((64, Cin, 7, 7), dtype=torch.float32, device=dy.device)
        call("ducosy_unpack_stem_wgrad", ptr(dwp), ptr(dw), Cin, Kpad, ptr(gs), stream_ptr())
        dx = None
        if want_dx:
            # dcol[p][k] = sum_o dy[p][o] * w[o][0][k]  (k < 49): a 1x1 conv with "Cout" = 64 columns
            def build():
                w1 = torch.zeros((64, 64, 1, 1), dtype=torch.float32, device=dy.device)
                w1[:49, :, 0, 0] = w.detach().to(torch.float32)[:, 0].reshape(64, 49).t()
                return _pack_conv_weight_build(w1, dt)
            dcol, _ = conv2d_nhwc(dy, _cached_pack("stem_dgrad", w, dt, build), 1, 1, 1, want_stats=False)
            dx = torch.empty((B, 1, H, W), dtype=torch.float32, device=dy.device)
            call("ducosy_stem_col2im", ptr(dcol), ptr(dx), ptr(gs), B, H, W, dtype_code(dt), stream_ptr())
    return dw, dx


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def cbam_forward_train(sv, stats, fc0, fc2, wsa, out_mode):
    """Training-mode tail of a ResidualBlockWithCBAM (reference modules/model.py:68-87): from the raw second conv output
    sv["yb"] and its InstanceNorm statistics ``stats`` = (scale, shift, normalised channel max) of conv2d_nhwc_in to the padded
    block output, keeping what the backward needs in sv."""
    yb = sv["yb"]
    B, H, W, Cn = yb.shape
    dev = yb.device
    scale_n, shift_n, chmax = stats
    with _dev(yb):
        scale_v, shift_v, ca = (torch.empty((B, Cn), dtype=torch.float32, device=dev) for _ in range(3))
        hidden = torch.empty((B, Cn // 16), dtype=torch.float32, device=dev)
        call("ducosy_cbam_channel_train", ptr(chmax), ptr(_f32c(fc0)), ptr(_f32c(fc2)), ptr(scale_n), ptr(shift_n), ptr(scale_v),
             ptr(shift_v), ptr(ca), ptr(hidden), B, Cn, stream_ptr())
    pooled = cbam_pool(yb, scale_v, shift_v)
    sa = cbam_spatial_conv(pooled, wsa)
    sv.update(nb=(scale_n, shift_n), nv=(scale_v, shift_v), ca=ca, hidden=hidden, chmax=chmax, pooled=pooled, sa=sa)
    return residual_apply_pad(yb, scale_v, shift_v, sa, sv["r"], 1, 1, out_mode)


def cbam_backward(sv, dout, fc0, fc2, wsa, gs, accumulate_into=None):
    """CBAM backward: dout [B,H,W,C] 16-bit (gradient of the block output) -> (dn 16-bit gradient w.r.t. InstanceNorm(yb),
    [d fc.0.weight, d fc.2.weight, d spatial conv weight] fp32 true scale).  ``accumulate_into``: the three parameters' existing
    fp32 gradients (all three or none), which receive ``+=`` -- the list returned is then [None, None, None]."""
    yb = sv["yb"]
    B, H, W, Cn = yb.shape
    dev = yb.device
    lib = _lib.load()
    with _dev(yb):
        scratch = torch.empty(lib.ducosy_cbam_backward_scratch_bytes(B, H, W, Cn) // 4, dtype=torch.float32, device=dev)
        dn = torch.empty_like(yb)
        acc = accumulate_into is not None and all(g is not None for g in accumulate_into)
        if acc:
            dfc0, dfc2, dwsa = accumulate_into
            assert dfc0.shape == fc0.shape and dfc2.shape == fc2.shape and dwsa.shape == wsa.shape
        else:
            dfc0 = torch.empty(tuple(fc0.shape), dtype=torch.float32, device=dev)
            dfc2 = torch.empty(tuple(fc2.shape), dtype=torch.float32, device=dev)
            dwsa = torch.empty(tuple(wsa.shape), dtype=torch.float32, device=dev)
        call("ducosy_cbam_backward_acc" if acc else "ducosy_cbam_backward", ptr(dout), ptr(yb), ptr(sv["nb"][0]), ptr(sv["nb"][1]), ptr(sv["nv"][0]), ptr(sv["nv"][1]),
             ptr(sv["ca"]), ptr(sv["hidden"]), ptr(sv["chmax"]), ptr(sv["pooled"]), ptr(sv["sa"]), ptr(_f32c(fc0)), ptr(_f32c(fc2)),
             ptr(_f32c(wsa)), ptr(dn), ptr(dfc0), ptr(dfc2), ptr(dwsa), ptr(scratch), ptr(gs), B, H, W, Cn, dtype_code(yb.dtype),
             stream_ptr())
    return dn, ([None, None, None] if acc else [dfc0, dfc2, dwsa])


# ------------------------------------------------------------------ stand-alone building blocks (NCHW fp32, reference layout)
def channel_attention_nchw(x, fc0, fc2):
    """reference modules/model.py:20-24 on [B,C,H,W] fp32 (both pooling branches)."""
    B, Cn, H, W = x.shape
    hidden = fc0.shape[0]
    x, fc0, fc2 = _f32c(x), _f32c(fc0), _f32c(fc2)
    lib = _lib.load()
    with _dev(x):
        out = torch.empty_like(x)
        scratch = torch.empty(lib.ducosy_channel_attention_scratch_bytes(B, Cn) // 4, dtype=torch.float32, device=x.device)
        call("ducosy_channel_attention_nchw", ptr(x), ptr(fc0), ptr(fc2), ptr(out), ptr(scratch), B, Cn, hidden, H, W, stream_ptr())
    return out


def spatial_attention_nchw(x, w):
    """reference modules/model.py:34-39 on [B,C,H,W] fp32; w = conv.weight [1,2,k,k]."""
    B, Cn, H, W = x.shape
    k = w.shape[-1]
    x, w = _f32c(x), _f32c(w)
    lib = _lib.load()
    with _dev(x):
        out = torch.empty_like(x)
        scratch = torch.empty(lib.ducosy_spatial_attention_scratch_bytes(B, H, W) // 4, dtype=torch.float32, device=x.device)
        call("ducosy_spatial_attention_nchw", ptr(x), ptr(w), ptr(out), ptr(scratch), B, Cn, H, W, int(k), stream_ptr())
    return out


def nchw_to_nhwc_pad(x, pad, pad_mode, dtype=torch.float16):
    """fp32 [B,C,H,W] -> 16-bit [B,H+2p,W+2p,C] (reflect / zero padding)."""
    B, Cn, H, W = x.shape
    x = _f32c(x)
    with _dev(x):
        out = torch.empty((B, H + 2 * pad, W + 2 * pad, Cn), dtype=dtype, device=x.device)
        call("ducosy_nchw_to_nhwc_pad", ptr(x), ptr(out), B, Cn, H, W, int(pad), int(pad_mode), dtype_code(dtype), stream_ptr())
    return out


def nhwc_to_nchw(y):
    """16-bit [B,H,W,C] -> fp32 [B,C,H,W]."""
    B, H, W, Cn = y.shape
    assert y.is_contiguous()
    with _dev(y):
        out = torch.empty((B, Cn, H, W), dtype=torch.float32, device=y.device)
        call("ducosy_nhwc_to_nchw", ptr(y), ptr(out), B, Cn, H, W, dtype_code(y.dtype), stream_ptr())
    return out
