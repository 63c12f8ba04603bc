"""Drop-in for the reference's ``modules/model.py`` (reference file:line in each docstring).

Same constructors, same ``forward`` signatures, same parameter tree -- so ``state_dict()`` keys/shapes are
identical, existing ``.pth`` checkpoints load with ``strict=True`` and ``.apply(weights_init_normal)`` finds
the same 51 (generator) / 5 (discriminator) ``Conv2d`` modules.  The torch layers below are parameter
holders only: ``Generator.forward`` hands the whole network to ``libducosy_sm100.so``
(tcgen05 implicit-GEMM convolutions + fused InstanceNorm/CBAM/residual kernels, see ../csrc).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib

def _ordered_params(module):
    """Parameters in ``named_parameters()`` order.  A replica made by ``nn.DataParallel`` (reference modules/trainer.py:335-338)
    has no parameters of its own -- ``replicate`` stores the broadcast copies in ``_former_parameters`` of every sub-module, in
    registration order -- so those are collected instead (they carry the grad_fn that reduces gradients back to the master)."""
    ps = [p for _, p in module.named_parameters()]
    if ps or not getattr(module, "_is_replica", False):
        return ps
    out = []
    for m in module.modules():
        out.extend(getattr(m, "_former_parameters", {}).values())
    return out


def default_operand_dtype() -> int:
    """16-bit operand type of the tensor-core convolutions: fp16 (default; TF32-class 10-bit mantissa, the
    reference's own GPU precision) or bf16 via DUCOSY_PRECISION=bf16.  (DUCOSY_PRECISION=fp16x2 selects the split-operand
    arm of the generator's inference path only -- see ``inference_operand_dtype``; everything else then runs in fp16.)"""
    code = _lib.dtype_code(os.environ.get("DUCOSY_PRECISION", "fp16").lower())
    return _lib.F16 if code == _lib.F16X2 else code


def inference_operand_dtype(override=None) -> int:
    """Operand mode of ``Generator.forward`` without autograd (generate.py:96-97).  ``fp16x2`` is the <= 1 HU precision arm:
    every activation and weight is a (hi, lo) pair of fp16 values and each convolution tap runs three tensor-core products
    (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32 accumulate) -- fp32-class results at about a third of the fp16 throughput.
    Selected by ``Generator.precision = "fp16x2"`` or DUCOSY_PRECISION=fp16x2."""
    return _lib.dtype_code((override or os.environ.get("DUCOSY_PRECISION", "fp16")).lower())


class _DerivedState:
    """Mixin of Generator / Discriminator.  The per-device engines (packed 16-bit weights, workspaces, captured graphs) are
    derived state: never part of ``state_dict``, dropped by pickle / ``copy.deepcopy`` / ``torch.save(module)`` and rebuilt
    lazily.  The packed-weight caches are keyed by ``(data_ptr, _version)`` of every parameter, which sees
    ``load_state_dict``, optimiser steps, ``.to()`` and any in-place op on the parameter itself; writes through ``p.data``
    do not bump ``_version``, so the two module-level entry points that commonly do that (``.apply(fn)`` and ``._apply``)
    invalidate explicitly, and ``invalidate_packed_weights()`` is public for code that edits ``p.data`` by hand."""

    def invalidate_packed_weights(self):
        for eng in self._engines.values():
            eng.invalidate()

    def apply(self, fn):
        out = super().apply(fn)
        self.invalidate_packed_weights()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engines = {}
        return out

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engines"] = {}
        return state


def _forward_only(module, x):
    """The stand-alone building blocks run forward-only (inside ``Generator`` the same layers are differentiable through the
    fused training path): refuse to hand autograd a tensor without history instead of training on silent zeros."""
    if not x.is_cuda:
        raise RuntimeError(f"ducosy_gan_b200.{type(module).__name__} needs CUDA tensors on an sm_100 (B200) device; no CPU path exists")
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        raise RuntimeError(f"ducosy_gan_b200.{type(module).__name__}.forward on its own is forward-only: call it under torch.no_grad() "
                           "(training goes through Generator.forward, which differentiates these layers in its fused path)")
    _lib.check(_lib.load().ducosy_check_device(), "check_device")


class ChannelAttention(nn.Module):
    """reference modules/model.py:6-24 (parameters: fc.0.weight [C/r,C,1,1], fc.2.weight [C,C/r,1,1]).
    ``forward(x[B,C,H,W] fp32) -> x * sigmoid(fc(avgpool x) + fc(maxpool x))``, both branches (csrc/blocks.cu)."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.fc = nn.Sequential(nn.Conv2d(channels, channels // reduction, 1, bias=False), nn.ReLU(inplace=True),
                                nn.Conv2d(channels // reduction, channels, 1, bias=False))
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        from .. import ops
        _forward_only(self, x)
        return ops.channel_attention_nchw(x, self.fc[0].weight.detach(), self.fc[2].weight.detach())


class SpatialAttention(nn.Module):
    """reference modules/model.py:27-39 (parameter: conv.weight [1,2,k,k]).
    ``forward(x[B,C,H,W] fp32) -> x * sigmoid(conv_kxk(cat[mean_C x, max_C x]))`` (csrc/blocks.cu)."""

    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        from .. import ops
        _forward_only(self, x)
        return ops.spatial_attention_nchw(x, self.conv.weight.detach())


class CBAM(nn.Module):
    """reference modules/model.py:42-52: channel attention, then spatial attention."""

    def __init__(self, channels, reduction=16, kernel_size=7):
        super().__init__()
        self.channel_attention = ChannelAttention(channels, reduction)
        self.spatial_attention = SpatialAttention(kernel_size)

    def forward(self, x):
        _forward_only(self, x)
        return self.spatial_attention(self.channel_attention(x))


def _res_layers(c):
    return nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(c, c, 3), nn.InstanceNorm2d(c), nn.ReLU(inplace=True),
                         nn.ReflectionPad2d(1), nn.Conv2d(c, c, 3), nn.InstanceNorm2d(c))


def _residual_block_forward(module, x, cbam):
    """Stand-alone forward of a residual block on the generator's kernels: NCHW fp32 -> reflect-padded NHWC 16-bit ->
    tcgen05 conv + InstanceNorm statistics -> IN/ReLU/pad -> conv -> IN (+ channel attention, spatial attention) ->
    residual add -> NCHW fp32.  Same arithmetic as inside Generator.forward (16-bit operands, fp32 accumulate; behind the
    non-affine InstanceNorm the avg-pool branch of the channel attention is identically zero and is not evaluated)."""
    from .. import ops
    _forward_only(module, x)
    if x.dim() != 4:
        raise RuntimeError(f"expected [B,C,H,W], got {tuple(x.shape)}")
    B, Cn, H, W = x.shape
    conv1, conv2 = module.block[1], module.block[5]
    if Cn != conv1.in_channels:
        raise RuntimeError(f"expected {conv1.in_channels} channels, got {Cn}")
    Wt = min(W, 128)
    if Cn % 64 or Cn not in (64, 128, 256) or W % Wt or Wt & (Wt - 1) or Wt < 8 or H % (128 // Wt) or H < 2 or W < 2:
        raise RuntimeError(f"stand-alone residual block: channels must be 64/128/256, W one of 8/16/32/64 or a multiple of 128, H a "
                           f"multiple of 128/min(W,128) (got C={Cn}, {H}x{W})")
    if cbam is not None and (Cn != 256 or cbam.channel_attention.fc[0].out_channels != 16 or cbam.spatial_attention.conv.kernel_size != (7, 7)):
        raise RuntimeError("stand-alone ResidualBlockWithCBAM: the fused CBAM kernels are built for 256 channels, reduction 16, 7x7")
    dt = _lib.torch_dtype(default_operand_dtype())
    cache = module.__dict__.setdefault("_packed", {})
    sig = (conv1.weight.data_ptr(), conv1.weight._version, conv2.weight.data_ptr(), conv2.weight._version, dt)
    if cache.get("sig") != sig:
        cache.clear()
        cache.update(sig=sig, w1=ops.pack_conv_weight(conv1.weight, dt), w2=ops.pack_conv_weight(conv2.weight, dt))
    xp = ops.nchw_to_nhwc_pad(x, 1, _lib.PAD_REFLECT, dt)
    ya, part = ops.conv2d_nhwc(xp, cache["w1"], 3, 3, 1)
    pa = ops.in_apply_pad(ya, *ops.in_finalize(part, H * W), 1, _lib.PAD_REFLECT, _lib.ACT_RELU)
    yb, part = ops.conv2d_nhwc(pa, cache["w2"], 3, 3, 1)
    if cbam is None:
        scale, shift = ops.in_finalize(part, H * W)
        sa = None
    else:
        fc0 = cbam.channel_attention.fc[0].weight.detach().to(torch.float32).contiguous()
        fc2 = cbam.channel_attention.fc[2].weight.detach().to(torch.float32).contiguous()
        scale, shift = ops.in_finalize(part, H * W, fc0, fc2)          # InstanceNorm and channel attention as one affine map
        sa = ops.cbam_spatial_conv(ops.cbam_pool(yb, scale, shift), cbam.spatial_attention.conv.weight)
    out = ops.residual_apply_pad(yb, scale, shift, sa, xp, 1, 0, _lib.PAD_ZERO)
    return ops.nhwc_to_nchw(out)


class ResidualBlock(nn.Module):
    """reference modules/model.py:56-65: x + IN(conv(refpad(ReLU(IN(conv(refpad(x)))))))."""

    def __init__(self, in_features):
        super().__init__()
        self.block = _res_layers(in_features)

    def forward(self, x):
        return _residual_block_forward(self, x, None)

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_packed", None)
        return state


class ResidualBlockWithCBAM(nn.Module):
    """reference modules/model.py:68-87: x + CBAM(block(x))."""

    def __init__(self, in_features):
        super().__init__()
        self.block = _res_layers(in_features)
        self.cbam = CBAM(in_features)

    def forward(self, x):
        return _residual_block_forward(self, x, self.cbam)

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_packed", None)
        return state


class Generator(_DerivedState, nn.Module):
    """reference modules/model.py:90-115.  ``forward(x[B,Cin,H,W] fp32 cuda) -> [B,1,H,W] fp32``.

    H must be a multiple of 32 and W of 128 with W/4 in {32, 64, 128k} (512x512 is what generate.py and
    train.py use).  Only sm_100 devices are accepted.
    """

    def __init__(self, input_channels=1, num_residual_blocks=9, use_cbam=True):
        super().__init__()
        layers = [nn.ReflectionPad2d(3), nn.Conv2d(input_channels, 64, 7), nn.InstanceNorm2d(64), nn.ReLU(inplace=True)]
        ch = 64
        for _ in range(2):
            layers += [nn.Conv2d(ch, ch * 2, 3, stride=2, padding=1), nn.InstanceNorm2d(ch * 2), nn.ReLU(inplace=True)]
            ch *= 2
        block_cls = ResidualBlockWithCBAM if use_cbam else ResidualBlock
        layers += [block_cls(ch) for _ in range(num_residual_blocks)]
        for _ in range(2):
            layers += [nn.Upsample(scale_factor=2), nn.Conv2d(ch, ch // 2, 3, stride=1, padding=1),
                       nn.InstanceNorm2d(ch // 2), nn.ReLU(inplace=True)]
            ch //= 2
        layers += [nn.ReflectionPad2d(3), nn.Conv2d(ch, 1, 7), nn.Tanh()]
        self.model = nn.Sequential(*layers)
        self._cfg_tuple = (int(input_channels), int(num_residual_blocks), bool(use_cbam))
        self._engines = {}  # device index -> _GeneratorEngine (derived caches; never part of state_dict)
        self.precision = None  # None: DUCOSY_PRECISION (default fp16); "fp16" | "bf16" | "fp16x2" (inference_operand_dtype)

    # -- engine ------------------------------------------------------------------------------------------
    def _engine(self, device) -> "_GeneratorEngine":
        key = (device.index if device.index is not None else torch.cuda.current_device(),
               inference_operand_dtype(getattr(self, "precision", None)))
        eng = self._engines.get(key)
        if eng is None:
            eng = _GeneratorEngine(self._cfg_tuple, key[1], torch.device("cuda", key[0]))
            self._engines[key] = eng
        return eng

    def _ordered_params(self):
        return _ordered_params(self)

    def _train_pack_cache(self):
        """Packed operand forms of this module's weights for the training path (ops.pack_cache): per device, derived state."""
        dev = torch.cuda.current_device()
        return self._engines.setdefault(("train_packs", dev), _PackCache())

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("ducosy_gan_b200.Generator needs CUDA tensors on an sm_100 (B200) device; no CPU path exists")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self._ordered_params())):
            if x.dim() != 4 or x.shape[1] != self._cfg_tuple[0]:
                raise RuntimeError(f"expected input [B,{self._cfg_tuple[0]},H,W], got {tuple(x.shape)}")
            _lib.check(_lib.load().ducosy_check_device(), "check_device")
            return _GeneratorFunction.apply((self._cfg_tuple, self._train_pack_cache()), x, *self._ordered_params())
        eng = self._engine(x.device)
        eng.sync_weights(self._ordered_params())
        if eng.wants_graph(x, self):
            return eng.forward_graphed(x)
        return eng.forward(x)

    def forward_hu(self, px, slope, intercept, hu_min, hu_max):
        """Fused preprocess_dicom windowing (reference modules/preprocess.py:72-84) + forward:
        ``px`` int16 [B,H,W] stored values on the GPU -> [B,1,H,W] fp32.  Needs input_channels == 1."""
        eng = self._engine(px.device)
        eng.sync_weights(self._ordered_params())
        return eng.forward_hu(px, slope, intercept, hu_min, hu_max)


class _PackCache(dict):
    """ops.pack_cache store of one module on one device; ``invalidate`` makes it an engine-like entry of ``_engines``."""

    def invalidate(self):
        self.clear()


class _GeneratorFunction(torch.autograd.Function):
    """autograd bridge of the training path (reference modules/trainer.py:455-500): the forward keeps the raw conv outputs
    and InstanceNorm statistics, the backward runs ..training.generator_backward.  The image gradient flows to channel 0
    (the CT slice); the mask channels the reference concatenates (trainer.py:430-450) are data."""

    @staticmethod
    def forward(ctx, cfg_and_cache, x, *params):
        from .. import ops, training
        cfg, cache = cfg_and_cache
        dtype = torch.float16 if default_operand_dtype() == _lib.F16 else torch.bfloat16
        xs = x.detach().to(dtype=torch.float32).contiguous()
        ps = [p.detach() for p in params]
        with torch.cuda.device(x.device), ops.pack_cache(cache):
            out, saved = training.generator_forward_train(ps, cfg, xs, dtype)
        saved.pop("out")
        ctx.cfg, ctx.saved, ctx.params, ctx.in_shape, ctx.cache = cfg, saved, ps, tuple(x.shape), cache
        ctx.param_refs = params
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        from .. import ops, training
        (out,) = ctx.saved_tensors
        saved = ctx.saved
        saved["out"] = out
        # a leaf parameter that already holds a contiguous fp32 gradient on this device (the views of a data-parallel bucket, an
        # earlier backward) lets the weight-gradient kernels accumulate into it directly; autograd then gets None for it
        direct = os.environ.get("DUCOSY_WGRAD_DIRECT", "1") != "0"
        grad_out = [p.grad if (direct and p.is_leaf and p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                               and p.grad.device == out.device and p.grad.shape == p.shape) else None for p in ctx.param_refs]
        with torch.cuda.device(out.device), ops.pack_cache(ctx.cache):
            grads, dx = training.generator_backward(ctx.params, ctx.cfg, saved, dout, ctx.needs_input_grad[1], grad_out)
            if dx is not None and ctx.in_shape[1] > 1:
                full = torch.zeros(ctx.in_shape, dtype=torch.float32, device=out.device)
                full[:, :1] = dx
                dx = full
        ctx.saved = None
        # exact-zero gradients (dead biases): a parameter that already holds a gradient tensor (flat data-parallel bucket, an
        # earlier backward) is left alone -- adding zeros would be two launches per bias for nothing; one without gets real zeros
        grads = [g if g is not None or (p.is_leaf and p.grad is not None) else torch.zeros_like(p, dtype=torch.float32)
                 for g, p in zip(grads, ctx.param_refs)]
        ctx.param_refs = None
        return (None, dx, *grads)


class _GeneratorEngine:
    """Per-device derived state of a Generator: packed 16-bit weights and the activation workspace."""

    def __init__(self, cfg_tuple, dtype_code, device):
        self.device = device
        self.cfg = _lib.GenConfig(cfg_tuple[0], cfg_tuple[1], int(cfg_tuple[2]), dtype_code)
        lib = _lib.load()
        self.num_params = lib.ducosy_generator_num_params(C.byref(self.cfg))
        with torch.cuda.device(device):
            self.packed = torch.empty(lib.ducosy_generator_packed_bytes(C.byref(self.cfg)) + 256,
                                      dtype=torch.uint8, device=device)
        self._versions = None
        self._ws = {}
        self._graphs = {}        # input shape -> (CUDAGraph, static input, static output, dedicated workspace)

    def _packed_ptr(self):
        return (self.packed.data_ptr() + 255) // 256 * 256

    def invalidate(self):
        self._versions = None

    # -- small-batch calls replayed from a CUDA graph ------------------------------------------------------------------
    GRAPH_MAX_PIXELS = 2 * 512 * 512

    def wants_graph(self, x, module):
        """generate.py:89-102 calls ``model(x)`` one slice at a time: ~100 dependent kernels of a few microseconds each, so the
        call is bound by launch latency, not by the GPU.  Such calls (eval mode, no autograd, at most two 512x512 slices) are
        captured once per input shape and replayed.  The graph holds only the launches: the packed weights it reads are
        refreshed outside it whenever a parameter changes, so it never goes stale.  DUCOSY_FORWARD_GRAPH=0 disables it."""
        return (not module.training and x.dim() == 4 and x.shape[1] == self.cfg.input_channels
                and x.shape[0] * x.shape[2] * x.shape[3] <= self.GRAPH_MAX_PIXELS
                and not getattr(module, "_is_replica", False)
                and os.environ.get("DUCOSY_FORWARD_GRAPH", "1") != "0"
                and not torch.cuda.is_current_stream_capturing())

    def forward_graphed(self, x):
        key = tuple(x.shape)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(x)
        graph, xin, out, _ = entry
        with torch.cuda.device(self.device):
            xin.copy_(x)                      # any dtype / layout the eager call accepts
            graph.replay()
            return out.clone()                # the static buffer is overwritten by the next call

    def _capture(self, x):
        B, _, H, W = x.shape
        need = _lib.load().ducosy_generator_workspace_bytes(C.byref(self.cfg), B, H, W)
        if need == 0:
            raise _lib.DucosyError(f"unsupported generator input shape B={B} H={H} W={W}: "
                                   "B >= 1, H a multiple of 32, W of 128, W/4 in {32,64,128k}")
        with torch.cuda.device(self.device):
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            wptr = (ws.data_ptr() + 1023) // 1024 * 1024
            xin = torch.empty((B, self.cfg.input_channels, H, W), dtype=torch.float32, device=self.device)
            out = torch.empty((B, 1, H, W), dtype=torch.float32, device=self.device)
            xin.copy_(x)

            def launch():
                _lib.call("ducosy_generator_forward", C.byref(self.cfg), C.c_void_p(self._packed_ptr()), _lib.ptr(xin),
                          _lib.ptr(out), B, H, W, C.c_void_p(wptr), ws.numel() - 1024, _lib.stream_ptr())

            launch()                          # eager once: per-device kernel attributes and TMA descriptors exist before capture
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                launch()
        if len(self._graphs) >= 4:            # a handful of shapes at most; drop the oldest
            self._graphs.pop(next(iter(self._graphs)))
        entry = self._graphs[tuple(x.shape)] = (graph, xin, out, ws)
        return entry

    def sync_weights(self, params):
        """Re-pack when any fp32 master parameter changed (load_state_dict, optimizer step, .to())."""
        sig = tuple((p.data_ptr(), p._version) for p in params)
        if sig == self._versions:
            return
        if len(params) != self.num_params:
            raise RuntimeError(f"expected {self.num_params} parameter tensors, found {len(params)}")
        keep = []
        for p in params:
            t = p.detach()
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
        arr = (C.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
        with torch.cuda.device(self.device):
            _lib.call("ducosy_generator_pack", C.byref(self.cfg), arr, len(keep), C.c_void_p(self._packed_ptr()),
                      _lib.stream_ptr())
        self._keepalive = keep
        self._versions = sig

    def workspace(self, B, H, W):
        """One live workspace per engine.  A request that fits the live one (a ragged tail of a volume, a smaller batch)
        reuses it: the library lays its buffers out from the sizes of the call, so any large-enough block will do."""
        need = _lib.load().ducosy_generator_workspace_bytes(C.byref(self.cfg), B, H, W)
        if need == 0:
            raise _lib.DucosyError(f"unsupported generator input shape B={B} H={H} W={W}: "
                                   "B >= 1, H a multiple of 32, W of 128, W/4 in {32,64,128k}")
        ws = self._ws.get("ws")
        if ws is None or ws.numel() - 1024 < need:
            self._ws.clear()
            ws = None
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            self._ws["ws"] = ws
        return ws, (ws.data_ptr() + 1023) // 1024 * 1024, ws.numel() - 1024

    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.cfg.input_channels:
            raise RuntimeError(f"expected input [B,{self.cfg.input_channels},H,W], got {tuple(x.shape)}")
        x = x.to(dtype=torch.float32).contiguous()
        B, _, H, W = x.shape
        with torch.cuda.device(self.device):
            _, wptr, wbytes = self.workspace(B, H, W)
            out = torch.empty((B, 1, H, W), dtype=torch.float32, device=self.device)
            _lib.call("ducosy_generator_forward", C.byref(self.cfg), C.c_void_p(self._packed_ptr()), _lib.ptr(x),
                      _lib.ptr(out), B, H, W, C.c_void_p(wptr), wbytes, _lib.stream_ptr())
        return out

    def forward_hu(self, px, slope, intercept, lo, hi, out=None):
        if px.dtype != torch.int16 or px.dim() != 3 or not px.is_contiguous():
            raise RuntimeError("forward_hu expects a contiguous int16 [B,H,W] tensor of stored pixel values")
        B, H, W = px.shape
        with torch.cuda.device(self.device):
            _, wptr, wbytes = self.workspace(B, H, W)
            if out is None:
                out = torch.empty((B, 1, H, W), dtype=torch.float32, device=self.device)
            _lib.call("ducosy_generator_forward_hu", C.byref(self.cfg), C.c_void_p(self._packed_ptr()), _lib.ptr(px),
                      float(slope), float(intercept), float(lo), float(hi), _lib.ptr(out), B, H, W, C.c_void_p(wptr),
                      wbytes, _lib.stream_ptr())
        return out


class Discriminator(_DerivedState, nn.Module):
    """reference modules/model.py:118-131 (PatchGAN).  ``forward(img[B,1,H,W] fp32 cuda) -> [B,1,H/16,W/16] fp32``;
    H and W multiples of 256.  Differentiable w.r.t. its parameters and the input image: under autograd the forward and
    the backward both run in libducosy_sm100.so (BASELINE config 3: forward/backward with the MSE adversarial loss of
    reference modules/trainer.py:347,518-524)."""

    def __init__(self, input_channels=1):
        super().__init__()
        layers, ch_in = [], input_channels
        for i, ch_out in enumerate((64, 128, 256, 512)):
            layers.append(nn.Conv2d(ch_in, ch_out, 4, stride=2, padding=1))
            if i > 0:
                layers.append(nn.InstanceNorm2d(ch_out))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
            ch_in = ch_out
        layers += [nn.ZeroPad2d((1, 0, 1, 0)), nn.Conv2d(512, 1, 4, padding=1)]
        self.model = nn.Sequential(*layers)
        self._input_channels = int(input_channels)
        self._engines = {}

    def forward(self, img):
        if self._input_channels != 1:
            raise NotImplementedError("ducosy_gan_b200.Discriminator supports input_channels == 1 (what the reference trains)")
        if not img.is_cuda:
            raise RuntimeError("ducosy_gan_b200.Discriminator needs CUDA tensors on an sm_100 (B200) device; no CPU path exists")
        dev = img.device
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), default_operand_dtype())
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = _DiscriminatorEngine(key[1], torch.device("cuda", key[0]))
        params = _ordered_params(self)
        eng.sync_weights(params)
        if torch.is_grad_enabled() and (img.requires_grad or any(p.requires_grad for p in params)):
            return _DiscriminatorFunction.apply(eng, img, *params)
        return eng.forward(img)


class _DiscriminatorFunction(torch.autograd.Function):
    """autograd bridge: forward keeps the activation workspace, backward calls ducosy_discriminator_backward."""

    @staticmethod
    def forward(ctx, eng, img, *params):
        img = img.detach().to(dtype=torch.float32).contiguous()
        out, ws = eng.forward(img, keep_workspace=True)
        ctx.eng, ctx.ws, ctx.img = eng, ws, img
        ctx.shapes = [tuple(p.shape) for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        grads, dx = ctx.eng.backward(ctx.img, dout, ctx.ws, ctx.shapes, ctx.needs_input_grad[1])
        ctx.ws = None
        return (None, dx, *grads)


class _DiscriminatorEngine:
    """Per-device packed weights + workspace of a Discriminator."""

    def __init__(self, dtype_code, device):
        self.device, self.dtype_code = device, dtype_code
        with torch.cuda.device(device):
            self.packed = torch.empty(_lib.load().ducosy_discriminator_packed_bytes() + 256, dtype=torch.uint8, device=device)
        self._versions, self._ws = None, {}

    def _packed_ptr(self):
        return (self.packed.data_ptr() + 255) // 256 * 256

    def invalidate(self):
        self._versions = None

    def sync_weights(self, params):
        sig = tuple((p.data_ptr(), p._version) for p in params)
        if sig == self._versions:
            return
        keep = [p.detach().to(device=self.device, dtype=torch.float32).contiguous() for p in params]
        arr = (C.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
        with torch.cuda.device(self.device):
            _lib.call("ducosy_discriminator_pack", arr, len(keep), C.c_void_p(self._packed_ptr()), self.dtype_code,
                      _lib.stream_ptr())
        self._keepalive, self._versions = keep, sig

    @staticmethod
    def _aligned(t):
        return (t.data_ptr() + 1023) // 1024 * 1024

    def forward(self, img, keep_workspace=False):
        if img.dim() != 4 or img.shape[1] != 1:
            raise RuntimeError(f"expected input [B,1,H,W], got {tuple(img.shape)}")
        img = img.to(dtype=torch.float32).contiguous()
        B, _, H, W = img.shape
        with torch.cuda.device(self.device):
            ws = None if keep_workspace else self._ws.get((B, H, W))
            if ws is None:
                need = _lib.load().ducosy_discriminator_workspace_bytes(B, H, W)
                if need == 0:
                    raise _lib.DucosyError(f"unsupported discriminator input shape {tuple(img.shape)}: H, W multiples of 256")
                ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
                if not keep_workspace:   # inference: one cached workspace; training: one per call, owned by autograd
                    self._ws.clear()
                    self._ws[(B, H, W)] = ws
            out = torch.empty((B, 1, H // 16, W // 16), dtype=torch.float32, device=self.device)
            _lib.call("ducosy_discriminator_forward", C.c_void_p(self._packed_ptr()), _lib.ptr(img), _lib.ptr(out), B, H, W,
                      C.c_void_p(self._aligned(ws)), ws.numel() - 1024, self.dtype_code, _lib.stream_ptr())
        return (out, ws) if keep_workspace else out

    def backward(self, img, dout, fwd_ws, shapes, need_dx):
        B, _, H, W = img.shape
        dout = dout.detach().to(dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            need = _lib.load().ducosy_discriminator_backward_workspace_bytes(B, H, W)
            bws = self._bws.get((B, H, W)) if hasattr(self, "_bws") else None
            if bws is None:
                self._bws = {(B, H, W): torch.empty(need + 1024, dtype=torch.uint8, device=self.device)}
                bws = self._bws[(B, H, W)]
            grads = [torch.empty(s, dtype=torch.float32, device=self.device) for s in shapes]
            dx = torch.empty_like(img) if need_dx else None
            arr = (C.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
            _lib.call("ducosy_discriminator_backward", C.c_void_p(self._packed_ptr()), _lib.ptr(img), _lib.ptr(dout),
                      C.c_void_p(self._aligned(fwd_ws)), arr, _lib.ptr(dx), B, H, W, C.c_void_p(self._aligned(bws)),
                      bws.numel() - 1024, self.dtype_code, _lib.stream_ptr())
        return grads, dx


def weights_init_normal(m):
    """reference modules/model.py:134-140: Conv* weights ~ N(0, 0.02); BatchNorm2d weights ~ N(1, 0.02), bias 0."""
    name = type(m).__name__
    # same draws as the reference's ``.data`` form; writing through the Parameter bumps its ``_version``, which keys the
    # packed-weight caches (a ``.data`` write is invisible to them)
    if "Conv" in name:
        nn.init.normal_(m.weight, 0.0, 0.02)
    elif "BatchNorm2d" in name:
        nn.init.normal_(m.weight, 1.0, 0.02)
        nn.init.constant_(m.bias, 0.0)
