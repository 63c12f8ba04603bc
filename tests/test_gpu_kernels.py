"""Kernel-level parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI.

Integer / mask kernels: bit-exact against the oracle.  Floating-point kernels: against a plain PyTorch fp32
evaluation of the same op on the same (16-bit-rounded) operands, tolerance stated per test.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ducosy_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from ducosy_gan_b200 import _lib, ops as _ops
    _lib.check(_lib.load().ducosy_check_device(), "check_device")
    return _ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


DTYPES = [torch.float16, torch.bfloat16]


# ------------------------------------------------------------------ HU kernels (bit-exact)
@pytest.mark.parametrize("slope,intercept", [(1.0, -1024.0), (2.0, -1000.0), (0.5, -512.25)])
@pytest.mark.parametrize("n", [0, 5, 8, 4099, 512 * 512])
def test_hu_window_bit_exact(ops, slope, intercept, n):
    px = orc.synthetic_volume(1, 512, 512, seed=7)[0].reshape(-1)[:n].copy()
    if n >= 8:
        px[:6] = [874, 873, 24, 23, 1274, 1275]
    d = torch.from_numpy(px).cuda()
    s, l = ops.hu_window(d, slope, intercept)
    assert np.array_equal(s.cpu().numpy(), orc.hu_window(px, slope, intercept, -150, 250).astype(np.float32))
    assert np.array_equal(l.cpu().numpy(), orc.hu_window(px, slope, intercept, -1000, -150).astype(np.float32))


def test_hu_thresholds_bit_exact(ops):
    px = orc.synthetic_volume(2, 64, 96, seed=8)
    px.flat[:7] = [24, 23, 25, 724, 725, 1224, 1223]
    b, l, o = ops.hu_thresholds(torch.from_numpy(px).cuda(), 1.0, -1024.0)
    rb, rl, ro = orc.threshold_candidates(orc.stored_to_hu(px, 1.0, -1024.0))
    assert np.array_equal(b.cpu().numpy(), rb) and np.array_equal(l.cpu().numpy(), rl) and np.array_equal(o.cpu().numpy(), ro)


@pytest.mark.parametrize("slope,intercept", [(1.0, -1024.0), (2.0, -1000.0), (0.5, -512.25)])
@pytest.mark.parametrize("n", [3, 4096 + 5, 300 * 1024])
def test_dewindow_composite_bit_exact(ops, slope, intercept, n):
    rng = np.random.Generator(np.random.PCG64(9))
    raw = rng.integers(0, 2500, size=n, dtype=np.int16)
    raw[:3] = [874, 24, 1275]
    ys = rng.uniform(-1, 1, size=n).astype(np.float32)
    yl = rng.uniform(-1, 1, size=n).astype(np.float32)
    ys[:2] = [-1.0, 1.0]
    merged, sp, lp, mk = ops.dewindow_composite(torch.from_numpy(raw).cuda(), torch.from_numpy(ys).cuda(),
                                                torch.from_numpy(yl).cuda(), slope, intercept, want_parts=True)
    rsp = orc.dewindow_to_stored(ys, slope, intercept, -150, 250)
    rlp = orc.dewindow_to_stored(yl, slope, intercept, -1000, -150)
    rm, sm, lm = orc.composite(raw, rsp, rlp, slope, intercept)
    assert np.array_equal(sp.cpu().numpy(), rsp) and np.array_equal(lp.cpu().numpy(), rlp)
    assert np.array_equal(merged.cpu().numpy(), rm)
    assert np.array_equal(mk.cpu().numpy(), sm.astype(np.uint8) | (lm.astype(np.uint8) << 1))


# ------------------------------------------------------------------ tcgen05 implicit-GEMM convolution
def _conv_case(ops, dtype, B, H, W, Cin, Cout, k, stride, pad, seed):
    """Random NHWC input, conv through the C ABI, compare with torch fp32 conv on the rounded operands."""
    x = _rand((B, Cin, H, W), seed).to(dtype)
    w = _rand((Cout, Cin, k, k), seed + 1, 0.05)
    xp = F.pad(x.float(), (pad, pad, pad, pad)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    wp = ops.pack_conv_weight(w.cuda(), dtype)
    y, partials = ops.conv2d_nhwc(xp, wp, k, k, stride)
    torch.cuda.synchronize()
    w_r = wp.float().view(Cout, k, k, Cin).permute(0, 3, 1, 2).contiguous()      # rounded weights back in OIHW
    ref = F.conv2d(F.pad(x.float().cuda(), (pad, pad, pad, pad)), w_r, stride=stride)  # fp32
    got = y.float().permute(0, 3, 1, 2)
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    tol = eps * ref.abs().max().item() + 1e-3
    assert (got - ref).abs().max().item() <= tol, ((got - ref).abs().max().item(), tol)
    # statistics of the stored values: sum / sum of squares / max per (sample, channel)
    yf = y.float()
    s1 = partials[:, :, 0, :].sum(1)
    s2 = partials[:, :, 1, :].sum(1)
    mx = partials[:, :, 2, :].amax(1)
    assert torch.allclose(s1, yf.sum((1, 2)), rtol=1e-3, atol=0.05)
    assert torch.allclose(s2, (yf * yf).sum((1, 2)), rtol=1e-3, atol=0.05)
    assert torch.equal(mx, yf.amax((1, 2)))
    return y, partials


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_1x1(ops, dtype):
    _conv_case(ops, dtype, B=1, H=8, W=128, Cin=64, Cout=64, k=1, stride=1, pad=0, seed=1)
    _conv_case(ops, dtype, B=2, H=16, W=16, Cin=192, Cout=128, k=1, stride=1, pad=0, seed=2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv3x3_s1_c256(ops, dtype):
    _conv_case(ops, dtype, B=2, H=32, W=32, Cin=256, Cout=256, k=3, stride=1, pad=1, seed=3)


def test_conv3x3_s1_c256_full_res_block_shape(ops):
    # the headline shape: 256 -> 256 on 128x128 (one output row per tile), more tiles than SMs
    _conv_case(ops, torch.float16, B=2, H=128, W=128, Cin=256, Cout=256, k=3, stride=1, pad=1, seed=4)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv3x3_s2(ops, dtype):
    _conv_case(ops, dtype, B=1, H=64, W=128, Cin=64, Cout=128, k=3, stride=2, pad=1, seed=5)
    _conv_case(ops, dtype, B=2, H=32, W=64, Cin=128, Cout=256, k=3, stride=2, pad=1, seed=6)


def test_conv4x4_s2_patchgan_shapes(ops):
    _conv_case(ops, torch.float16, B=1, H=64, W=64, Cin=64, Cout=128, k=4, stride=2, pad=1, seed=7)
    _conv_case(ops, torch.float16, B=2, H=16, W=64, Cin=256, Cout=512, k=4, stride=2, pad=1, seed=8)   # 2 n-blocks


def test_conv_bias_lrelu_epilogue(ops):
    dtype = torch.float16
    x = _rand((1, 64, 16, 64), 11).to(dtype)
    w = _rand((64, 64, 1, 1), 12, 0.1)
    bias = _rand((64,), 13, 0.5).cuda()
    xp = x.float().permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    wp = ops.pack_conv_weight(w.cuda(), dtype)
    y, partials = ops.conv2d_nhwc(xp, wp, 1, 1, 1, want_stats=False, bias=bias, act=ops.ACT_LRELU02)
    assert partials is None
    ref = F.leaky_relu(F.conv2d(x.float().cuda(), wp.float().view(64, 64, 1, 1), bias), 0.2)
    assert (y.float().permute(0, 3, 1, 2) - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(1, 16, 32, 128, 64), (2, 32, 32, 256, 128)])
def test_upconv2x_matches_upsample_conv(ops, dtype, shape):
    B, Hs, Ws, Cin, Cout = shape
    x = _rand((B, Cin, Hs, Ws), 21).to(dtype)
    w = _rand((Cout, Cin, 3, 3), 22, 0.05)
    xp = F.pad(x.float(), (1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    y, partials = ops.upconv2x_nhwc(xp, ops.pack_upconv_weight(w.cuda(), dtype))
    # reference in fp32 with the original (un-summed) weights: differences come from rounding the pre-summed weights
    ref = F.conv2d(F.interpolate(x.float().cuda(), scale_factor=2, mode="nearest"), w.cuda(), padding=1)
    got = y.float().permute(0, 3, 1, 2)
    eps = 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6
    assert (got - ref).abs().max().item() <= eps * ref.abs().max().item() + 1e-3
    yf = y.float()
    assert torch.allclose(partials[:, :, 0, :].sum(1), yf.sum((1, 2)), rtol=1e-3, atol=0.05)
    assert torch.equal(partials[:, :, 2, :].amax(1), yf.amax((1, 2)))


@pytest.mark.parametrize("dtype", DTYPES)
def test_upconv2x_merged_phases(ops, dtype):
    B, Hs, Ws, Cin, Cout = 2, 16, 32, 128, 64
    x = _rand((B, Cin, Hs, Ws), 23).to(dtype)
    w = _rand((Cout, Cin, 3, 3), 24, 0.05)
    xp = F.pad(x.float(), (1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    y, partials = ops.upconv2x_merged_nhwc(xp, ops.pack_upconv_merged_weight(w.cuda(), dtype))
    y4, _ = ops.upconv2x_nhwc(xp, ops.pack_upconv_weight(w.cuda(), dtype))
    assert torch.equal(y, y4)                      # same operands, same k order per phase: bit-identical to the 4-phase path
    ref = F.conv2d(F.interpolate(x.float().cuda(), scale_factor=2, mode="nearest"), w.cuda(), padding=1)
    eps = 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6
    assert (y.float().permute(0, 3, 1, 2) - ref).abs().max().item() <= eps * ref.abs().max().item() + 1e-3
    yf = y.float()
    assert partials.shape == (B, Hs * Ws // 128, 3, Cout)
    assert torch.allclose(partials[:, :, 0, :].sum(1), yf.sum((1, 2)), rtol=1e-3, atol=0.05)
    assert torch.allclose(partials[:, :, 1, :].sum(1), (yf * yf).sum((1, 2)), rtol=1e-3, atol=0.05)
    assert torch.equal(partials[:, :, 2, :].amax(1), yf.amax((1, 2)))


# ------------------------------------------------------------------ stem im2col
def test_stem_im2col_and_hu_variant(ops):
    x = _rand((2, 3, 16, 24), 31)
    a = ops.stem_im2col(x.cuda(), torch.float16)
    cols = F.unfold(F.pad(x, (3, 3, 3, 3), mode="reflect"), 7)          # [B, Cin*49, H*W], k = c*49 + r*7 + s
    ref = cols.transpose(1, 2).reshape(2, 16, 24, 147).to(torch.float16)
    assert torch.equal(a[..., :147].cpu(), ref)
    assert torch.count_nonzero(a[..., 147:]).item() == 0
    px = orc.synthetic_volume(1, 16, 24, seed=32)
    ah = ops.stem_im2col_hu(torch.from_numpy(px).cuda(), 1.0, -1024.0, -150.0, 250.0, torch.float16)
    xw = torch.from_numpy(orc.hu_window(px, 1.0, -1024.0, -150, 250).astype(np.float32))[:, None]
    refh = F.unfold(F.pad(xw, (3, 3, 3, 3), mode="reflect"), 7).transpose(1, 2).reshape(1, 16, 24, 49).to(torch.float16)
    assert torch.equal(ah[..., :49].cpu(), refh)


@pytest.mark.parametrize("dtype", DTYPES)
def test_stem_fused_two_pass(ops, dtype):
    # conv7x7(reflect pad 3) -> IN -> ReLU -> zero pad 1, from a fp32 tensor and from stored pixels through the HU window
    w = _rand((64, 1, 7, 7), 33, 0.1)
    wp = ops.pack_stem_weight(w.cuda(), dtype)
    w_r = wp.float()[:, :49].reshape(64, 1, 7, 7)
    tol = 6e-3 if dtype == torch.float16 else 4e-2
    x = _rand((2, 1, 32, 128), 34).to(dtype).float()        # representable in the operand type
    out = ops.stem_fused(wp, x=x.cuda())
    ref = F.pad(F.relu(F.instance_norm(F.conv2d(F.pad(x.cuda(), (3, 3, 3, 3), mode="reflect"), w_r))), (1, 1, 1, 1))
    assert (out.float().permute(0, 3, 1, 2) - ref).abs().max().item() < tol
    assert torch.count_nonzero(out[:, 0]).item() == 0 and torch.count_nonzero(out[:, :, -1]).item() == 0
    px = orc.synthetic_volume(2, 16, 128, seed=35)
    outh = ops.stem_fused(wp, px=torch.from_numpy(px).cuda(), window=(1.0, -1024.0, -150.0, 250.0))
    xw = torch.from_numpy(orc.hu_window(px, 1.0, -1024.0, -150, 250).astype(np.float32))[:, None].to(dtype).float().cuda()
    refh = F.pad(F.relu(F.instance_norm(F.conv2d(F.pad(xw, (3, 3, 3, 3), mode="reflect"), w_r))), (1, 1, 1, 1))
    assert (outh.float().permute(0, 3, 1, 2) - refh).abs().max().item() < tol


# ------------------------------------------------------------------ InstanceNorm / CBAM / residual kernels
def _fake_partials(y):
    """Build [B, tiles, 3, C] partials from an NHWC tensor exactly as the conv epilogue would (128-pixel tiles)."""
    B, H, W, Cn = y.shape
    t = y.float().reshape(B, H * W // 128, 128, Cn)
    return torch.stack([t.sum(2), (t * t).sum(2), t.amax(2)], dim=2).contiguous()


@pytest.mark.parametrize("dtype", DTYPES)
def test_instance_norm_apply_pad(ops, dtype):
    y = (_rand((2, 16, 32, 64), 41) * 3 + 0.7).to(dtype).cuda()
    scale, shift = ops.in_finalize(_fake_partials(y), 16 * 32)
    ref = F.instance_norm(y.float().permute(0, 3, 1, 2))
    for pad, mode, act in [(1, ops.PAD_REFLECT, ops.ACT_RELU), (3, ops.PAD_REFLECT, ops.ACT_RELU),
                           (1, ops.PAD_ZERO, ops.ACT_RELU), (0, ops.PAD_ZERO, ops.ACT_NONE),
                           (1, ops.PAD_ZERO, ops.ACT_LRELU02)]:
        out = ops.in_apply_pad(y, scale, shift, pad, mode, act)
        r = ref
        r = F.relu(r) if act == ops.ACT_RELU else (F.leaky_relu(r, 0.2) if act == ops.ACT_LRELU02 else r)
        if pad:
            r = F.pad(r, (pad,) * 4, mode="reflect" if mode == ops.PAD_REFLECT else "constant")
        tol = 4e-3 if dtype == torch.float16 else 3e-2
        assert (out.float().permute(0, 3, 1, 2) - r).abs().max().item() < tol


def test_cbam_chain_matches_oracle(ops):
    dtype = torch.float16
    B, H, W, Cn = 2, 16, 32, 256
    y = (_rand((B, H, W, Cn), 51) * 2 + 0.3).to(dtype).cuda()
    res = _rand((B, H, W, Cn), 52).to(dtype).cuda()
    fc0 = _rand((16, 256, 1, 1), 53, 0.2).cuda()
    fc2 = _rand((256, 16, 1, 1), 54, 0.2).cuda()
    wsa = _rand((1, 2, 7, 7), 55, 0.2).cuda()
    scale, shift = ops.in_finalize(_fake_partials(y), H * W, fc0.contiguous(), fc2.contiguous())
    pooled = ops.cbam_pool(y, scale, shift)
    sa = ops.cbam_spatial_conv(pooled, wsa)
    res_pad = F.pad(res.float().permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1).contiguous().to(dtype)
    out = ops.residual_apply_pad(y, scale, shift, sa, res_pad, 1, 1, ops.PAD_REFLECT)
    # oracle (reference modules/model.py:20-24,34-39,83-87) in fp32 on the same rounded inputs
    n = orc.instance_norm(y.float().permute(0, 3, 1, 2))
    c = orc.spatial_attention(orc.channel_attention(n, fc0, fc2), wsa)
    ref = F.pad(res.float().permute(0, 3, 1, 2) + c, (1, 1, 1, 1), mode="reflect")
    assert (out.float().permute(0, 3, 1, 2) - ref).abs().max().item() < 8e-3
    # plain ResidualBlock variant (no attention)
    scale2, shift2 = ops.in_finalize(_fake_partials(y), H * W)
    out2 = ops.residual_apply_pad(y, scale2, shift2, None, res_pad, 1, 1, ops.PAD_ZERO)
    ref2 = F.pad(res.float().permute(0, 3, 1, 2) + n, (1, 1, 1, 1))
    assert (out2.float().permute(0, 3, 1, 2) - ref2).abs().max().item() < 8e-3


@pytest.mark.parametrize("dtype", DTYPES)
def test_out_conv7x7_tanh(ops, dtype):
    B, H, W = 2, 16, 128
    x = _rand((B, 64, H, W), 61).to(dtype)
    w = _rand((1, 64, 7, 7), 62, 0.02)
    bias = torch.tensor([0.05])
    xp = F.pad(x.float(), (3, 3, 3, 3), mode="reflect").permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    wp = ops.pack_out_weight(w.cuda(), dtype)
    out = ops.out_conv7x7_tanh(xp, wp, bias.cuda())
    w_r = wp.float().view(7, 8, 64)[:, :7, :].permute(2, 0, 1).reshape(1, 64, 7, 7)
    ref = torch.tanh(F.conv2d(F.pad(x.float().cuda(), (3, 3, 3, 3), mode="reflect"), w_r, bias.cuda()))
    assert (out - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_out_conv_fused_equals_apply_then_conv(ops, dtype):
    B, H, W = 2, 16, 128
    y = (_rand((B, H, W, 64), 71) * 2 + 0.5).to(dtype).cuda()
    scale, shift = ops.in_finalize(_fake_partials(y), H * W)
    w = _rand((1, 64, 7, 7), 72, 0.02)
    bias = torch.tensor([-0.03]).cuda()
    wp = ops.pack_out_weight(w.cuda(), dtype)
    ref = ops.out_conv7x7_tanh(ops.in_apply_pad(y, scale, shift, 3, ops.PAD_REFLECT, ops.ACT_RELU), wp, bias)
    got = ops.out_conv7x7_tanh_fused(y, scale, shift, wp, bias)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("case", [(2, 32, 32, 256, 256, 3, 1), (1, 64, 64, 64, 128, 3, 2), (2, 32, 64, 128, 256, 4, 2),
                                  (1, 16, 128, 192, 128, 1, 1), (2, 32, 64, 128, 64, 3, 1), (1, 16, 128, 64, 64, 1, 1)])
def test_conv2d_wgrad_matches_autograd(ops, case):
    B, H, W, Cin, Cout, k, stride = case
    dtype = torch.float16
    pad = 0 if k == 1 else 1
    x = _rand((B, Cin, H, W), 81).to(dtype)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    dy = _rand((B, Cout, Ho, Wo), 82).to(dtype)
    xp = F.pad(x.float(), (pad,) * 4).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dyn = dy.float().permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dw = ops.conv2d_wgrad_nhwc(xp, dyn, k, k, stride)                      # [Cout][k*k*Cin]
    w0 = torch.zeros((Cout, Cin, k, k), device="cuda", requires_grad=True)
    F.conv2d(x.float().cuda(), w0, stride=stride, padding=pad).backward(dy.float().cuda())
    ref = w0.grad.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin)            # packed layout: (r*kw+s)*Cin + c
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-2, (err, ref.abs().max().item())


def test_in_backward_matches_autograd(ops):
    dtype = torch.float16
    B, H, W, Cn = 2, 16, 32, 128
    y = (_rand((B, H, W, Cn), 91) * 1.5 + 0.8).to(dtype).cuda()
    da = _rand((B, H, W, Cn), 92).to(dtype).cuda()
    scale, shift = ops.in_finalize(_fake_partials(y), H * W)
    for act, fn in ((ops.ACT_LRELU02, lambda t: F.leaky_relu(t, 0.2)), (ops.ACT_RELU, F.relu), (ops.ACT_NONE, lambda t: t)):
        got = ops.in_backward_pad(da, y, scale, shift, 1, act)
        yr = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        fn(F.instance_norm(yr)).backward(da.float().permute(0, 3, 1, 2))
        ref = F.pad(yr.grad, (1, 1, 1, 1))
        err = (got.float().permute(0, 3, 1, 2) - ref).abs().max().item()
        assert err < 4e-3 * max(1.0, ref.abs().max().item()), (act, err, ref.abs().max().item())


@pytest.mark.parametrize("mode", ["reflect", "zero"])
def test_in_backward_folded_matches_autograd_through_the_padding(ops, mode):
    """ducosy_in_backward_pad_folded: the gradient w.r.t. pad1(act(IN(y))) goes in, the padding adjoint (reference
    nn.ReflectionPad2d(1) in front of modules/model.py:60-62, or zero padding) is folded while loading.  Against autograd
    through F.pad + instance_norm + relu, and bit-identical interior / near-identical frame against the two-step form."""
    dtype = torch.float16
    B, H, W, Cn = 2, 16, 32, 256
    y = (_rand((B, H, W, Cn), 191) * 1.5 + 0.8).to(dtype).cuda()
    dpad = _rand((B, H + 2, W + 2, Cn), 192).to(dtype).cuda()
    scale, shift = ops.in_finalize(_fake_partials(y), H * W)
    pm = ops.PAD_REFLECT if mode == "reflect" else ops.PAD_ZERO
    got = ops.in_backward_pad_folded(dpad.clone(), pm, y, scale, shift, 2, ops.ACT_RELU)   # the padded map is consumed (folded in place)
    yr = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    a = F.relu(F.instance_norm(yr))
    ap = F.pad(a, (1, 1, 1, 1), mode="reflect") if mode == "reflect" else F.pad(a, (1, 1, 1, 1))
    ap.backward(dpad.float().permute(0, 3, 1, 2))
    ref = F.pad(yr.grad, (2, 2, 2, 2))
    # relu'(n) is discontinuous at n = 0 and the kernel's statistics (fp32 sums of the 16-bit map) differ from F.instance_norm's in
    # the last bits: the handful of elements with |n| < 1e-3 may take either branch -- they are left out of the comparison
    away = F.pad((F.instance_norm(y.float().permute(0, 3, 1, 2)).abs() > 1e-3).float(), (2, 2, 2, 2), value=1.0)
    assert away.mean().item() > 0.99
    err = ((got.float().permute(0, 3, 1, 2) - ref).abs() * away).max().item()
    assert err < 4e-3 * max(1.0, ref.abs().max().item()), (mode, err, ref.abs().max().item())
    two_step = ops.in_backward_pad(ops.pad_fold(dpad, 1, pm), y, scale, shift, 2, ops.ACT_RELU)
    assert (got.float() - two_step.float()).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(2, 16, 16, 256, 512, 4), (1, 32, 64, 64, 128, 4), (2, 32, 32, 128, 256, 3), (1, 16, 64, 64, 128, 3)])
def test_stride2_dgrad_matches_autograd(ops, case):
    B, Ho, Wo, Cin, Cout, ksz = case
    dtype = torch.float16
    dy = _rand((B, Cout, Ho, Wo), 93).to(dtype)
    w = (_rand((Cout, Cin, ksz, ksz), 94, 0.05)).to(dtype).float()
    dyp = F.pad(dy.float(), (1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dx = ops.convs2_dgrad_nhwc(dyp, w.cuda())
    x0 = torch.zeros((B, Cin, 2 * Ho, 2 * Wo), device="cuda", requires_grad=True)
    F.conv2d(x0, w.cuda(), stride=2, padding=1).backward(dy.float().cuda())
    ref = x0.grad
    err = (dx.float().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-2, (err, ref.abs().max().item())


def test_conv2d_wgrad_with_padded_dy(ops):
    dtype = torch.float16
    B, H, W, Cin, Cout = 2, 32, 32, 256, 512
    x = _rand((B, Cin, H, W), 95).to(dtype)
    dy = _rand((B, Cout, H // 2, W // 2), 96).to(dtype)
    xp = F.pad(x.float(), (1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dyp = F.pad(dy.float(), (1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dw = ops.conv2d_wgrad_nhwc(xp, dyp, 4, 4, 2, dy_pad=1)
    w0 = torch.zeros((Cout, Cin, 4, 4), device="cuda", requires_grad=True)
    F.conv2d(x.float().cuda(), w0, stride=2, padding=1).backward(dy.float().cuda())
    ref = w0.grad.permute(0, 2, 3, 1).reshape(Cout, 16 * Cin)
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-2, (err, ref.abs().max().item())


def test_soft_squeeze_window_and_display_windowing(ops):
    px = orc.synthetic_volume(2, 64, 96, seed=101)
    d = torch.from_numpy(px).cuda()
    for lo, hi in ((-150, 250), (-1000, -150)):
        got = ops.hu_window_soft(d, 1.0, -1024.0, lo, hi).cpu().numpy()
        ref = orc.soft_squeeze_window(px, 1.0, -1024.0, lo, hi)
        assert np.abs(got - ref).max() <= 2.4e-7          # same fp32 steps; exp() may differ by one ulp of the result
        below = ref < 2 * 0.9 - 1 - 1e-6                   # the linear part of the curve is bit-exact
        assert np.array_equal(got[below], ref[below].astype(np.float32))
    y = torch.from_numpy(np.random.Generator(np.random.PCG64(5)).uniform(-1, 1, size=(2, 1, 64, 64)).astype(np.float32))
    for lo, hi, wc, ww in ((-150, 250, 40, 400), (-1000, -150, -600, 1500)):
        got = ops.apply_windowing(y.cuda(), lo, hi, wc, ww).cpu().numpy()
        assert np.array_equal(got, orc.apply_windowing(y.numpy(), lo, hi, wc, ww))


@pytest.mark.parametrize("mode", ["reflect", "zero"])
@pytest.mark.parametrize("case", [(2, 32, 128, 256, 256), (1, 16, 128, 128, 64), (1, 8, 256, 64, 128), (1, 16, 64, 256, 256)])
def test_conv3x3s1_dgrad_with_pad_fold(ops, case, mode):
    B, H, W, Cin, Cout = case
    dtype = torch.float16
    dy = _rand((B, Cout, H, W), 111).to(dtype)
    w = _rand((Cout, Cin, 3, 3), 112, 0.05).to(dtype).float()
    dyp = F.pad(dy.float(), (2, 2, 2, 2)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dx, dxpad = ops.conv3x3s1_dgrad(dyp, w.cuda(), ops.PAD_REFLECT if mode == "reflect" else ops.PAD_ZERO)
    x0 = torch.zeros((B, Cin, H, W), device="cuda", requires_grad=True)
    xp = F.pad(x0, (1, 1, 1, 1), mode="reflect" if mode == "reflect" else "constant")
    xp.retain_grad()
    F.conv2d(xp, w.cuda()).backward(dy.float().cuda())
    ref_pad, ref = xp.grad, x0.grad
    tol = 2e-3 * ref_pad.abs().max().item() + 1e-2
    assert (dxpad.float().permute(0, 3, 1, 2) - ref_pad).abs().max().item() <= tol       # incl. the two CUDA-core columns
    assert (dx.float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 2 * tol


@pytest.mark.parametrize("case", [(2, 32, 32, 256, 128), (1, 16, 64, 128, 128)])
def test_upconv2x_backward_matches_autograd(ops, case):
    B, Hs, Ws, Cin, Cout = case
    dtype = torch.float16
    x = _rand((B, Cin, Hs, Ws), 121).to(dtype)
    dy = _rand((B, Cout, 2 * Hs, 2 * Ws), 122).to(dtype)
    w = _rand((Cout, Cin, 3, 3), 123, 0.05)
    xp = F.pad(x.float(), (1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dyp = F.pad(dy.float(), (2, 2, 2, 2)).permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dsrc, dw = ops.upconv2x_backward(xp, dyp, w.cuda())
    xr = x.float().cuda().requires_grad_(True)
    wr = w.cuda().requires_grad_(True)
    F.conv2d(F.interpolate(xr, scale_factor=2, mode="nearest"), wr, padding=1).backward(dy.float().cuda())
    assert (dsrc.float().permute(0, 3, 1, 2) - xr.grad).abs().max().item() <= 4e-3 * xr.grad.abs().max().item() + 1e-2
    assert (dw - wr.grad).abs().max().item() <= 2e-3 * wr.grad.abs().max().item() + 1e-2


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride", [(3, 32, 128, 256, 256, 1), (2, 64, 64, 64, 128, 2), (1, 128, 128, 128, 64, 1)])
def test_conv_with_fused_instance_norm_finalize_matches_separate_finalize(B, H, W, Cin, Cout, stride, monkeypatch):
    """ducosy_conv2d_nhwc_in (the InstanceNorm finalize done by the CTA that completes a sample's last tile; opt-in via
    DUCOSY_FUSED_FINALIZE=1) against conv + ducosy_in_finalize: identical conv output, statistics equal to fp32 rounding of the
    two summation orders, ticket array left zero (re-launchable), deterministic."""
    from ducosy_gan_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + Cout)
    x = (torch.randn(B, H + 2, W + 2, Cin, generator=g)).to(torch.float16).cuda()
    w = ops.pack_conv_weight((torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).cuda(), torch.float16)
    monkeypatch.setenv("DUCOSY_FUSED_FINALIZE", "0")
    y0, (s0, h0, m0) = ops.conv2d_nhwc_in(x, w, 3, 3, stride, want_chmax=True)
    monkeypatch.setenv("DUCOSY_FUSED_FINALIZE", "1")
    for _ in range(2):                                  # second launch: the tickets reset themselves
        y1, (s1, h1, m1) = ops.conv2d_nhwc_in(x, w, 3, 3, stride, want_chmax=True)
        assert torch.equal(y0, y1)
        for a, b in ((s0, s1), (h0, h1), (m0, m1)):
            assert (a - b).abs().max().item() <= 2e-6 * b.abs().max().item() + 1e-7
    y2, (s2, h2, m2) = ops.conv2d_nhwc_in(x, w, 3, 3, stride, want_chmax=True)
    assert torch.equal(s1, s2) and torch.equal(h1, h2) and torch.equal(m1, m2)
    assert int(ops._tickets(x.device, B).abs().sum()) == 0
