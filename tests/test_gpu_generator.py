"""Generator / dual-HU path parity on the B200 (pytest -m gpu): CUDA product through the drop-in
modules/model.py API and the C ABI, checked against the oracle and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ducosy_oracle as orc  # noqa: E402

# Stated tolerance (tanh units) of the 16-bit-operand / fp32-accumulate path against the fp32 oracle.
# 1 HU = 0.005 (soft-tissue window, 400 HU span) or 0.00235 (lung window, 850 HU span).
#   fp16 operands (default): max <= 0.015 (3 HU soft window), mean <= 0.002 (0.4 HU); measured on B200 at 512x512,
#     9 CBAM blocks: max 8.5e-3 (1.7 HU), mean 1.0e-3 (0.2 HU)  -- profiles/r01_accuracy.json
#   bf16 operands: max <= 0.10 (20 HU), mean <= 0.015 (3 HU); measured max 6.5e-2, mean 8.5e-3
# The same figures are produced by a fp32 CPU oracle whose stored activations are rounded to the operand type
# (oracle.generator_forward_rounded), i.e. the error is operand quantisation noise, not a defect.
TOL_TANH = {"fp16": 1.5e-2, "bf16": 1.0e-1}
TOL_MEAN = {"fp16": 2e-3, "bf16": 1.5e-2}


def _x(seed, shape):
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).uniform(-1, 1, size=shape).astype(np.float32))


def _gen(cin, nb, cbam, sd):
    from ducosy_gan_b200.modules.model import Generator
    G = Generator(cin, nb, cbam)
    G.load_state_dict(sd, strict=True)
    return G.cuda().eval()


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_generator_matches_golden_128(golden_dir, precision, monkeypatch):
    monkeypatch.setenv("DUCOSY_PRECISION", precision)
    g = np.load(os.path.join(golden_dir, "gen_c2_b1_cbam_128.npz"))
    sd = orc.make_state_dict(orc.generator_param_shapes(2, 1, True), int(g["wseed"]), attn_std=float(g["attn_std"]))
    G = _gen(2, 1, True, sd)
    with torch.no_grad():
        y = G(_x(int(g["xseed"]), (1, 2, 128, 128)).cuda()).cpu().numpy()
    err = np.abs(y - g["y"]).max()
    print(f"gen 128 golden [{precision}]: max abs err {err:.3e}")
    assert err < TOL_TANH[precision]


@pytest.mark.parametrize("cin,nb,cbam,B,H,W", [(1, 2, True, 2, 128, 128), (3, 1, False, 1, 128, 256), (1, 0, True, 1, 256, 128)])
def test_generator_matches_oracle_small(cin, nb, cbam, B, H, W):
    sd = orc.make_state_dict(orc.generator_param_shapes(cin, nb, cbam), 77, attn_std=0.2)
    G = _gen(cin, nb, cbam, sd)
    x = _x(5, (B, cin, H, W))
    with torch.no_grad():
        y = G(x.cuda()).cpu()
        ref = orc.generator_forward(sd, x, nb, cbam)
    err = (y - ref).abs().max().item()
    print(f"gen cin={cin} nb={nb} cbam={cbam} {H}x{W}: max abs err {err:.3e}")
    assert err < TOL_TANH["fp16"]


def test_generator_full_size_vs_oracle_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "gen_full_512.npz"))
    sd = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), int(g["wseed"]), attn_std=float(g["attn_std"]))
    G = _gen(1, 9, True, sd)
    px = orc.synthetic_volume(1, 512, 512, seed=int(g["vseed"]))
    x = torch.from_numpy(orc.hu_window(px[0], 1.0, -1024.0, *orc.SOFT_HU).astype(np.float32))[None, None]
    with torch.no_grad():
        y = G(x.cuda())
        y_hu = G.forward_hu(torch.from_numpy(px).cuda(), 1.0, -1024.0, *orc.SOFT_HU)
        ref = orc.generator_forward(sd, x)
    assert torch.equal(y, y_hu)                       # fused-window entry == windowed tensor entry
    y = y.cpu()
    err = (y - ref).abs().max().item()
    err_hu = orc.hu_error(y.numpy(), ref.numpy(), *orc.SOFT_HU)
    mean_err = float((y - ref).abs().mean())
    print(f"gen 512 full: max abs err {err:.3e} tanh = {err_hu:.2f} HU (soft window); mean abs {mean_err:.2e}")
    assert mean_err < TOL_MEAN["fp16"]
    # the error must be no worse than the quantisation noise of a rounding-aware fp32 oracle
    with torch.no_grad():
        rref = orc.generator_forward_rounded(sd, x, 9, True, torch.float16)
    noise = float((rref - ref).abs().mean())
    print(f"   rounding-aware oracle vs fp32 oracle: mean abs {noise:.2e}")
    assert mean_err < 1.5 * noise + 1e-4
    assert np.abs(y[0, 0].numpy()[::8, ::8] - g["y_sub"]).max() < TOL_TANH["fp16"]
    assert err < TOL_TANH["fp16"]


def test_generator_config1_soft_tissue_with_mask_channels():
    """BASELINE configs[0]: soft-tissue Generator A2B (input_channels = 3: slice + bone / mediastinum masks, 9 CBAM blocks),
    batch 1, one 512x512 slice built as SURVEY 8d prescribes (windowed phantom slice, bone-candidate mask, Bernoulli(0.1)
    mask) against the fp32 oracle: the stated fp16 tolerance, 3 HU max / 0.4 HU mean in the soft-tissue window."""
    sd = orc.make_state_dict(orc.generator_param_shapes(3, 9, True), 1234)
    G = _gen(3, 9, True, sd)
    px = orc.phantom_volume(1, 512, 512, seed=0)[0]
    hu = orc.stored_to_hu(px, 1.0, -1024.0)
    ct = torch.from_numpy(orc.hu_window(px, 1.0, -1024.0, -150, 250).astype(np.float32))
    bone = torch.from_numpy(((hu >= 200) & (hu > -1000)).astype(np.float32))
    med = (torch.rand(512, 512, generator=torch.Generator().manual_seed(0)) < 0.1).float()
    x = torch.stack([ct, bone, med])[None]
    with torch.no_grad():
        y = G(x.cuda()).cpu()
        ref = orc.generator_forward(sd, x, 9, True)
    err, mean = (y - ref).abs().max().item(), (y - ref).abs().mean().item()
    print(f"config 1 (Cin=3, batch 1, 512x512): max {err:.3e} tanh units = {err * 200:.2f} HU, mean {mean * 200:.3f} HU")
    assert err < TOL_TANH["fp16"] and mean < TOL_MEAN["fp16"]


# Split-operand arm (DUCOSY_F16X2, Generator.precision = "fp16x2"): (hi, lo) fp16 pairs, three tensor-core products per tap,
# fp32 accumulate.  Stated tolerance against the fp32 oracle: max <= 2e-4 tanh units (0.04 HU soft window, 0.085 HU lung window),
# i.e. fp32-class; after the truncating de-window cast (preprocess.py:111) that is "<= 1 HU" (a value sitting on an integer
# boundary may fall either way).
TOL_SPLIT = 2e-4


@pytest.mark.parametrize("cin,nb,cbam,B,H,W", [(1, 2, True, 2, 128, 128), (3, 1, False, 1, 128, 256), (2, 1, True, 1, 256, 128)])
def test_generator_split_operand_small(cin, nb, cbam, B, H, W):
    sd = orc.make_state_dict(orc.generator_param_shapes(cin, nb, cbam), 77, attn_std=0.2)
    G = _gen(cin, nb, cbam, sd)
    G.precision = "fp16x2"
    x = _x(5, (B, cin, H, W))
    with torch.no_grad():
        y = G(x.cuda()).cpu()
        ref = orc.generator_forward(sd, x, nb, cbam)
    err = (y - ref).abs().max().item()
    print(f"gen fp16x2 cin={cin} nb={nb} cbam={cbam} {H}x{W}: max abs err {err:.3e}")
    assert err < TOL_SPLIT


def test_generator_split_operand_full_size_within_1hu(golden_dir):
    """The <= 1 HU arm north_star asks for: 512x512, 9 CBAM blocks, both HU windows' worth of error against the fp32 oracle and
    the reference's own golden output, then through the truncating de-window: |stored-value difference| <= 1."""
    g = np.load(os.path.join(golden_dir, "gen_full_512.npz"))
    sd = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), int(g["wseed"]), attn_std=float(g["attn_std"]))
    G = _gen(1, 9, True, sd)
    G.precision = "fp16x2"
    px = orc.synthetic_volume(1, 512, 512, seed=int(g["vseed"]))
    x = torch.from_numpy(orc.hu_window(px[0], 1.0, -1024.0, *orc.SOFT_HU).astype(np.float32))[None, None]
    with torch.no_grad():
        y = G(x.cuda())
        y_hu = G.forward_hu(torch.from_numpy(px).cuda(), 1.0, -1024.0, *orc.SOFT_HU)
        ref = orc.generator_forward(sd, x)
    assert torch.equal(y, y_hu)
    y = y.cpu()
    err = (y - ref).abs().max().item()
    print(f"gen 512 full fp16x2: max abs err {err:.3e} tanh = {err * 200:.4f} HU (soft) / {err * 425:.4f} HU (lung); "
          f"mean {float((y - ref).abs().mean()):.2e}")
    assert err < TOL_SPLIT
    assert np.abs(y[0, 0].numpy()[::8, ::8] - g["y_sub"]).max() < TOL_SPLIT
    for lo, hi in (orc.SOFT_HU, orc.LUNG_HU):
        a = orc.dewindow_to_stored(y[0, 0].numpy(), 1.0, -1024.0, lo, hi)
        b = orc.dewindow_to_stored(ref[0, 0].numpy(), 1.0, -1024.0, lo, hi)
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        print(f"   de-windowed ({lo},{hi}): max |diff| {d.max()} stored units, {100.0 * (d > 0).mean():.3f} % of voxels differ")
        assert d.max() <= 1


def test_generator_is_deterministic_and_batch_invariant():
    sd = orc.make_state_dict(orc.generator_param_shapes(1, 2, True), 5, attn_std=0.2)
    G = _gen(1, 2, True, sd)
    x = _x(9, (3, 1, 128, 128)).cuda()
    with torch.no_grad():
        a, b = G(x), G(x)
        single = G(x[1:2].contiguous())
    assert torch.equal(a, b)
    assert torch.equal(a[1:2], single)               # InstanceNorm is per sample: no cross-sample coupling


def test_load_state_dict_invalidates_packed_weights():
    shapes = orc.generator_param_shapes(1, 1, True)
    G = _gen(1, 1, True, orc.make_state_dict(shapes, 1))
    x = _x(3, (1, 1, 128, 128)).cuda()
    with torch.no_grad():
        y1 = G(x)
        G.load_state_dict(orc.make_state_dict(shapes, 2))
        y2 = G(x)
    assert not torch.equal(y1, y2)


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (1, 512, 512), (1, 256, 512)])
def test_discriminator_forward_matches_oracle(B, H, W):
    from ducosy_gan_b200.modules.model import Discriminator
    sd = orc.make_state_dict(orc.discriminator_param_shapes(1), 21)
    D = Discriminator(1)
    D.load_state_dict(sd, strict=True)
    D = D.cuda().eval()
    x = _x(31, (B, 1, H, W))
    with torch.no_grad():
        y = D(x.cuda()).cpu()
        ref = orc.discriminator_forward(sd, x)
    assert y.shape == ref.shape == (B, 1, H // 16, W // 16)
    err = (y - ref).abs().max().item()
    print(f"disc {B}x{H}x{W}: max abs err {err:.3e} (ref abs max {ref.abs().max().item():.3f})")
    assert err < 1.5e-2 * max(1.0, ref.abs().max().item())
    # MSE-GAN validation loss (reference modules/trainer.py:243,347) from the patch map
    mse = ((y - 1) ** 2).mean().item()
    assert abs(mse - orc.mse_gan_loss(ref, True).item()) < 2e-2


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (1, 512, 512)])
def test_discriminator_backward_matches_autograd(B, H, W):
    """BASELINE config 3: PatchGAN forward/backward with the MSE adversarial loss (reference trainer.py:347,518-524):
    parameter and input gradients of the CUDA path vs torch autograd through the fp32 oracle."""
    from ducosy_gan_b200.modules.model import Discriminator
    sd = orc.make_state_dict(orc.discriminator_param_shapes(1), 21)
    D = Discriminator(1)
    D.load_state_dict(sd, strict=True)
    D = D.cuda().train()
    x = _x(41, (B, 1, H, W))
    xg = x.clone().cuda().requires_grad_(True)
    loss = torch.nn.functional.mse_loss(D(xg), torch.ones(B, 1, H // 16, W // 16, device="cuda"))
    loss.backward()
    # oracle + autograd (fp32, CPU)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    ref_loss = orc.mse_gan_loss(orc.discriminator_forward(ref_sd, xr), True)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * max(1.0, ref_loss.item())
    for name, p in D.named_parameters():
        g, r = p.grad.cpu(), ref_sd[name].grad
        if name in ("model.2.bias", "model.5.bias", "model.8.bias"):
            assert g.abs().max().item() == 0.0 and r.abs().max().item() < 1e-4      # dead under the non-affine InstanceNorm
            continue
        l2 = ((g - r).norm() / (r.norm() + 1e-20)).item()
        mx = (g - r).abs().max().item() / (r.abs().max().item() + 1e-12)
        print(f"  {name}: rel L2 err {l2:.3e}, max err / max |ref| {mx:.3e}")
        # stated tolerance (fp16 operands, 16-bit stored activations and gradients): 5 % relative L2, 20 % of the largest
        # reference entry.  Every building block is checked at 2e-3 in test_gpu_kernels.py; what is left is the
        # amplification of 16-bit rounding through three InstanceNorms (LeakyReLU sign flips at |n| ~ 0, few pixels per
        # weight in the deepest layer) -- it grows 8-10x with bf16 operands (tools/disc_grad_check.py).
        assert l2 < 5e-2 and mx < 0.2, (name, l2, mx)
    gx, rx = xg.grad.cpu(), xr.grad
    l2x = ((gx - rx).norm() / rx.norm()).item()
    print(f"  input grad: rel L2 err {l2x:.3e}")
    assert l2x < 5e-2


def test_reinit_through_data_invalidates_packed_weights():
    """the reference's weights_init_normal writes through ``.data`` (modules/model.py:134-140), which does not bump the
    parameters' ``_version``: ``.apply(fn)`` must invalidate the packed 16-bit weights whatever ``fn`` does"""
    from ducosy_gan_b200.modules.model import Discriminator, Generator, weights_init_normal

    def data_init(m):                                  # the reference's exact form
        if "Conv" in type(m).__name__:
            torch.nn.init.normal_(m.weight.data, 0.0, 0.02)

    x = _x(3, (1, 1, 128, 128)).cuda()
    for net in (Generator(1, 1, True).cuda().eval(), Discriminator(1).cuda().eval()):
        xin = x if isinstance(net, Generator) else _x(3, (1, 1, 256, 256)).cuda()
        with torch.no_grad():
            torch.manual_seed(1)
            net.apply(weights_init_normal)
            y1 = net(xin).clone()
            torch.manual_seed(2)
            net.apply(data_init)
            y2 = net(xin).clone()
            for p in net.parameters():                 # hand edit through .data + the documented explicit hook
                p.data.mul_(0.5)
            net.invalidate_packed_weights()
            y3 = net(xin)
        assert not torch.equal(y1, y2) and not torch.equal(y2, y3)


def test_modules_deepcopy_and_pickle_after_forward():
    """EMA copies (copy.deepcopy) and whole-module checkpoints (torch.save(model)) work on the reference's modules; the
    engines (device buffers, captured graphs) are derived state and must not get in the way"""
    import copy
    import io
    from ducosy_gan_b200.modules.model import Discriminator, Generator
    sd = orc.make_state_dict(orc.generator_param_shapes(1, 1, True), 5, attn_std=0.2)
    G = _gen(1, 1, True, sd)
    D = Discriminator(1).cuda().eval()
    x = _x(3, (1, 1, 256, 256)).cuda()
    with torch.no_grad():
        y, d = G(x), D(x)
        G2, D2 = copy.deepcopy(G), copy.deepcopy(D)
        buf = io.BytesIO()
        torch.save(G, buf)
        buf.seek(0)
        G3 = torch.load(buf, weights_only=False)
        assert torch.equal(G2(x), y) and torch.equal(G3(x), y) and torch.equal(D2(x), d)


def test_batch1_forward_replays_a_cuda_graph_bit_identical_to_eager(monkeypatch):
    """generate.py:96-97 calls model(x) slice by slice; the eval-mode small-batch call is captured once per shape and
    replayed.  Same kernels, same buffers layout: the bytes must equal the eager call's, for fresh inputs, after a weight
    change (the graph reads the re-packed weights), for a second shape, and the returned tensors must not alias."""
    shapes = orc.generator_param_shapes(1, 2, True)
    G = _gen(1, 2, True, orc.make_state_dict(shapes, 3, attn_std=0.2))
    xs = [_x(s, (1, 1, 256, 256)).cuda() for s in (1, 2, 3)]
    with torch.no_grad():
        monkeypatch.setenv("DUCOSY_FORWARD_GRAPH", "0")
        eager = [G(x).clone() for x in xs]
        monkeypatch.setenv("DUCOSY_FORWARD_GRAPH", "1")
        got = [G(x) for x in xs]                      # first call captures, the others replay
        eng = next(iter(G._engines.values()))
        assert len(eng._graphs) == 1
        assert all(torch.equal(a, b) for a, b in zip(eager, got))
        assert got[0].data_ptr() != got[1].data_ptr() and not torch.equal(got[0], got[1])
        two = torch.cat(xs[:2])
        assert torch.equal(G(two), torch.cat(eager[:2]))          # another shape: its own graph, batch invariant
        assert len(eng._graphs) == 2
        G.load_state_dict(orc.make_state_dict(shapes, 4, attn_std=0.2))
        after = G(xs[0])
        monkeypatch.setenv("DUCOSY_FORWARD_GRAPH", "0")
        assert torch.equal(after, G(xs[0])) and not torch.equal(after, eager[0])
        G.train()
        monkeypatch.setenv("DUCOSY_FORWARD_GRAPH", "1")
        assert torch.equal(G(xs[0]), after)           # train mode: eager launches, same result (no dropout / running stats)


# ------------------------------------------------------------------ stand-alone building blocks (reference modules/model.py:6-87)
def test_channel_and_spatial_attention_modules_match_oracle():
    """ChannelAttention / SpatialAttention / CBAM called on their own (NCHW fp32, both pooling branches, inputs that are NOT
    behind an InstanceNorm so the avg branch matters): fp32 kernels, 1e-5 against the oracle."""
    from ducosy_gan_b200.modules.model import CBAM, ChannelAttention, SpatialAttention
    torch.manual_seed(0)
    for C, r, k, B, H, W in [(256, 16, 7, 2, 32, 48), (64, 8, 3, 1, 17, 23), (48, 16, 7, 3, 8, 8)]:
        x = torch.randn(B, C, H, W) * 1.5 + 0.7
        ca, sa, cb = ChannelAttention(C, r).cuda(), SpatialAttention(k).cuda(), CBAM(C, r, k).cuda()
        with torch.no_grad():
            for m in (ca, sa, cb):
                for p in m.parameters():
                    p.normal_(0, 0.2)
            y_ca, y_sa, y_cb = ca(x.cuda()).cpu(), sa(x.cuda()).cpu(), cb(x.cuda()).cpu()
            cpu = lambda t: t.detach().cpu()
            r_ca = orc.channel_attention(x, cpu(ca.fc[0].weight), cpu(ca.fc[2].weight))
            r_sa = orc.spatial_attention(x, cpu(sa.conv.weight))
            r_cb = orc.spatial_attention(orc.channel_attention(x, cpu(cb.channel_attention.fc[0].weight), cpu(cb.channel_attention.fc[2].weight)),
                                         cpu(cb.spatial_attention.conv.weight))
        for got, ref, name in ((y_ca, r_ca, "channel"), (y_sa, r_sa, "spatial"), (y_cb, r_cb, "cbam")):
            err = (got - ref).abs().max().item() / ref.abs().max().item()
            assert err < 1e-5, (name, C, err)
    with pytest.raises(RuntimeError):
        ca(x.cuda().requires_grad_(True))            # forward-only on their own: loud, not silent


@pytest.mark.parametrize("cbam", [True, False])
def test_residual_block_modules_match_oracle(cbam):
    """BASELINE config 5 call: ResidualBlockWithCBAM(256)(randn(B,256,128,128)) (and the plain ResidualBlock) through the
    tensor-core path, against oracle.residual_block: 16-bit operands, stated tolerance of the generator tests."""
    from ducosy_gan_b200.modules.model import ResidualBlock, ResidualBlockWithCBAM
    torch.manual_seed(3)
    blk = (ResidualBlockWithCBAM if cbam else ResidualBlock)(256).cuda()
    with torch.no_grad():
        for n, p in blk.named_parameters():
            p.normal_(0, 0.2 if "cbam" in n else 0.02)
    sd = {f"m.{k}": v.detach().cpu() for k, v in blk.state_dict().items()}
    for B, H, W in [(2, 128, 128), (1, 32, 64)]:
        x = torch.randn(B, 256, H, W)
        with torch.no_grad():
            y = blk(x.cuda()).cpu()
            ref = orc.residual_block(x, sd, "m", cbam)
        err = (y - ref).abs().max().item()
        print(f"residual block cbam={cbam} {B}x256x{H}x{W}: max abs err {err:.3e} (|ref| max {ref.abs().max().item():.2f})")
        assert y.shape == ref.shape and err < 1.5e-2
