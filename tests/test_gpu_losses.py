"""Loss terms of the CycleGAN step on the CUDA path: value and gradient parity against the
oracle restatement (pinned to the reference's own classes by tests/golden/losses.npz) evaluated with torch autograd."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ducosy_oracle as orc  # noqa: E402


def _x(seed, shape):
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).uniform(-1, 1, size=shape).astype(np.float32))


def _smooth(seed, shape):
    """smoother images (blurred noise): closer to CT slices than white noise, keeps |.| kinks away from zero"""
    t = _x(seed, shape)
    return torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0


def _check(cuda_fn, ref_fn, inputs, tol_val=2e-5, tol_grad=2e-4):
    p = inputs[0].clone().cuda().requires_grad_(True)
    rest = [t.cuda() for t in inputs[1:]]
    out = cuda_fn(p, *rest)
    (out * 1.7).backward()                       # non-trivial upstream gradient
    pr = inputs[0].clone().requires_grad_(True)
    ref = ref_fn(pr, *inputs[1:])
    (ref * 1.7).backward()
    assert abs(out.item() - ref.item()) <= tol_val * max(1.0, abs(ref.item())), (out.item(), ref.item())
    g, r = p.grad.cpu(), pr.grad
    err = (g - r).abs().max().item()
    assert err <= tol_grad * r.abs().max().item() + 1e-9, (err, r.abs().max().item())
    return out.item(), ref.item()


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (8, 1, 512, 512)])
def test_l1_and_mse_gan(shape):
    from ducosy_gan_b200 import losses
    a, b = _x(1, shape), _x(2, shape)
    _check(losses.l1_loss, torch.nn.functional.l1_loss, [a, b])
    d = _x(3, (shape[0], 1, shape[2] // 16, shape[3] // 16)) * 3
    _check(lambda t: losses.mse_gan_loss(t, True), lambda t: orc.mse_gan_loss(t, True), [d])
    _check(lambda t: losses.mse_gan_loss(t, False), lambda t: orc.mse_gan_loss(t, False), [d])


def test_losses_match_golden_values(golden_dir):
    """the exact tensors of tests/golden/losses.npz (values produced by the reference's own classes)"""
    from ducosy_gan_b200 import losses
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    p, t, s = (_x(int(k), (2, 1, 64, 64)).cuda() for k in g["seeds"])
    assert abs(losses.GradientLoss()(p, t).item() - float(g["grad"])) < 2e-6
    assert abs(losses.ContrastAttentionLoss(0.15, 1.0, 3.0, 7)(p, t, s).item() - float(g["att"])) < 2e-6
    assert abs(losses.ContrastRegionLoss(0.15, 1.5)(p, t, s).item() - float(g["region"])) < 2e-6
    assert abs(losses.ContrastEdgeLoss()(p, t, s).item() - float(g["edge"])) < 2e-6


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (4, 1, 256, 512)])
def test_gradient_loss(shape):
    from ducosy_gan_b200 import losses
    _check(losses.GradientLoss(), orc.gradient_loss, [_smooth(11, shape), _smooth(12, shape)])


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (4, 1, 256, 512)])
def test_contrast_attention_loss(shape):
    from ducosy_gan_b200 import losses
    mod = losses.ContrastAttentionLoss(sigma=0.15, min_weight=1.0, max_weight=3.0, blur_kernel=7)
    _check(mod, orc.contrast_attention_loss, [_smooth(21, shape), _smooth(22, shape), _smooth(23, shape)])


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (4, 1, 256, 512)])
def test_contrast_region_loss(shape):
    from ducosy_gan_b200 import losses
    mod = losses.ContrastRegionLoss(threshold=0.15, weight=1.5)
    _check(mod, orc.contrast_region_loss, [_smooth(31, shape), _smooth(32, shape), _smooth(33, shape)], tol_grad=5e-4)


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (8, 1, 512, 512)])
def test_contrast_edge_loss(shape):
    from ducosy_gan_b200 import losses
    _check(losses.ContrastEdgeLoss(), lambda p, t: orc.contrast_edge_loss(p, t), [_smooth(41, shape), _smooth(42, shape)],
           tol_val=5e-5, tol_grad=1e-3)


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (2, 1, 256, 512)])
def test_ssim_unpinned_restatement(shape):
    """PARITY UNPINNED (pytorch_msssim is not part of the reference tree): checked against the oracle restatement."""
    from ducosy_gan_b200 import losses
    mod = losses.SSIM(data_range=1.0, size_average=True, channel=1)
    _check(mod, orc.ssim, [_smooth(51, shape), _smooth(52, shape)], tol_val=2e-5, tol_grad=2e-3)
    _check(lambda a, b: 1 - mod(a, b), lambda a, b: 1 - orc.ssim(a, b), [_smooth(53, shape), _smooth(51, shape) * 0.8 + 0.1],
           tol_val=2e-5, tol_grad=2e-3)


def test_loss_mix_matches_reference_weights():
    """trainer.py:493-512 mix on the CUDA losses vs the oracle (weights: GAN 1, cyc 10, id 5, grad_cyc 5, grad_id 2.5,
    ssim 2, att 2, region 1.5, edge 1)."""
    from ducosy_gan_b200 import losses
    shape = (2, 1, 128, 128)
    fake, rec, idt, real_a, real_b = (_smooth(60 + i, shape) for i in range(5))
    d_out = _x(70, (2, 1, 8, 8))

    def mix(L, d, f, r, i, a, b, ssim):
        return (L["gan"](d) + 10 * L["l1"](r, a) + 5 * L["l1"](i, b) + 5 * L["grad"](r, a) + 2.5 * L["grad"](i, b) +
                2 * (1 - ssim(r, a)) + 2 * L["att"](f, b, a) + 1.5 * L["region"](f, b, a) + 1.0 * L["edge"](f, b, a))

    cu = {"gan": lambda d: losses.mse_gan_loss(d, True), "l1": losses.l1_loss, "grad": losses.GradientLoss(),
          "att": losses.ContrastAttentionLoss(0.15, 1.0, 3.0, 7), "region": losses.ContrastRegionLoss(0.15, 1.5),
          "edge": losses.ContrastEdgeLoss()}
    rf = {"gan": lambda d: orc.mse_gan_loss(d, True), "l1": torch.nn.functional.l1_loss, "grad": orc.gradient_loss,
          "att": orc.contrast_attention_loss, "region": orc.contrast_region_loss,
          "edge": lambda p, t, s: orc.contrast_edge_loss(p, t)}
    c = [t.cuda() for t in (d_out, fake, rec, idt, real_a, real_b)]
    got = mix(cu, *c, losses.SSIM(1.0)).item()
    ref = mix(rf, d_out, fake, rec, idt, real_a, real_b, orc.ssim).item()
    assert abs(got - ref) < 1e-4 * max(1.0, abs(ref))
