"""Generator backward parity (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI; the reference
values are torch autograd in fp32 on the same 16-bit-rounded operands (tolerance stated per test)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False      # the fp32 autograd reference must not run on TF32 tensor cores
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops():
    from ducosy_gan_b200 import _lib, ops as _ops
    _lib.check(_lib.load().ducosy_check_device(), "check_device")
    return _ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def _rel(got, ref):
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 64, 128)])
def test_out_conv_backward_matches_autograd(ops, shape):
    """modules/model.py:112-113.  16-bit activations, fp32 everything else: 2e-3 relative (da is rounded to 16 bit)."""
    B, H, W = shape
    dtype = torch.float16
    a = F.relu(_rand((B, 64, H, W), 201)).to(dtype)
    w = _rand((1, 64, 7, 7), 202, 0.02)
    bias = _rand((1,), 203, 0.1)
    dout = _rand((B, 1, H, W), 204) * 3e-6          # mean-reduced losses give gradients of this size
    ar = a.float().cuda().requires_grad_(True)
    wr = w.cuda().requires_grad_(True)
    br = bias.cuda().requires_grad_(True)
    out = torch.tanh(F.conv2d(F.pad(ar, (3, 3, 3, 3), mode="reflect"), wr, br))
    out.backward(dout.cuda())
    in_pad = F.pad(a.float(), (3, 3, 3, 3), mode="reflect").permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    gs = ops.grad_scale(dout.cuda())
    da, dw, db = ops.out_conv_backward(dout.cuda(), out.detach(), in_pad, w.cuda(), gs)
    got_da = da.float().permute(0, 3, 1, 2) * gs[1]
    assert _rel(got_da, ar.grad) < 2e-3, _rel(got_da, ar.grad)
    assert _rel(dw, wr.grad) < 3e-4, _rel(dw, wr.grad)
    assert abs(db.item() - br.grad.item()) <= 1e-4 * abs(br.grad.item()) + 1e-12


@pytest.mark.parametrize("Cin", [1, 3])
def test_stem_backward_matches_autograd(ops, Cin):
    """modules/model.py:90-91: weight gradient through the saved im2col matrix, image gradient (channel 0) through the
    col2im adjoint.  Operands are 16-bit: 3e-3 relative L2."""
    B, H, W = 2, 32, 128
    dtype = torch.float16
    x = _rand((B, Cin, H, W), 211, 0.5).clamp(-1, 1).to(dtype).float()
    w = _rand((64, Cin, 7, 7), 212, 0.02).to(dtype).float()
    dy = _rand((B, 64, H, W), 213).to(dtype)
    xr = x.cuda().requires_grad_(True)
    wr = w.cuda().requires_grad_(True)
    F.conv2d(F.pad(xr, (3, 3, 3, 3), mode="reflect"), wr).backward(dy.float().cuda() * 2.0 ** -10)
    cols = ops.stem_im2col(x.cuda(), dtype)
    gs = torch.tensor([2.0 ** 10, 2.0 ** -10], device="cuda")
    dw, dx = ops.stem_backward(dy.permute(0, 2, 3, 1).contiguous().cuda(), cols, w.cuda(), gs)
    assert _rel(dw, wr.grad) < 3e-3, _rel(dw, wr.grad)
    assert _rel(dx[:, 0], xr.grad[:, 0]) < 3e-3, _rel(dx[:, 0], xr.grad[:, 0])
