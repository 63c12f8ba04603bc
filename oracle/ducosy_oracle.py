"""CPU oracle for the DuCoSy-GAN hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (plain torch fp32 functional ops + numpy) of the
reference algorithm for the dual-HU synthesis path.  It exists so that the CUDA
product in ``ducosy_gan_b200/`` can be checked; it is never part of the product
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity status
-------------
* generator / discriminator / windowing / de-windowing / composite / HU
  threshold candidates: PINNED against the reference itself.  ``oracle/make_golden.py``
  imports ``/root/reference/modules/{model,preprocess}.py`` (and executes the
  composite statements of ``generate.py:140-145,218-237``) in the authoring
  container and commits small input/output vectors under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` replays them against this file.
* SSIM (``ssim``): "parity unpinned" -- the reference delegates to the third
  party ``pytorch_msssim`` package (unpinned in ``requirements.txt:28``, not
  vendored, not installed here).  The restatement follows the package's
  published algorithm (gaussian 11/1.5, valid conv, C1/C2 from data_range).

* volume smoothing (``postprocess_volume``), anatomical masks (``mask_*``), the CycleGAN step (``cyclegan_step``) and the
  image-quality metrics (``metric_mae / psnr / cs / ed / normalize``): PINNED by ``oracle/make_golden_postprocess.py``,
  ``make_golden_masks.py``, ``make_golden_trainstep.py`` and ``make_golden_metrics.py``, which run the reference's own
  functions (``modules/postprocess.py``, ``modules/mask_generator.py``, the loop body of ``modules/trainer.py:448-525``,
  ``calculate.py:232-263,360-381``) on seeded inputs.
* ``skimage_structural_similarity`` (the core of ``calculate_ssim``): "parity unpinned" -- scikit-image is absent and
  unpinned; restated with the ``scipy.ndimage.uniform_filter`` it calls.

Every function cites the reference ``file:line`` it follows.
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

# HU ranges of the two CycleGANs (modules/argmanager.py:121-152).
SOFT_HU = (-150.0, 250.0)
LUNG_HU = (-1000.0, -150.0)


# --------------------------------------------------------------------------------------
# deterministic, RNG-library-independent weights (no shipped checkpoints: SURVEY 8c)
# --------------------------------------------------------------------------------------
def _param_rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def generator_param_shapes(input_channels=1, num_residual_blocks=9, use_cbam=True):
    """state_dict key -> shape, in the reference's order (modules/model.py:92-113)."""
    shapes = {}
    shapes["model.1.weight"] = (64, input_channels, 7, 7)
    shapes["model.1.bias"] = (64,)
    shapes["model.4.weight"] = (128, 64, 3, 3)
    shapes["model.4.bias"] = (128,)
    shapes["model.7.weight"] = (256, 128, 3, 3)
    shapes["model.7.bias"] = (256,)
    idx = 10
    for _ in range(num_residual_blocks):
        for j in (1, 5):
            shapes[f"model.{idx}.block.{j}.weight"] = (256, 256, 3, 3)
            shapes[f"model.{idx}.block.{j}.bias"] = (256,)
        if use_cbam:
            shapes[f"model.{idx}.cbam.channel_attention.fc.0.weight"] = (16, 256, 1, 1)
            shapes[f"model.{idx}.cbam.channel_attention.fc.2.weight"] = (256, 16, 1, 1)
            shapes[f"model.{idx}.cbam.spatial_attention.conv.weight"] = (1, 2, 7, 7)
        idx += 1
    shapes[f"model.{idx + 1}.weight"] = (128, 256, 3, 3)
    shapes[f"model.{idx + 1}.bias"] = (128,)
    shapes[f"model.{idx + 5}.weight"] = (64, 128, 3, 3)
    shapes[f"model.{idx + 5}.bias"] = (64,)
    shapes[f"model.{idx + 9}.weight"] = (1, 64, 7, 7)
    shapes[f"model.{idx + 9}.bias"] = (1,)
    return shapes


def discriminator_param_shapes(input_channels=1):
    """modules/model.py:120-129."""
    chans = [(64, input_channels), (128, 64), (256, 128), (512, 256), (1, 512)]
    shapes = {}
    for key, (co, ci) in zip((0, 2, 5, 8, 12), chans):
        shapes[f"model.{key}.weight"] = (co, ci, 4, 4)
        shapes[f"model.{key}.bias"] = (co,)
    return shapes


def make_state_dict(shapes: dict, seed: int, weight_std: float = 0.02, attn_std: float | None = None):
    """Weights ~ N(0, weight_std) as weights_init_normal does (modules/model.py:134-140); biases
    uniform in +-0.05.  Drawn with numpy PCG64 keyed by (seed, key) so the same tensors can be
    rebuilt anywhere without shipping a checkpoint.  ``attn_std`` optionally gives the CBAM convs a
    larger spread so the attention maps are not all ~0.5 (exercises the sigmoid properly)."""
    sd = {}
    for name, shape in shapes.items():
        rng = _param_rng(seed, name)
        if name.endswith("bias"):
            a = rng.uniform(-0.05, 0.05, size=shape)
        else:
            std = weight_std
            if attn_std is not None and "cbam" in name:
                std = attn_std
            a = rng.standard_normal(size=shape) * std
        sd[name] = torch.from_numpy(a.astype(np.float32))
    return sd


# --------------------------------------------------------------------------------------
# model restatement (modules/model.py)
# --------------------------------------------------------------------------------------
def instance_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm2d defaults: biased variance, eps 1e-5, no affine, no running stats
    (modules/model.py:61,75,79,94,97,110)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def channel_attention(x, w_fc0, w_fc2):
    """modules/model.py:20-24."""
    avg = x.mean(dim=(2, 3), keepdim=True)
    mx = x.amax(dim=(2, 3), keepdim=True)

    def fc(v):
        return F.conv2d(F.relu(F.conv2d(v, w_fc0)), w_fc2)

    return x * torch.sigmoid(fc(avg) + fc(mx))


def spatial_attention(x, w_conv):
    """modules/model.py:34-39."""
    avg = x.mean(dim=1, keepdim=True)
    mx = x.amax(dim=1, keepdim=True)
    att = torch.sigmoid(F.conv2d(torch.cat([avg, mx], dim=1), w_conv, padding=w_conv.shape[-1] // 2))
    return x * att


def residual_block(x, sd, prefix, use_cbam=True):
    """modules/model.py:56-65 (plain) and 68-87 (CBAM)."""
    h = F.pad(x, (1, 1, 1, 1), mode="reflect")
    h = F.conv2d(h, sd[f"{prefix}.block.1.weight"], sd[f"{prefix}.block.1.bias"])
    h = F.relu(instance_norm(h))
    h = F.pad(h, (1, 1, 1, 1), mode="reflect")
    h = F.conv2d(h, sd[f"{prefix}.block.5.weight"], sd[f"{prefix}.block.5.bias"])
    h = instance_norm(h)
    if use_cbam:
        h = channel_attention(h, sd[f"{prefix}.cbam.channel_attention.fc.0.weight"],
                              sd[f"{prefix}.cbam.channel_attention.fc.2.weight"])
        h = spatial_attention(h, sd[f"{prefix}.cbam.spatial_attention.conv.weight"])
    return x + h


def generator_forward(sd, x, num_residual_blocks=9, use_cbam=True, return_intermediates=False):
    """Generator.forward (modules/model.py:92-115).  x: [B,Cin,H,W] fp32."""
    inter = {}
    h = F.pad(x, (3, 3, 3, 3), mode="reflect")
    h = F.conv2d(h, sd["model.1.weight"], sd["model.1.bias"])
    h = F.relu(instance_norm(h))
    inter["stem"] = h
    h = F.relu(instance_norm(F.conv2d(h, sd["model.4.weight"], sd["model.4.bias"], stride=2, padding=1)))
    inter["down1"] = h
    h = F.relu(instance_norm(F.conv2d(h, sd["model.7.weight"], sd["model.7.bias"], stride=2, padding=1)))
    inter["down2"] = h
    idx = 10
    for _ in range(num_residual_blocks):
        h = residual_block(h, sd, f"model.{idx}", use_cbam)
        inter[f"block{idx}"] = h
        idx += 1
    for _ in range(2):
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        h = F.conv2d(h, sd[f"model.{idx + 1}.weight"], sd[f"model.{idx + 1}.bias"], padding=1)
        h = F.relu(instance_norm(h))
        inter[f"up{idx}"] = h
        idx += 4
    h = F.pad(h, (3, 3, 3, 3), mode="reflect")
    h = torch.tanh(F.conv2d(h, sd[f"model.{idx + 1}.weight"], sd[f"model.{idx + 1}.bias"]))
    if return_intermediates:
        return h, inter
    return h


def discriminator_forward(sd, x):
    """Discriminator.forward (modules/model.py:120-131)."""
    h = F.leaky_relu(F.conv2d(x, sd["model.0.weight"], sd["model.0.bias"], stride=2, padding=1), 0.2)
    for key in (2, 5, 8):
        h = F.conv2d(h, sd[f"model.{key}.weight"], sd[f"model.{key}.bias"], stride=2, padding=1)
        h = F.leaky_relu(instance_norm(h), 0.2)
    h = F.pad(h, (1, 0, 1, 0))
    return F.conv2d(h, sd["model.12.weight"], sd["model.12.bias"], padding=1)


# --------------------------------------------------------------------------------------
# HU windowing / de-windowing / composite / threshold candidates (numpy, float32 like the reference)
# --------------------------------------------------------------------------------------
def stored_to_hu(px: np.ndarray, slope: float, intercept: float) -> np.ndarray:
    """modules/preprocess.py:72-75 and generate.py:140-145: float32 px * slope + intercept."""
    image = px.astype(np.float32)
    return image * float(slope) + float(intercept)


def hu_window(px, slope, intercept, hu_min, hu_max) -> np.ndarray:
    """Linear inference windowing, modules/preprocess.py:72-84:  clip then 2*(x-lo)/(hi-lo)-1."""
    image = stored_to_hu(px, slope, intercept)
    image = np.clip(image, hu_min, hu_max)
    return 2 * (image - hu_min) / (hu_max - hu_min) - 1


def soft_squeeze_window(px, slope, intercept, hu_min, hu_max, sigma=50):
    """Training-side windowing, modules/preprocess.py:6-40,43-55 (apply_hu_transform with soft squeezing)."""
    image = stored_to_hu(px, slope, intercept)
    image = np.clip(image, hu_min, hu_max)
    normalized = (image - hu_min) / (hu_max - hu_min)
    threshold = 0.9
    k = 10.0 / sigma
    soft_mask = 1.0 / (1.0 + np.exp(-k * (normalized - threshold)))
    result = np.where(normalized < threshold, normalized, threshold + (1.0 - threshold) * soft_mask)
    return 2.0 * result - 1.0


def dewindow_to_stored(y: np.ndarray, slope, intercept, hu_min, hu_max, dtype=np.int16) -> np.ndarray:
    """postprocess_tensor, modules/preprocess.py:96-111: (y+1)/2*(hi-lo)+lo -> (hu-b)/m -> astype (truncation)."""
    y = np.asarray(y, dtype=np.float32)
    denorm = (y + 1.0) / 2.0 * (hu_max - hu_min) + hu_min
    new_px = (denorm - float(intercept)) / float(slope)
    return new_px.astype(dtype)


def apply_windowing(y, hu_min, hu_max, window_center, window_width):
    """apply_windowing, modules/preprocess.py:58-65 (display windowing of a tanh-range tensor, used for the validation
    JPEGs at modules/trainer.py:276-278): de-window, clamp to [wc - ww/2, wc + ww/2], scale to [0, 1]."""
    t = torch.as_tensor(y, dtype=torch.float32)
    hu = (t + 1.0) / 2.0 * (hu_max - hu_min) + hu_min
    lo, hi = window_center - window_width / 2.0, window_center + window_width / 2.0
    return ((torch.clamp(hu, lo, hi) - lo) / window_width).numpy()


def composite(raw_px, soft_px, lung_px, slope, intercept, soft_hu=SOFT_HU, lung_hu=LUNG_HU):
    """generate.py:213-237: start from the NCCT stored values, overwrite the soft-tissue HU range
    with the soft-tissue generator's pixels, then the lung range with the lung generator's
    (so lung wins where the ranges touch, HU == -150).  Returns (merged, soft_mask, lung_mask)."""
    merged = raw_px.copy()
    hu = stored_to_hu(raw_px, slope, intercept)
    soft_mask = np.logical_and(hu >= soft_hu[0], hu <= soft_hu[1])
    lung_mask = np.logical_and(hu >= lung_hu[0], hu <= lung_hu[1])
    merged[soft_mask] = soft_px[soft_mask]
    merged[lung_mask] = lung_px[lung_mask]
    return merged, soft_mask, lung_mask


def threshold_candidates(hu: np.ndarray):
    """HU threshold candidates of the anatomical mask generator:
    body hu>-1000, lung -1000<=hu<=-300 & body (mask_generator.py:14-20), bone hu>=200 & body
    (mask_generator.py:179-183).  Returns uint8 (body, lung, bone)."""
    body = (hu > -1000).astype(np.uint8)
    lung = np.logical_and(np.logical_and(hu >= -1000, hu <= -300).astype(np.uint8), body).astype(np.uint8)
    bone = np.logical_and((hu >= 200).astype(np.uint8), body).astype(np.uint8)
    return body, lung, bone


def dual_hu_synthesize(raw_px, slope, intercept, sd_soft, sd_lung, num_residual_blocks=9, use_cbam=True,
                       return_parts=False):
    """The whole north-star path for a volume [S,H,W] of stored values:
    generate.py:89-102 (window -> 2 generators -> de-window) then generate.py:213-237 (composite)."""
    S = raw_px.shape[0]
    merged = np.empty_like(raw_px)
    ys, yl = [], []
    with torch.no_grad():
        for i in range(S):
            xs = torch.from_numpy(hu_window(raw_px[i], slope, intercept, *SOFT_HU).astype(np.float32))[None, None]
            xl = torch.from_numpy(hu_window(raw_px[i], slope, intercept, *LUNG_HU).astype(np.float32))[None, None]
            y_soft = generator_forward(sd_soft, xs, num_residual_blocks, use_cbam)[0, 0].numpy()
            y_lung = generator_forward(sd_lung, xl, num_residual_blocks, use_cbam)[0, 0].numpy()
            soft_px = dewindow_to_stored(y_soft, slope, intercept, *SOFT_HU, dtype=raw_px.dtype)
            lung_px = dewindow_to_stored(y_lung, slope, intercept, *LUNG_HU, dtype=raw_px.dtype)
            merged[i] = composite(raw_px[i], soft_px, lung_px, slope, intercept)[0]
            ys.append(y_soft)
            yl.append(y_lung)
    if return_parts:
        return merged, np.stack(ys), np.stack(yl)
    return merged


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synthetic_volume(S: int, H: int = 512, W: int = 512, seed: int = 0) -> np.ndarray:
    """Stored pixels uniform in [0,2500) (slope 1, intercept -1024 => HU in [-1024,1475])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 2500, size=(S, H, W), dtype=np.int16)


def phantom_volume(S: int, H: int = 512, W: int = 512, seed: int = 0) -> np.ndarray:
    """Structured chest phantom: air, two lung ellipses, body ellipse, spine disc (stored values, intercept -1024)."""
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    cy, cx = H / 2, W / 2
    vol = np.empty((S, H, W), np.float32)
    for s in range(S):
        hu = np.full((H, W), -1000.0, np.float32)
        body = ((yy - cy) / (0.36 * H)) ** 2 + ((xx - cx) / (0.44 * W)) ** 2 <= 1
        hu[body] = 40 + 30 * rng.standard_normal(int(body.sum())).astype(np.float32)
        for sgn in (-1, 1):
            lung = ((yy - cy) / (0.22 * H)) ** 2 + ((xx - (cx + sgn * 0.2 * W)) / (0.14 * W)) ** 2 <= 1
            hu[lung] = -800 + 60 * rng.standard_normal(int(lung.sum())).astype(np.float32)
        spine = (yy - (cy + 0.22 * H)) ** 2 + (xx - cx) ** 2 <= (0.05 * H) ** 2
        hu[spine] = 400 + 100 * rng.standard_normal(int(spine.sum())).astype(np.float32)
        vol[s] = hu
    return np.clip(np.rint(vol + 1024.0), 0, 4095).astype(np.int16)


# --------------------------------------------------------------------------------------
# losses (modules/trainer.py:22-184) -- restated for the training-step rows
# --------------------------------------------------------------------------------------
def gradient_loss(pred, target):
    """GradientLoss, modules/trainer.py:29-40."""
    pdy = torch.abs(pred[:, :, 1:, :] - pred[:, :, :-1, :])
    pdx = torch.abs(pred[:, :, :, 1:] - pred[:, :, :, :-1])
    tdy = torch.abs(target[:, :, 1:, :] - target[:, :, :-1, :])
    tdx = torch.abs(target[:, :, :, 1:] - target[:, :, :, :-1])
    return torch.mean(torch.abs(pdy - tdy)) + torch.mean(torch.abs(pdx - tdx))


def _gauss_window(size=11, sigma=1.5):
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim(x, y, data_range=1.0, size=11, sigma=1.5):
    """pytorch_msssim.SSIM(data_range=1.0,size_average=True,channel=1) as called at
    modules/trainer.py:351,485.  PARITY UNPINNED (third-party package absent): separable valid
    gaussian filtering, C1=(0.01L)^2, C2=(0.03L)^2, mean over the map then over batch."""
    win = _gauss_window(size, sigma)
    C = x.shape[1]

    def filt(t):
        t = F.conv2d(t, win.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)
        return F.conv2d(t, win.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)

    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = filt(x), filt(y)
    s11 = filt(x * x) - mu1 * mu1
    s22 = filt(y * y) - mu2 * mu2
    s12 = filt(x * y) - mu1 * mu2
    cs = (2 * s12 + c2) / (s11 + s22 + c2)
    ssim_map = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs
    return ssim_map.flatten(2).mean(-1).mean(1).mean()


def contrast_attention_loss(pred, target, source, sigma=0.15, min_w=1.0, max_w=3.0, k=7):
    """ContrastAttentionLoss, modules/trainer.py:62-86."""
    blur = lambda t: F.avg_pool2d(t, k, 1, k // 2)
    pb, tb, sb = blur(pred), blur(target), blur(source)
    diff = torch.abs(tb - sb)
    w = min_w + (max_w - min_w) * (1.0 - torch.exp(-diff / sigma))
    return torch.mean(w * torch.abs(pb - tb))


def mse_gan_loss(d_out, is_real: bool):
    """nn.MSELoss against ones/zeros, modules/trainer.py:347,459-460,470."""
    tgt = torch.ones_like(d_out) if is_real else torch.zeros_like(d_out)
    return F.mse_loss(d_out, tgt)


def hu_error(y_test, y_ref, hu_min, hu_max):
    """max-abs error of a tanh-unit output expressed in HU after de-windowing (SURVEY 8d parity gate)."""
    return float(np.max(np.abs(np.asarray(y_test, np.float64) - np.asarray(y_ref, np.float64)))) * (hu_max - hu_min) / 2.0


def contrast_region_loss(pred, target, source, threshold=0.15, weight=1.5):
    """ContrastRegionLoss, modules/trainer.py:104-130 (ctor values from trainer.py:357)."""
    pool = lambda t: F.avg_pool2d(t, 8, 8)
    pp, tp, sp = pool(pred), pool(target), pool(source)
    mask = torch.sigmoid(5 * ((tp - sp) - threshold))
    region = torch.mean(mask * torch.abs(pp - tp))
    dist = torch.abs(pred.mean() - target.mean()) + torch.abs(pred.std() - target.std())
    return weight * (region + 0.5 * dist)


def sobel_edges(img):
    """ContrastEdgeLoss.get_edges, modules/trainer.py:150-155."""
    sx = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=torch.float32).view(1, 1, 3, 3)
    sy = torch.tensor([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=torch.float32).view(1, 1, 3, 3)
    ex = F.conv2d(img, sx, padding=1)
    ey = F.conv2d(img, sy, padding=1)
    return torch.sqrt(ex ** 2 + ey ** 2 + 1e-6)


def contrast_edge_loss(pred, target, source=None):
    """ContrastEdgeLoss.forward, modules/trainer.py:157-184: |d mean| + |d std| + |d top-10% mean| of Sobel magnitudes."""
    pe, te = sobel_edges(pred), sobel_edges(target)
    stats = torch.abs(pe.mean() - te.mean()) + torch.abs(pe.std() - te.std())
    k = 0.1
    ptop = torch.topk(pe.flatten(), int(pe.numel() * k)).values.mean()
    ttop = torch.topk(te.flatten(), int(te.numel() * k)).values.mean()
    return stats + torch.abs(ptop - ttop)


def generator_forward_rounded(sd, x, nb=9, cbam=True, dt=torch.float16):
    """generator_forward with activations / conv weights rounded to the 16-bit operand type ``dt`` at the points
    where the CUDA path stores them (fp32 arithmetic otherwise).  Its distance to ``generator_forward`` is the
    quantisation noise a 16-bit-operand implementation must show; used to tell that noise from real defects."""
    r = lambda t: t.to(dt).float()
    inorm = instance_norm
    h = F.conv2d(r(F.pad(x, (3, 3, 3, 3), mode="reflect")), r(sd["model.1.weight"]))
    h = r(F.relu(inorm(r(h))))
    h = r(F.relu(inorm(r(F.conv2d(h, r(sd["model.4.weight"]), stride=2, padding=1)))))
    h = r(F.relu(inorm(r(F.conv2d(h, r(sd["model.7.weight"]), stride=2, padding=1)))))
    idx = 10
    for _ in range(nb):
        p = f"model.{idx}"
        t = r(F.conv2d(F.pad(h, (1, 1, 1, 1), mode="reflect"), r(sd[p + ".block.1.weight"])))
        t = r(F.relu(inorm(t)))
        t = r(F.conv2d(F.pad(t, (1, 1, 1, 1), mode="reflect"), r(sd[p + ".block.5.weight"])))
        t = inorm(t)
        if cbam:
            t = channel_attention(t, sd[p + ".cbam.channel_attention.fc.0.weight"], sd[p + ".cbam.channel_attention.fc.2.weight"])
            t = spatial_attention(t, sd[p + ".cbam.spatial_attention.conv.weight"])
        h = r(h + t)
        idx += 1
    for _ in range(2):
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        h = r(F.relu(inorm(r(F.conv2d(h, r(sd[f"model.{idx + 1}.weight"]), padding=1)))))
        idx += 4
    h = F.conv2d(F.pad(h, (3, 3, 3, 3), mode="reflect"), r(sd[f"model.{idx + 1}.weight"]), sd[f"model.{idx + 1}.bias"])
    return torch.tanh(h)


# --------------------------------------------------------------------------------------
# post-composite volume smoothing
# --------------------------------------------------------------------------------------
def postprocess_test_volume(S: int, H: int, W: int, seed: int) -> np.ndarray:
    """Seeded int16 volume for the post-processing tests: noisy soft tissue, a flat air region (exactly constant: the
    case where 1-ulp differences of a float64 blur would flip the int16 truncation), bone >= 750 (kept voxels), and
    slice-to-slice jumps for the z filters."""
    g = np.random.Generator(np.random.PCG64(seed))
    vol = g.normal(1060.0, 40.0, size=(S, H, W))
    vol += (np.arange(S) % 3)[:, None, None] * 25.0                     # stair-steps along z
    vol[:, : H // 4, : W // 3] = 24.0                                    # air, flat
    vol[:, H // 2: H // 2 + 6, W // 2: W // 2 + 9] = g.normal(1900.0, 200.0, size=(S, 6, 9))   # bone
    vol[S // 2:, -5:, :] = 700.0 + g.integers(0, 120, size=(S - S // 2, 5, W))                 # straddles the 750 threshold
    return np.clip(np.rint(vol), -2000, 4000).astype(np.int16)


def postprocess_volume(merged: np.ndarray, pre_sigma_z=0.8, sigma_z=0.7, sigma_xy=0.05, sharpen_amount=1.7,
                       sharpen_radius=1.2, hu_threshold=750) -> np.ndarray:
    """generate.py:254-263 followed through modules/postprocess.py:45-50 (copy + mask), :58-60 (gaussian3d), :99-103
    (unsharp call), :106 (restore), :109 (int16) and unsharp_mask :140-160, with scipy doing the filtering exactly as in
    the reference.  merged: int16 [S,H,W] -> int16 [S,H,W]."""
    from scipy.ndimage import gaussian_filter, gaussian_filter1d
    vol = np.array(merged, dtype=np.float32)                                         # generate.py:255
    vol = gaussian_filter1d(vol, sigma=pre_sigma_z, axis=0)                          # generate.py:258-259
    original = vol.copy()                                                            # postprocess.py:47
    high = vol >= hu_threshold                                                       # postprocess.py:50
    pp = gaussian_filter(vol, sigma=(sigma_z, sigma_xy, sigma_xy))                   # postprocess.py:58-60
    sm, orig = pp.astype(np.float64), original.astype(np.float64)                    # postprocess.py:140-141
    hf = sm - gaussian_filter(sm, sigma=(0, sharpen_radius, sharpen_radius))         # postprocess.py:145-146
    ohf = orig - gaussian_filter(orig, sigma=(0, sharpen_radius, sharpen_radius))    # postprocess.py:149-150
    comb = (1 - sharpen_amount) * hf + sharpen_amount * ohf                          # postprocess.py:153
    sharp = np.clip(sm + comb * sharpen_amount, orig.min(), orig.max())              # postprocess.py:156-159
    sharp[high] = original[high]                                                     # postprocess.py:106
    return sharp.astype(np.int16)                                                    # postprocess.py:109


# --------------------------------------------------------------------------------------
# CycleGAN optimisation step
# --------------------------------------------------------------------------------------
def cyclegan_generator_loss(sdGA, sdGB, sdDA, sdDB, real_A, real_B, masks=None, num_residual_blocks=9, use_cbam=True,
                            lambda_cyc=10.0, lambda_id=5.0):
    """The generator half of the loop body, modules/trainer.py:447-512, on state dicts: returns (loss_G, dict of the nine
    terms, fake_A, fake_B).  Loss weights: trainer.py:493-512 and argmanager.py:97-98."""
    G = lambda sd, x: generator_forward(sd, x, num_residual_blocks, use_cbam)
    cat = (lambda t: torch.cat([t, masks], 1)) if masks is not None else (lambda t: t)
    l1 = F.l1_loss
    fake_B, fake_A = G(sdGA, cat(real_A)), G(sdGB, cat(real_B))                          # trainer.py:464
    id_A, id_B = G(sdGB, cat(real_A)), G(sdGA, cat(real_B))                              # trainer.py:467
    t = {}
    t["id"] = (l1(id_A, real_A) + l1(id_B, real_B)) / 2                                  # trainer.py:469
    t["GAN"] = (mse_gan_loss(discriminator_forward(sdDB, fake_B), True)
                + mse_gan_loss(discriminator_forward(sdDA, fake_A), True)) / 2           # trainer.py:470
    rec_A, rec_B = G(sdGB, cat(fake_B)), G(sdGA, cat(fake_A))                            # trainer.py:474-480
    t["cycle"] = (l1(rec_A, real_A) + l1(rec_B, real_B)) / 2                             # trainer.py:482
    t["grad_cycle"] = (gradient_loss(rec_A, real_A) + gradient_loss(rec_B, real_B)) / 2  # trainer.py:483
    t["grad_id"] = (gradient_loss(id_A, real_A) + gradient_loss(id_B, real_B)) / 2       # trainer.py:484
    t["ssim"] = 1 - (ssim(rec_A, real_A) + ssim(rec_B, real_B)) / 2                      # trainer.py:485
    t["contrast_attention"] = contrast_attention_loss(fake_B, real_B, real_A)            # trainer.py:489
    t["contrast_region"] = contrast_region_loss(fake_B, real_B, real_A)                  # trainer.py:490
    t["contrast_edge"] = contrast_edge_loss(fake_B, real_B, real_A)                      # trainer.py:491
    loss_G = (t["GAN"] + lambda_cyc * t["cycle"] + lambda_id * t["id"] + 5.0 * t["grad_cycle"] + 2.5 * t["grad_id"]
              + 2.0 * t["ssim"] + 2.0 * t["contrast_attention"] + 1.5 * t["contrast_region"] + 1.0 * t["contrast_edge"])
    return loss_G, t, fake_A, fake_B


def cyclegan_discriminator_loss(sdD, real, fake):
    """modules/trainer.py:518 / :523."""
    return (mse_gan_loss(discriminator_forward(sdD, real), True)
            + mse_gan_loss(discriminator_forward(sdD, fake.detach()), False)) / 2


def cyclegan_step(sds, opts, real_A, real_B, masks=None, num_residual_blocks=9, use_cbam=True):
    """One iteration of modules/trainer.py:462-524 with torch autograd and the given optimisers (opt_G, opt_D_A, opt_D_B) over
    the state dicts (sdGA, sdGB, sdDA, sdDB) whose tensors require grad.  Returns the logged losses as floats."""
    sdGA, sdGB, sdDA, sdDB = sds
    oG, oDA, oDB = opts
    oG.zero_grad()
    loss_G, t, fake_A, fake_B = cyclegan_generator_loss(sdGA, sdGB, sdDA, sdDB, real_A, real_B, masks, num_residual_blocks, use_cbam)
    loss_G.backward()
    oG.step()
    oDA.zero_grad()
    loss_DA = cyclegan_discriminator_loss(sdDA, real_A, fake_A)
    loss_DA.backward()
    oDA.step()
    oDB.zero_grad()
    loss_DB = cyclegan_discriminator_loss(sdDB, real_B, fake_B)
    loss_DB.backward()
    oDB.step()
    out = {k: float(v.detach()) for k, v in t.items()}
    out.update(G=float(loss_G.detach()), D_A=float(loss_DA.detach()), D_B=float(loss_DB.detach()))
    return out


# --------------------------------------------------------------------------------------
# anatomical masks (first half) -- modules/mask_generator.py restated with the same scipy calls
# PINNED: oracle/make_golden_masks.py runs the reference's own detect_lung / detect_lung_vessels (matplotlib.path stubbed: those
# two functions never touch it) on mask_test_slices() and stores the masks in tests/golden/masks.npz.
# --------------------------------------------------------------------------------------
def mask_test_slices(B: int, H: int = 256, W: int = 256, seed: int = 0) -> np.ndarray:
    """HU slices (float32, as modules/dataset.py:112-113 builds them) that exercise every branch of the mask code: a chest
    phantom with two lungs, vessel-like discs and a few air pockets inside them (holes for binary_fill_holes), speckle
    components below and above min_size, structures inside the border margin, one slice with a single lung (fails the
    ">= 2 regions" test) and one empty slice."""
    rng = np.random.Generator(np.random.PCG64(seed + 4242))
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    cy, cx = H / 2, W / 2
    out = np.empty((B, H, W), np.float32)
    for b in range(B):
        hu = np.full((H, W), -1000.0, np.float32)
        if b % 5 != 4:                                                # every fifth slice stays empty air
            body = ((yy - cy) / (0.40 * H)) ** 2 + ((xx - cx) / (0.46 * W)) ** 2 <= 1
            hu[body] = 40 + 30 * rng.standard_normal(int(body.sum())).astype(np.float32)
            sides = (-1, 1) if b % 5 != 3 else (1,)                   # slice 3: one lung only
            for sgn in sides:
                lung = ((yy - cy) / (0.24 * H)) ** 2 + ((xx - (cx + sgn * 0.21 * W)) / (0.15 * W)) ** 2 <= 1
                hu[lung] = -800 + 60 * rng.standard_normal(int(lung.sum())).astype(np.float32)
                for _ in range(6):                                    # vessels / nodules and air pockets inside the lung
                    vy = cy + rng.uniform(-0.18, 0.18) * H
                    vx = cx + sgn * 0.21 * W + rng.uniform(-0.09, 0.09) * W
                    r = rng.uniform(1.5, 6.0)
                    disc = (yy - vy) ** 2 + (xx - vx) ** 2 <= r * r
                    hu[disc] = rng.choice(np.array([60.0, 300.0, 700.0, -1000.0, -250.0], np.float32))
            for _ in range(40):                                       # lung-range speckle of 1..120 pixels anywhere (also in the margin)
                sy, sx = int(rng.integers(0, H - 12)), int(rng.integers(0, W - 12))
                hh, ww = int(rng.integers(1, 11)), int(rng.integers(1, 13))
                hu[sy:sy + hh, sx:sx + ww] = -600.0
        out[b] = hu
    return out


def mask_label4(mask: np.ndarray):
    """scipy.ndimage.label with the default structure on every slice of [B,H,W] (mask_generator.py:32): (labels int32, counts)."""
    from scipy import ndimage
    labels = np.zeros(mask.shape, np.int32)
    nums = np.zeros(mask.shape[0], np.int32)
    for z in range(mask.shape[0]):
        labels[z], nums[z] = ndimage.label(mask[z])
    return labels, nums


def mask_fill_holes(mask: np.ndarray) -> np.ndarray:
    """scipy.ndimage.binary_fill_holes on every slice (mask_generator.py:69)."""
    from scipy import ndimage
    return np.stack([ndimage.binary_fill_holes(m).astype(np.uint8) for m in mask])


def mask_detect_lung(hu: np.ndarray, lung_lower=-1000, lung_upper=-300, min_size=64, border_margin=32) -> np.ndarray:
    """modules/mask_generator.py:11-52 (3-D branch: every slice on its own)."""
    from scipy import ndimage
    body_mask = (hu > -1000).astype(np.uint8)                                            # :14
    lung_hu_mask = np.logical_and(hu >= lung_lower, hu <= lung_upper).astype(np.uint8)   # :17
    lung_mask = np.logical_and(lung_hu_mask, body_mask).astype(np.uint8)                 # :20
    _, height, width = lung_mask.shape
    lung_mask[:, :border_margin, :] = 0                                                  # :40-43
    lung_mask[:, height - border_margin:, :] = 0
    lung_mask[:, :, :border_margin] = 0
    lung_mask[:, :, width - border_margin:] = 0
    for z in range(lung_mask.shape[0]):                                                  # :45-50
        labeled, n = ndimage.label(lung_mask[z])
        sizes = np.bincount(labeled.ravel(), minlength=n + 1)
        small = np.flatnonzero(sizes < min_size)
        small = small[small > 0]
        lung_mask[z][np.isin(labeled, small)] = 0
    return lung_mask


def mask_detect_lung_vessels(hu: np.ndarray, lung_mask: np.ndarray, vessel_lower=-300, vessel_upper=600) -> np.ndarray:
    """modules/mask_generator.py:55-99 (3-D branch)."""
    from scipy import ndimage
    body_mask = (hu > -1000).astype(np.uint8)
    vessel = np.zeros_like(lung_mask, dtype=np.uint8)
    for z in range(lung_mask.shape[0]):
        lung_slice, body_slice = lung_mask[z], body_mask[z]
        _, num = ndimage.label(lung_slice)                                               # :85
        body_area, lung_area = body_slice.sum(), lung_slice.sum()
        if num >= 2 and body_area > 0 and (lung_area / body_area) >= 0.1:                # :89
            filled = ndimage.binary_fill_holes(lung_slice).astype(np.uint8)
            cand = filled - lung_slice
        else:
            cand = np.zeros_like(lung_slice)
        cond = np.logical_and(hu[z] >= vessel_lower, hu[z] <= vessel_upper)              # :96
        vessel[z] = np.logical_and(cand, cond).astype(np.uint8)
    return vessel


# --------------------------------------------------------------------------------------
# image-quality metrics (calculate.py) -- SURVEY 8f row N4
# --------------------------------------------------------------------------------------
def metrics_test_volumes(S: int, H: int, W: int, seed: int):
    """A seeded (target, prediction) pair of int16 stored-value volumes: phantom slices and a perturbed copy (noise, a contrast
    shift inside the body, one slice left identical so the inf / zero branches of PSNR are exercised)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tgt = phantom_volume(S, H, W, seed=seed).astype(np.int16)
    noise = rng.normal(0.0, 25.0, size=tgt.shape)
    pred = tgt.astype(np.float64) + noise + 40.0 * (tgt > 900)
    pred[:, : H // 8, : W // 8] += 300.0          # a block of large differences: (a-b)**2 leaves the int16 range
    pred = np.clip(np.rint(pred), -2000, 4000).astype(np.int16)
    if S > 1:
        pred[S - 1] = tgt[S - 1]
    return tgt, pred


def metric_normalize(data):
    """calculate.py:232-238."""
    mn, mx = data.min(), data.max()
    if mx - mn == 0:
        return np.zeros_like(data)
    return (data - mn) / (mx - mn)


def metric_mae(img1, img2):
    """calculate.py:243-245 (numpy dtype semantics kept: int16 inputs subtract in int16)."""
    diff = np.abs(img1 - img2)
    return np.mean(diff), [np.mean(s) for s in diff]


def metric_psnr(img1, img2):
    """calculate.py:247-263."""
    mse = np.mean((img1 - img2) ** 2)
    if mse == 0:
        return float("inf"), [float("inf")] * len(img1)
    rng = img1.max() - img1.min()
    max_pixel = 1.0 if rng == 0 else rng
    out = []
    for s1, s2 in zip(img1, img2):
        m = np.mean((s1 - s2) ** 2)
        out.append(float("inf") if m == 0 else 20 * np.log10(max_pixel / np.sqrt(m)))
    return 20 * np.log10(max_pixel / np.sqrt(mse)), out


def skimage_structural_similarity(im1, im2, data_range):
    """skimage.metrics.structural_similarity with its defaults as calculate.py:270 calls it (win_size 7, uniform filter,
    use_sample_covariance=True, K1 0.01, K2 0.03, crop of (win_size-1)//2, float64 mean).  PARITY UNPINNED: scikit-image is
    not installed here (and unpinned in requirements.txt); restated from its published implementation with the same
    scipy.ndimage.uniform_filter it calls."""
    from scipy.ndimage import uniform_filter
    ft = np.float32 if im1.dtype == np.float32 else np.float64
    im1, im2 = im1.astype(ft, copy=False), im2.astype(ft, copy=False)
    win, NP = 7, 49
    cov_norm = NP / (NP - 1)
    f = lambda a: uniform_filter(a, size=win)
    ux, uy = f(im1), f(im2)
    uxx, uyy, uxy = f(im1 * im1), f(im2 * im2), f(im1 * im2)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    return S[pad:-pad, pad:-pad].mean(dtype=np.float64)


def metric_ssim(img1, img2):
    """calculate.py:265-272."""
    data_range = img2.max() - img2.min()
    out = [skimage_structural_similarity(s1, s2, data_range) for s1, s2 in zip(img1, img2)]
    return np.mean(out), out


def metric_cs(img1, img2):
    """calculate.py:360-367 (sklearn cosine_similarity: rows scaled to unit norm in float64, then the dot product)."""
    out = []
    for s1, s2 in zip(img1, img2):
        v1, v2 = s1.reshape(-1).astype(np.float64), s2.reshape(-1).astype(np.float64)
        n1, n2 = np.sqrt(v1 @ v1), np.sqrt(v2 @ v2)
        out.append(float((v1 / (n1 if n1 else 1.0)) @ (v2 / (n2 if n2 else 1.0))))
    return np.mean(out), out


def metric_ed(img1, img2):
    """calculate.py:369-381."""
    out = []
    for s1, s2 in zip(img1, img2):
        a = (s1 - s1.min()) / (s1.max() - s1.min() + 1e-8)
        b = (s2 - s2.min()) / (s2.max() - s2.min() + 1e-8)
        out.append(np.linalg.norm(a - b) / np.prod(a.shape))
    return np.mean(out), out


def metric_emd(img1, img2):
    """calculate.py:320-337 (scipy.stats.wasserstein_distance is present here: pinned)."""
    from scipy.stats import wasserstein_distance
    gmin, gmax = min(img1.min(), img2.min()), max(img1.max(), img2.max())
    out = []
    for s1, s2 in zip(img1, img2):
        a = (s1 - gmin) / (gmax - gmin + 1e-8)
        b = (s2 - gmin) / (gmax - gmin + 1e-8)
        out.append(wasserstein_distance(a.flatten(), b.flatten()) / np.prod(s1.shape))
    return np.mean(out), out


def skimage_sobel(image):
    """skimage.filters.sobel (>= 0.18, mask=None, mode='reflect') up to the constant factor of its int -> float conversion, which
    calculate_ts's ratio cancels: separable [1,2,1]/4 smoothing x [1,0,-1] difference per axis, sqrt(mean of squares).
    PARITY UNPINNED: scikit-image is absent here."""
    from scipy import ndimage as ndi
    img = image.astype(np.float64)
    sm, ed = np.array([1.0, 2.0, 1.0]) / 4.0, np.array([1.0, 0.0, -1.0])
    g0 = ndi.correlate1d(ndi.correlate1d(img, ed, axis=0, mode="reflect"), sm, axis=1, mode="reflect")
    g1 = ndi.correlate1d(ndi.correlate1d(img, ed, axis=1, mode="reflect"), sm, axis=0, mode="reflect")
    return np.sqrt((g0 * g0 + g1 * g1) / 2.0)


def metric_ts(img1, img2):
    """calculate.py:340-358."""
    out = []
    for s1, s2 in zip(img1, img2):
        g1, g2 = skimage_sobel(s1), skimage_sobel(s2)
        diff = np.mean(np.abs(g1 - g2))
        mx = np.max([np.abs(g1).max(), np.abs(g2).max()])
        out.append(1.0 - (diff / mx if mx > 0 else 0))
    return np.mean(out), out


# --------------------------------------------------------------------------------------
# anatomical masks, second half: convex hull of the lungs (mediastinum / bone masks)
# --------------------------------------------------------------------------------------
def mask_test_slices_bone(B: int, H: int = 256, W: int = 256, seed: int = 0) -> np.ndarray:
    """``mask_test_slices`` plus bone-range structures placed to hit every branch of detect_bone: a spine disc inside the
    preserved bottom quarter, ribs outside the lung hull, a calcification between the lungs (inside the hull: removed), and a
    bridge that connects an inside-hull piece to a rib (restored by the region growing); soft tissue between the lungs for the
    mediastinum mask."""
    rng = np.random.Generator(np.random.PCG64(seed + 977))
    hu = mask_test_slices(B, H, W, seed).copy()
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    cy, cx = H / 2, W / 2
    for b in range(B):
        if b % 5 == 4:
            continue
        disc = lambda y, x, r: (yy - y) ** 2 + (xx - x) ** 2 <= r * r
        hu[b][disc(cy + 0.30 * H, cx, 0.045 * H)] = 500.0                                   # spine (row > 0.75 H)
        hu[b][disc(cy + 0.30 * H, cx, 0.015 * H)] = 30.0                                    # spinal canal: a hole to fill
        for sgn in (-1, 1):
            for k in range(5):                                                              # ribs around the lungs
                ang = np.deg2rad(-60 + 30 * k)
                ry, rx = cy + 0.33 * H * np.sin(ang), cx + sgn * 0.41 * W * np.cos(ang)
                hu[b][disc(ry, rx, 0.012 * H + 1)] = 350.0 + 40 * k
        hu[b][disc(cy - 0.05 * H, cx, 0.02 * H)] = 420.0                                    # calcification between the lungs
        hu[b][disc(cy + 0.10 * H, cx + 0.01 * W, 0.015 * H)] = 260.0
        if b % 2 == 0:                                                                      # bridge: inside-hull bone joined to a rib
            y0 = int(cy - 0.33 * H * np.sin(np.deg2rad(60)))
            x0, x1 = int(cx - 0.41 * W * np.cos(np.deg2rad(60))), int(cx - 0.10 * W)
            hu[b][y0 - 1:y0 + 2, min(x0, x1):max(x0, x1)] = 300.0
        hu[b] += (rng.standard_normal((H, W)) * 0.0).astype(np.float32)
    return hu


def path_contains_points(verts: np.ndarray, points: np.ndarray) -> np.ndarray:
    """``matplotlib.path.Path(verts).contains_points(points)`` (radius 0, no codes: the polyline is closed implicitly), restated
    from matplotlib's ``point_in_path_impl`` (src/_path.h) -- the crossings test with its exact inequality conventions:
        yflag = (vertex_y >= ty);  for an edge whose end points straddle ty:
        if ((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == yflag1: inside ^= 1
    PARITY UNPINNED: matplotlib is absent here; only points exactly on the boundary depend on these conventions."""
    v = np.asarray(verts, dtype=np.float64)
    tx, ty = np.asarray(points, dtype=np.float64).T
    inside = np.zeros(len(tx), dtype=bool)
    n = len(v)
    for i in range(n):
        x0, y0 = v[i]
        x1, y1 = v[(i + 1) % n]
        f0, f1 = y0 >= ty, y1 >= ty
        hit = (f0 != f1) & ((((y1 - ty) * (x0 - x1)) >= ((x1 - tx) * (y0 - y1))) == f1)
        inside ^= hit
    return inside


def mask_convex_hull(lung_slice: np.ndarray):
    """The hull rasterisation shared by detect_mediastinum / detect_bone (mask_generator.py:115-127,204-216): returns
    (hull mask uint8, ok) -- ok False where the reference falls into its ``except`` / ``len(lung_coords) < 3`` branches."""
    from scipy.spatial import ConvexHull
    coords = np.argwhere(lung_slice == 1)
    if len(coords) < 3:
        return lung_slice.copy(), False
    try:
        hull = ConvexHull(coords)
    except Exception:
        return lung_slice.copy(), False
    H, W = lung_slice.shape
    ys, xs = np.mgrid[:H, :W]
    pts = np.vstack((ys.flatten(), xs.flatten())).T
    return path_contains_points(coords[hull.vertices], pts).reshape(H, W).astype(np.uint8), True


def _hull_slice_ok(lung_slice, body_slice):
    """the plausibility test in front of every hull use (mask_generator.py:110-114)"""
    from scipy import ndimage
    _, num = ndimage.label(lung_slice)
    body_area, lung_area = body_slice.sum(), lung_slice.sum()
    return num >= 2 and body_area > 0 and (lung_area / body_area) >= 0.1


def mask_detect_mediastinum(hu: np.ndarray, lung_mask: np.ndarray, mediastinum_lower=-300, mediastinum_upper=450) -> np.ndarray:
    """modules/mask_generator.py:100-170 (3-D branch)."""
    body = (hu > -1000).astype(np.uint8)
    out = np.zeros_like(lung_mask, dtype=np.uint8)
    for z in range(lung_mask.shape[0]):
        if _hull_slice_ok(lung_mask[z], body[z]):
            hull, _ = mask_convex_hull(lung_mask[z])
            cand = hull - lung_mask[z]
            cond = np.logical_and(hu[z] >= mediastinum_lower, hu[z] <= mediastinum_upper)
            out[z] = np.logical_and(cand, cond).astype(np.uint8)
    return out


def mask_detect_bone(hu: np.ndarray, lung_mask: np.ndarray, bone_threshold=200, spine_margin_ratio=0.25) -> np.ndarray:
    """modules/mask_generator.py:173-311 (3-D branch): bone candidates minus the lung hull's interior (outside the lungs, above
    the preserved spine rows), components of the candidates that still touch the remainder restored, holes filled."""
    from scipy import ndimage
    body = (hu > -1000).astype(np.uint8)
    cand = np.logical_and((hu >= bone_threshold).astype(np.uint8), body).astype(np.uint8)
    bone = cand.copy()
    for z in range(lung_mask.shape[0]):
        if _hull_slice_ok(lung_mask[z], body[z]):
            hull, ok = mask_convex_hull(lung_mask[z])
            if ok:
                height = lung_mask[z].shape[0]
                spine = np.zeros_like(lung_mask[z], dtype=np.uint8)
                spine[int(height * (1 - spine_margin_ratio)):, :] = 1
                region = np.logical_and(np.logical_and(hull, 1 - lung_mask[z]), 1 - spine)
                bone[z] = np.logical_and(bone[z], 1 - region).astype(np.uint8)
    for z in range(bone.shape[0]):
        removed = np.logical_and(cand[z], 1 - bone[z]).astype(np.uint8)
        if removed.sum() == 0:
            continue
        combined = np.logical_or(bone[z], removed).astype(np.uint8)
        labeled, _ = ndimage.label(combined)
        keep = set(np.unique(labeled[bone[z] > 0]))
        keep.discard(0)
        for lab in keep:
            bone[z] = np.logical_or(bone[z], np.logical_and(labeled == lab, hu[z] >= bone_threshold)).astype(np.uint8)
    for z in range(bone.shape[0]):
        if bone[z].sum() > 0:
            bone[z] = ndimage.binary_fill_holes(bone[z]).astype(np.uint8)
    return bone
