"""Image-quality metrics of whole volumes on the GPU: the reference's ``calculate.py`` metric functions
with the same names, arguments and return values -- ``(mean, per-slice list)`` -- evaluated by ``csrc/metrics.cu``.

    normalize(data)                calculate.py:232-238
    calculate_mae(img1, img2)      calculate.py:243-245
    calculate_psnr(img1, img2)     calculate.py:247-263
    calculate_ssim(img1, img2)     calculate.py:265-272  (skimage.metrics.structural_similarity defaults, data_range of img2)
    calculate_cs(img1, img2)       calculate.py:360-367  (sklearn cosine_similarity of the flattened slices)
    calculate_ed(img1, img2)       calculate.py:369-381
    calculate_emd(img1, img2)      calculate.py:320-337  (scipy wasserstein_distance of the jointly normalised slices; int16
                                   volumes: cumulative histograms, integer-exact; float volumes: sorted samples)
    calculate_ts(img1, img2)       calculate.py:340-358  (skimage.filters.sobel magnitudes; parity unpinned, scikit-image absent)

Inputs are [S,H,W] volumes: numpy arrays (copied to the current CUDA device) or CUDA tensors, int16 (the stored pixel arrays
calculate.py:226-228 saves), float32 or float64.  Arithmetic is float64 like numpy's; int16 volumes reproduce numpy's int16
wrap-around in ``img1 - img2`` / ``(img1 - img2) ** 2`` (a reference quirk: the raw-HU MAE / PSNR rows are computed that way).
Not built: MS-SSIM / LPIPS (third-party networks, absent here).
``volume_metrics`` evaluates everything in one pass over the pair.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

_IN = {torch.int16: 0, torch.float32: 1, torch.float64: 2}


def _vol(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.is_tensor(x):
        raise TypeError("expected a numpy array or a torch tensor")
    if x.dtype not in _IN:
        raise TypeError(f"metrics: dtype {x.dtype} unsupported (int16, float32, float64)")
    if x.dim() == 2:
        x = x[None]
    if x.dim() != 3:
        raise ValueError(f"metrics: expected [S,H,W], got {tuple(x.shape)}")
    if not x.is_cuda:
        x = x.cuda()
    _lib.check(_lib.load().ducosy_check_device(), "check_device")
    return x.contiguous()


def _pair(a, b):
    a, b = _vol(a), _vol(b)
    if a.shape != b.shape or a.dtype != b.dtype or a.device != b.device:
        raise ValueError("metrics: the two volumes must have the same shape, dtype and device")
    return a, b


def _scratch(a):
    S, H, W = a.shape
    lib = _lib.load()
    n = max(lib.ducosy_metrics_chunks(H * W) * 12, lib.ducosy_metrics_ssim_tiles(H, W), 1)
    return torch.empty(S * n, dtype=torch.float64, device=a.device)


def _stats(a, b):
    """[S,12] float64 per-slice sums / extrema (see include/ducosy.h), on the host."""
    S, H, W = a.shape
    with torch.cuda.device(a.device):
        stats = torch.empty((S, 12), dtype=torch.float64, device=a.device)
        call("ducosy_metrics_slice_stats", ptr(a), ptr(b), _IN[a.dtype], S, H * W, ptr(stats), ptr(_scratch(a)), stream_ptr())
    return stats


def _wrap16(v, dtype):
    """numpy evaluates ``x.max() - x.min()`` of an int16 array in int16."""
    if dtype != torch.int16:
        return v
    return float(np.array(v, dtype=np.int64).astype(np.int16))


def normalize(data):
    """calculate.py:232-238 -> float64 CUDA tensor [S,H,W]."""
    x = _vol(data)
    with torch.cuda.device(x.device):
        st = _stats(x, x)
        mm = torch.stack([st[:, 7].min(), st[:, 8].max()])
        if x.dtype == torch.int16:       # `data - min_val` and `max_val - min_val` are int16 operations in numpy
            lo, hi = float(mm[0]), float(mm[1])
            if hi - lo > 32767:
                raise OverflowError("normalize: int16 range overflow (numpy would wrap here)")
        out = torch.empty(x.shape, dtype=torch.float64, device=x.device)
        call("ducosy_metrics_normalize", ptr(x), _IN[x.dtype], ptr(out), x.numel(), ptr(mm), stream_ptr())
    return out


def _mae(st, n):
    s = st[:, 0]
    return float(s.sum() / (n * len(s))), (s / n).tolist()


def _psnr(st, n, dtype):
    S = st.shape[0]
    sq = st[:, 1]
    mse = float(sq.sum()) / (n * S)
    if mse == 0:
        return float("inf"), [float("inf")] * S
    rng = _wrap16(float(st[:, 8].max()) - float(st[:, 7].min()), dtype)
    max_pixel = 1.0 if rng == 0 else rng
    psnr = 20 * np.log10(max_pixel / np.sqrt(mse))
    lst = [float("inf") if m == 0 else float(20 * np.log10(max_pixel / np.sqrt(m))) for m in (sq / n).tolist()]
    return float(psnr), lst


def _cs(st):
    # sklearn: normalize(X) . normalize(Y)^T with the norm of an all-zero row replaced by 1
    na, nb = st[:, 3].sqrt(), st[:, 4].sqrt()
    na = torch.where(na == 0, torch.ones_like(na), na)
    nb = torch.where(nb == 0, torch.ones_like(nb), nb)
    v = st[:, 2] / (na * nb)
    return float(v.mean()), v.tolist()


def _ed(a, b, st_dev):
    S, H, W = a.shape
    with torch.cuda.device(a.device):
        sums = torch.empty(S, dtype=torch.float64, device=a.device)
        call("ducosy_metrics_ed", ptr(a), ptr(b), _IN[a.dtype], S, H * W, ptr(st_dev), ptr(sums), ptr(_scratch(a)), stream_ptr())
    v = sums.sqrt().cpu() / (H * W)
    return float(v.mean()), v.tolist()


def _ssim(a, b, st):
    S, H, W = a.shape
    data_range = _wrap16(float(st[:, 10].max()) - float(st[:, 9].min()), a.dtype)
    with torch.cuda.device(a.device):
        sums = torch.empty(S, dtype=torch.float64, device=a.device)
        call("ducosy_metrics_ssim", ptr(a), ptr(b), _IN[a.dtype], S, H, W, float(data_range), ptr(sums), ptr(_scratch(a)), stream_ptr())
    v = sums.cpu() / ((H - 6) * (W - 6))
    return float(v.mean()), v.tolist()


def calculate_mae(img1, img2):
    a, b = _pair(img1, img2)
    return _mae(_stats(a, b).cpu(), a.shape[1] * a.shape[2])


def calculate_psnr(img1, img2):
    a, b = _pair(img1, img2)
    return _psnr(_stats(a, b).cpu(), a.shape[1] * a.shape[2], a.dtype)


def calculate_ssim(img1, img2):
    a, b = _pair(img1, img2)
    return _ssim(a, b, _stats(a, b).cpu())


def calculate_cs(img1, img2):
    a, b = _pair(img1, img2)
    return _cs(_stats(a, b).cpu())


def calculate_ed(img1, img2):
    a, b = _pair(img1, img2)
    return _ed(a, b, _stats(a, b))


def calculate_emd(img1, img2):
    a, b = _pair(img1, img2)
    S, H, W = a.shape
    n = H * W
    st = _stats(a, b).cpu()
    gmin = min(float(st[:, 7].min()), float(st[:, 9].min()))
    gmax = max(float(st[:, 8].max()), float(st[:, 10].max()))
    rng = _wrap16(gmax - gmin, a.dtype) + 1e-8
    with torch.cuda.device(a.device):
        if a.dtype == torch.int16:
            R = int(gmax) - int(gmin) + 1
            hist = torch.zeros(S * 2 * R, dtype=torch.int32, device=a.device)
            sums = torch.empty(S, dtype=torch.float64, device=a.device)
            call("ducosy_metrics_emd_i16", ptr(a), ptr(b), S, n, int(gmin), R, ptr(hist), ptr(sums), stream_ptr())
            d = sums.cpu() / n / rng
        else:
            # float volumes: W1 of two equal-size samples = mean |sorted difference| (the sort is torch's library sort; this
            # branch is not on any measured path)
            sa = torch.sort(a.reshape(S, -1).to(torch.float64), dim=1).values
            sb = torch.sort(b.reshape(S, -1).to(torch.float64), dim=1).values
            d = ((sa - sb).abs().mean(dim=1) / rng).cpu()
    v = d / n          # "scale by number of pixels" (calculate.py:334-335)
    return float(v.mean()), v.tolist()


def calculate_ts(img1, img2):
    a, b = _pair(img1, img2)
    S, H, W = a.shape
    with torch.cuda.device(a.device):
        stats = torch.empty((S, 3), dtype=torch.float64, device=a.device)
        scratch = torch.empty(S * ((H + 7) // 8) * 3, dtype=torch.float64, device=a.device)
        call("ducosy_metrics_ts", ptr(a), ptr(b), _IN[a.dtype], S, H, W, ptr(stats), ptr(scratch), stream_ptr())
    st = stats.cpu()
    diff = st[:, 0] / (H * W)
    mx = torch.maximum(st[:, 1], st[:, 2])
    v = 1.0 - torch.where(mx > 0, diff / torch.where(mx > 0, mx, torch.ones_like(mx)), torch.zeros_like(mx))
    return float(v.mean()), v.tolist()


def volume_metrics(target, pred):
    """What process_single_patient (calculate.py:383-470) computes for one (target, prediction) pair, minus the metrics listed
    as not built: raw and normalised MAE / PSNR / SSIM plus CS, ED, EMD and TS; every entry is ``(mean, per-slice list)``."""
    a, b = _pair(target, pred)
    n = a.shape[1] * a.shape[2]
    out = {}
    st_dev = _stats(a, b)
    st = st_dev.cpu()
    out["mae"], out["psnr"], out["ssim"] = _mae(st, n), _psnr(st, n, a.dtype), _ssim(a, b, st)
    out["cs"], out["ed"] = _cs(st), _ed(a, b, st_dev)
    out["emd"], out["ts"] = calculate_emd(a, b), calculate_ts(a, b)
    an, bn = normalize(a), normalize(b)
    stn = _stats(an, bn).cpu()
    out["mae_norm"], out["psnr_norm"], out["ssim_norm"] = _mae(stn, n), _psnr(stn, n, an.dtype), _ssim(an, bn, stn)
    return out


__all__ = ["normalize", "calculate_mae", "calculate_psnr", "calculate_ssim", "calculate_cs", "calculate_ed", "calculate_emd",
           "calculate_ts", "volume_metrics"]
