"""Device version of the volume post-processing generate.py:254-263 applies after the composite (reference
modules/postprocess.py:6-109 ``postprocess_ct_volume(method='gaussian3d', enhance_sharpness=True)`` and :114-160
``unsharp_mask``): bit-exact with the scipy pipeline, on the merged int16 volume while it is still in HBM."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def gaussian_kernel1d(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """The normalised float64 weights scipy.ndimage.gaussian_filter1d builds (order 0): radius int(truncate*sigma + 0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return np.ascontiguousarray(phi / phi.sum(), dtype=np.float64)


def postprocess_volume(merged: torch.Tensor, pre_sigma_z=0.8, sigma_z=0.7, sigma_xy=0.05, sharpen_amount=1.7, sharpen_radius=1.2,
                       hu_threshold=750, out: torch.Tensor | None = None) -> torch.Tensor:
    """merged: int16 CUDA tensor [S,H,W] of composite stored values -> int16 [S,H,W].  Defaults are the arguments of
    generate.py:258-263.  ``sigma_xy`` must be small enough for scipy to give it a radius-0 kernel (0.05 does): the
    reference's in-plane smoothing is an identity there."""
    if not merged.is_cuda or merged.dtype != torch.int16 or merged.dim() != 3 or not merged.is_contiguous():
        raise RuntimeError("postprocess_volume expects a contiguous int16 CUDA tensor [S,H,W] (no CPU path exists)")
    if int(4.0 * float(sigma_xy) + 0.5) != 0:
        raise NotImplementedError("in-plane Gaussian smoothing with a non-trivial kernel is not built (generate.py uses sigma_xy=0.05)")
    wz1, wz2, wxy = gaussian_kernel1d(pre_sigma_z), gaussian_kernel1d(sigma_z), gaussian_kernel1d(sharpen_radius)
    S, H, W = merged.shape
    lib = _lib.load()
    with torch.cuda.device(merged.device):
        if out is None:
            out = torch.empty_like(merged)
        scratch = torch.empty(lib.ducosy_postprocess_scratch_bytes(S, H, W) // 4, dtype=torch.float32, device=merged.device)
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        call("ducosy_postprocess_volume", ptr(merged), ptr(out), ptr(scratch), S, H, W, hp(wz1), len(wz1) // 2, hp(wz2), len(wz2) // 2,
             hp(wxy), len(wxy) // 2, float(sharpen_amount), float(hu_threshold), stream_ptr())
    return out
