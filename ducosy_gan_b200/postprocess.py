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


Z_HALO = 6   # two chained 7-tap z kernels (generate.py:259 then postprocess.py:60): an output slice sees +-6 input slices


def _weights(pre_sigma_z, sigma_z, sigma_xy, sharpen_radius):
    if int(4.0 * float(sigma_xy) + 0.5) != 0:
        raise NotImplementedError("in-plane Gaussian smoothing with a non-trivial kernel is not built (generate.py uses sigma_xy=0.05)")
    return gaussian_kernel1d(pre_sigma_z), gaussian_kernel1d(sigma_z), gaussian_kernel1d(sharpen_radius)


def _run(merged, out, scratch, w, sharpen_amount, hu_threshold, phases, mm_z0, mm_z1):
    S, H, W = merged.shape
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    call("ducosy_postprocess_volume", ptr(merged), ptr(out), ptr(scratch), S, H, W, hp(w[0]), len(w[0]) // 2, hp(w[1]), len(w[1]) // 2,
         hp(w[2]), len(w[2]) // 2, float(sharpen_amount), float(hu_threshold), int(phases), int(mm_z0), int(mm_z1), stream_ptr())


def _check(merged):
    if not merged.is_cuda or merged.dtype != torch.int16 or merged.dim() != 3 or not merged.is_contiguous():
        raise RuntimeError("postprocess_volume expects a contiguous int16 CUDA tensor [S,H,W] (no CPU path exists)")


def postprocess_volume(merged: torch.Tensor, pre_sigma_z=0.8, sigma_z=0.7, sigma_xy=0.05, sharpen_amount=1.7, sharpen_radius=1.2,
                       hu_threshold=750, out: torch.Tensor | None = None) -> torch.Tensor:
    """merged: int16 CUDA tensor [S,H,W] of composite stored values -> int16 [S,H,W].  Defaults are the arguments of
    generate.py:258-263.  ``sigma_xy`` must be small enough for scipy to give it a radius-0 kernel (0.05 does): the
    reference's in-plane smoothing is an identity there."""
    _check(merged)
    w = _weights(pre_sigma_z, sigma_z, sigma_xy, sharpen_radius)
    S, H, W = merged.shape
    lib = _lib.load()
    with torch.cuda.device(merged.device):
        if out is None:
            out = torch.empty_like(merged)
        scratch = torch.empty(lib.ducosy_postprocess_scratch_bytes(S, H, W) // 4, dtype=torch.float32, device=merged.device)
        _run(merged, out, scratch, w, sharpen_amount, hu_threshold, 3, 0, S)
    return out


def exchange_z_halo(local: torch.Tensor, halo: int = Z_HALO, group=None) -> tuple[torch.Tensor, int]:
    """Slices of a volume sharded over ranks in contiguous z ranges (``synthesis.shard_range``): returns this rank's slab
    extended by up to ``halo`` slices of each neighbour, and the index of its first own slice inside the slab.  Point to
    point only (each rank talks to rank-1 and rank+1); the outer ranks get no halo on the volume boundary."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local, 0
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = local.shape[0]
    if n < halo:
        raise RuntimeError(f"every rank needs at least {halo} slices for the z halo (this one has {n})")
    gr = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    lo_buf = torch.empty((halo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if rank > 0 else None
    hi_buf = torch.empty((halo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if rank < world - 1 else None
    raw = lambda t: t.view(torch.uint8)     # NCCL has no int16: slices travel as bytes
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, raw(local[:halo].contiguous()), gr(rank - 1), group),
                dist.P2POp(dist.irecv, raw(lo_buf), gr(rank - 1), group)]
    if rank < world - 1:
        ops += [dist.P2POp(dist.isend, raw(local[n - halo:].contiguous()), gr(rank + 1), group),
                dist.P2POp(dist.irecv, raw(hi_buf), gr(rank + 1), group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = [t for t in (lo_buf, local, hi_buf) if t is not None]
    return torch.cat(parts, dim=0), (halo if rank > 0 else 0)


def postprocess_volume_sharded(local: torch.Tensor, group=None, pre_sigma_z=0.8, sigma_z=0.7, sigma_xy=0.05, sharpen_amount=1.7,
                               sharpen_radius=1.2, hu_threshold=750) -> torch.Tensor:
    """``postprocess_volume`` for a volume whose slices are sharded over the ranks of ``group``: the one exchange the
    synthesis path has once the smoothing is on the device -- a 6-slice halo from each z neighbour (3 MB per side at 512x512)
    and a two-float min/max all-reduce for the clip range.  Bit-identical to the single-process result on the whole volume."""
    import torch.distributed as dist
    _check(local)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return postprocess_volume(local, pre_sigma_z, sigma_z, sigma_xy, sharpen_amount, sharpen_radius, hu_threshold)
    w = _weights(pre_sigma_z, sigma_z, sigma_xy, sharpen_radius)
    if max(len(w[0]) // 2 + len(w[1]) // 2, 1) > Z_HALO:
        raise NotImplementedError("z kernels wider than the 6-slice halo")
    slab, first = exchange_z_halo(local, Z_HALO, group)
    S, H, W = slab.shape
    n = local.shape[0]
    lib = _lib.load()
    with torch.cuda.device(local.device):
        out = torch.empty_like(slab)
        scratch = torch.empty(lib.ducosy_postprocess_scratch_bytes(S, H, W) // 4, dtype=torch.float32, device=local.device)
        _run(slab, out, scratch, w, sharpen_amount, hu_threshold, 1, first, first + n)
        mm = scratch[lib.ducosy_postprocess_minmax_offset_bytes(S, H, W) // 4:][:2]
        dist.all_reduce(mm[0:1], op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mm[1:2], op=dist.ReduceOp.MAX, group=group)
        _run(slab, out, scratch, w, sharpen_amount, hu_threshold, 2, first, first + n)
    return out[first:first + n].contiguous()
