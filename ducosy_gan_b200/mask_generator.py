"""Anatomical masks for training batches on the GPU: the reference's ``modules/mask_generator.py``.
First the scipy.ndimage work -- connected components, hole filling, the lung mask and the lung-vessel mask -- bit-exact with
scipy, on batches of HU slices that are already on the device.

The reference computes these per slice on the CPU inside the dataloader workers (``modules/dataset.py:129-158``, about 0.1 s per
512x512 slice, MASK_GENERATION_GUIDE.md:140-142); at the step rates of the CUDA training path that caps the input pipeline
well below what the GPUs consume.  Same names and argument meaning as the reference's functions; inputs are CUDA tensors
``[H,W]`` or ``[B,H,W]`` (every slice is treated on its own, exactly like the reference's 3-D branch), outputs uint8 {0,1}.

Second half of the row: ``detect_mediastinum`` / ``detect_bone`` on the rasterised convex hull of the lungs.  The hull
vertices are scipy's (``scipy.spatial.ConvexHull``: pinned), labelling / hole filling are the bit-exact primitives above; the
rasterisation follows ``matplotlib.path.Path.contains_points`` restated from matplotlib's crossings test -- matplotlib is absent
here, so THAT step is parity-unpinned (it decides only pixels exactly on the hull's boundary).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def _as_batch(t, dtype):
    if not t.is_cuda:
        raise RuntimeError("ducosy_gan_b200.mask_generator needs CUDA tensors on an sm_100 (B200) device; no CPU path exists")
    if t.dim() not in (2, 3):
        raise RuntimeError(f"expected [H,W] or [B,H,W], got {tuple(t.shape)}")
    squeeze = t.dim() == 2
    t = t.to(dtype).contiguous()
    return (t[None] if squeeze else t), squeeze


def _scratch(B, H, W, device):
    need = _lib.load().ducosy_masks_scratch_bytes(B, H, W)
    buf = torch.empty(need + 256, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 256
    return buf, buf.data_ptr() + off, need


def label(mask):
    """``scipy.ndimage.label(mask)`` (default structure: 4-connectivity) per slice.
    Returns (labels int32, num_features: int for a 2-D input, int32 tensor [B] for a batch)."""
    m, squeeze = _as_batch(mask != 0, torch.uint8)
    B, H, W = m.shape
    with torch.cuda.device(m.device):
        labels = torch.empty((B, H, W), dtype=torch.int32, device=m.device)
        num = torch.empty(B, dtype=torch.int32, device=m.device)
        buf, sp, sb = _scratch(B, H, W, m.device)
        call("ducosy_label4", ptr(m), ptr(labels), ptr(num), B, H, W, sp, sb, stream_ptr())
    return (labels[0], int(num[0].item())) if squeeze else (labels, num)


def binary_fill_holes(mask):
    """``scipy.ndimage.binary_fill_holes(mask)`` (default structure) per slice -> uint8 {0,1}."""
    m, squeeze = _as_batch(mask != 0, torch.uint8)
    B, H, W = m.shape
    with torch.cuda.device(m.device):
        out = torch.empty_like(m)
        buf, sp, sb = _scratch(B, H, W, m.device)
        call("ducosy_binary_fill_holes", ptr(m), ptr(out), B, H, W, sp, sb, stream_ptr())
    return out[0] if squeeze else out


def detect_lung(hu_volume, lung_lower=-1000, lung_upper=-300, min_size=64, border_margin=32):
    """reference modules/mask_generator.py:11-52."""
    hu, squeeze = _as_batch(hu_volume, torch.float32)
    B, H, W = hu.shape
    with torch.cuda.device(hu.device):
        out = torch.empty((B, H, W), dtype=torch.uint8, device=hu.device)
        buf, sp, sb = _scratch(B, H, W, hu.device)
        call("ducosy_detect_lung", ptr(hu), ptr(out), B, H, W, float(lung_lower), float(lung_upper), int(min_size), int(border_margin),
             sp, sb, stream_ptr())
    return out[0] if squeeze else out


def detect_lung_vessels(hu_volume, lung_mask, vessel_lower=-300, vessel_upper=600):
    """reference modules/mask_generator.py:55-99."""
    hu, squeeze = _as_batch(hu_volume, torch.float32)
    lm, _ = _as_batch(lung_mask != 0, torch.uint8)
    if lm.shape != hu.shape:
        raise RuntimeError(f"lung_mask {tuple(lung_mask.shape)} does not match hu_volume {tuple(hu_volume.shape)}")
    B, H, W = hu.shape
    with torch.cuda.device(hu.device):
        out = torch.empty((B, H, W), dtype=torch.uint8, device=hu.device)
        buf, sp, sb = _scratch(B, H, W, hu.device)
        call("ducosy_detect_lung_vessels", ptr(hu), ptr(lm), ptr(out), B, H, W, float(vessel_lower), float(vessel_upper), sp, sb,
             stream_ptr())
    return out[0] if squeeze else out


def _pair(hu_volume, lung_mask):
    hu, squeeze = _as_batch(hu_volume, torch.float32)
    lm, _ = _as_batch(lung_mask != 0, torch.uint8)
    if lm.shape != hu.shape:
        raise RuntimeError(f"lung_mask {tuple(lung_mask.shape)} does not match hu_volume {tuple(hu_volume.shape)}")
    return hu, lm, squeeze


def lung_hull(lung_mask, return_vertices=False):
    """Rasterised convex hull of every slice's lung pixels, as mask_generator.py:115-127 builds it (ConvexHull vertices ->
    ``Path.contains_points`` of every pixel).  ``return_vertices``: also (vertices int32 [B,2H+4,2] in scipy's counter-clockwise
    (row, col) order, count int32 [B]; count 0 = the reference's fallback, hull == lung)."""
    lm, squeeze = _as_batch(lung_mask != 0, torch.uint8)
    B, H, W = lm.shape
    with torch.cuda.device(lm.device):
        out = torch.empty((B, H, W), dtype=torch.uint8, device=lm.device)
        verts = torch.zeros((B, 2 * H + 4, 2), dtype=torch.int32, device=lm.device)
        nv = torch.zeros(B, dtype=torch.int32, device=lm.device)
        buf, sp, sb = _scratch(B, H, W, lm.device)
        call("ducosy_lung_hull", ptr(lm), ptr(out), ptr(verts), ptr(nv), B, H, W, sp, sb, stream_ptr())
    out = out[0] if squeeze else out
    return (out, verts, nv) if return_vertices else out


def detect_mediastinum(hu_volume, lung_mask, mediastinum_lower=-300, mediastinum_upper=450):
    """reference modules/mask_generator.py:100-170."""
    hu, lm, squeeze = _pair(hu_volume, lung_mask)
    B, H, W = hu.shape
    with torch.cuda.device(hu.device):
        out = torch.empty((B, H, W), dtype=torch.uint8, device=hu.device)
        buf, sp, sb = _scratch(B, H, W, hu.device)
        call("ducosy_detect_mediastinum", ptr(hu), ptr(lm), ptr(out), B, H, W, float(mediastinum_lower), float(mediastinum_upper), sp, sb,
             stream_ptr())
    return out[0] if squeeze else out


def detect_bone(hu_volume, lung_mask, bone_threshold=200, spine_margin_ratio=0.25):
    """reference modules/mask_generator.py:173-311."""
    hu, lm, squeeze = _pair(hu_volume, lung_mask)
    B, H, W = hu.shape
    spine_start = int(H * (1 - spine_margin_ratio))          # mask_generator.py:219 (python float arithmetic)
    with torch.cuda.device(hu.device):
        out = torch.empty((B, H, W), dtype=torch.uint8, device=hu.device)
        buf, sp, sb = _scratch(B, H, W, hu.device)
        call("ducosy_detect_bone", ptr(hu), ptr(lm), ptr(out), B, H, W, float(bone_threshold), int(spine_start), sp, sb, stream_ptr())
    return out[0] if squeeze else out


def generate_anatomical_masks(hu_image, mask_types=("lung", "mediastinum", "bone", "lung_vessel")):
    """reference modules/mask_generator.py:313-347."""
    masks = {}
    lung_mask = detect_lung(hu_image)
    if "lung" in mask_types:
        masks["lung"] = lung_mask
    if "mediastinum" in mask_types:
        masks["mediastinum"] = detect_mediastinum(hu_image, lung_mask)
    if "bone" in mask_types:
        masks["bone"] = detect_bone(hu_image, lung_mask)
    if "lung_vessel" in mask_types:
        masks["lung_vessel"] = detect_lung_vessels(hu_image, lung_mask)
    return masks
