"""Anatomical-mask primitives (first half) on the CUDA path, through the C ABI: bit-exact against
scipy.ndimage (labels with scipy's numbering, hole filling) and against the golden masks produced by the reference's own
modules/mask_generator.py:detect_lung / detect_lung_vessels."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ducosy_oracle as orc  # noqa: E402


def _masks(B, H, W, p, seed):
    return (np.random.Generator(np.random.PCG64(seed)).random((B, H, W)) < p).astype(np.uint8)


def _spiral(n):
    """one long winding component (a worst case for union-find chains) and its complement as a second one"""
    m = np.zeros((n, n), np.uint8)
    lo, hi = 0, n - 1
    while lo <= hi:
        m[lo, lo:hi + 1] = 1
        m[lo:hi + 1, hi] = 1
        m[hi, lo:hi + 1] = 1
        m[lo + 2:hi + 1, lo] = 1
        if lo + 2 <= hi:
            m[lo + 2, lo:lo + 2] = 1
        lo += 2
        hi -= 2
    return m


@pytest.mark.parametrize("shape,p", [((1, 1, 1), 1.0), ((2, 7, 13), 0.5), ((3, 64, 96), 0.3), ((2, 512, 512), 0.593), ((1, 512, 512), 0.9),
                                      ((4, 33, 1), 0.6), ((1, 300, 257), 0.05)])
def test_label4_matches_scipy_including_numbering(shape, p):
    from ducosy_gan_b200 import mask_generator as mg
    m = _masks(*shape, p, seed=sum(shape))
    labels, num = mg.label(torch.from_numpy(m).cuda())
    ref_labels, ref_num = orc.mask_label4(m)
    assert np.array_equal(num.cpu().numpy(), ref_num), (num, ref_num)
    assert np.array_equal(labels.cpu().numpy(), ref_labels)


def test_label4_and_fill_holes_special_cases():
    from scipy import ndimage
    from ducosy_gan_b200 import mask_generator as mg
    sp = _spiral(255)
    cases = [np.zeros((40, 50), np.uint8), np.ones((40, 50), np.uint8), sp, 1 - sp]
    ring = np.zeros((64, 64), np.uint8)
    ring[10:50, 10:50] = 1
    ring[20:40, 20:40] = 0
    ring[25:30, 25:30] = 1            # an island inside the hole
    ring[0:5, 60:64] = 1
    ring[1:4, 61:63] = 0              # a hole that is not one: it touches nothing but its own component... and is enclosed
    cases.append(ring)
    edge = np.ones((32, 32), np.uint8)
    edge[0, 5] = 0                    # background pixel on the border: connected to the outside, stays background
    edge[10:12, 10:12] = 0
    cases.append(edge)
    for m in cases:
        labels, num = mg.label(torch.from_numpy(m).cuda())          # 2-D input: (labels [H,W], int)
        ref_labels, ref_num = ndimage.label(m)
        assert num == ref_num and np.array_equal(labels.cpu().numpy(), ref_labels)
        filled = mg.binary_fill_holes(torch.from_numpy(m).cuda()).cpu().numpy()
        assert np.array_equal(filled, ndimage.binary_fill_holes(m).astype(np.uint8))


@pytest.mark.parametrize("shape,p", [((3, 64, 96), 0.55), ((2, 512, 512), 0.62), ((5, 17, 9), 0.7)])
def test_fill_holes_matches_scipy_on_random_masks(shape, p):
    from ducosy_gan_b200 import mask_generator as mg
    m = _masks(*shape, p, seed=7 + sum(shape))
    got = mg.binary_fill_holes(torch.from_numpy(m).cuda()).cpu().numpy()
    assert np.array_equal(got, orc.mask_fill_holes(m))
    assert got.sum() > m.sum()                      # there were holes to fill


def test_detect_lung_and_vessels_match_reference_golden(golden_dir):
    from ducosy_gan_b200 import mask_generator as mg
    g = np.load(os.path.join(golden_dir, "masks.npz"))
    for name in "ab":
        B, H, W, seed = (int(v) for v in g[f"shape_{name}"])
        hu = torch.from_numpy(orc.mask_test_slices(B, H, W, seed)).cuda()
        lung = mg.detect_lung(hu)
        vessel = mg.detect_lung_vessels(hu, lung)
        unpack = lambda key: np.unpackbits(g[key])[: B * H * W].reshape(B, H, W)
        assert np.array_equal(lung.cpu().numpy(), unpack(f"lung_{name}")), name
        assert np.array_equal(vessel.cpu().numpy(), unpack(f"vessel_{name}")), name
        one = mg.detect_lung(hu[1])                                     # 2-D call, like the reference's per-slice use in dataset.py
        assert np.array_equal(one.cpu().numpy(), unpack(f"lung_{name}")[1])
        masks = mg.generate_anatomical_masks(hu, ("lung", "lung_vessel"))
        assert torch.equal(masks["lung"], lung) and torch.equal(masks["lung_vessel"], vessel)
    with pytest.raises(NotImplementedError):
        mg.detect_bone(hu, lung)


def test_detect_lung_and_vessels_at_training_size_and_other_parameters():
    """a training batch: 8 slices of 512x512 (BASELINE config 4 size), default and non-default parameters, vs the oracle"""
    from ducosy_gan_b200 import mask_generator as mg
    hu = orc.mask_test_slices(8, 512, 512, seed=9)
    dev = torch.from_numpy(hu).cuda()
    for kw in (dict(), dict(lung_lower=-950, lung_upper=-400, min_size=200, border_margin=0), dict(min_size=0, border_margin=100)):
        lung = mg.detect_lung(dev, **kw)
        ref = orc.mask_detect_lung(hu, **kw)
        assert np.array_equal(lung.cpu().numpy(), ref), kw
        vessel = mg.detect_lung_vessels(dev, lung, vessel_lower=-300, vessel_upper=600)
        assert np.array_equal(vessel.cpu().numpy(), orc.mask_detect_lung_vessels(hu, ref))
    assert torch.equal(mg.detect_lung(dev), mg.detect_lung(dev))             # atomics inside, deterministic outside
