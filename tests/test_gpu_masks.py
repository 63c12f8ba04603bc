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


def test_lung_hull_vertices_match_scipy_convex_hull():
    """The GPU hull (per-row extremes + monotone chain) against scipy.spatial.ConvexHull on the lung masks of the phantom and on
    random blobs: same strict corners, same counter-clockwise cyclic order; degenerate sets (empty, < 3 pixels, one row, one
    diagonal) take the reference's fallback (hull == lung, no vertices)."""
    from scipy.spatial import ConvexHull
    from ducosy_gan_b200 import mask_generator as mg
    rng = np.random.Generator(np.random.PCG64(5))
    masks = [orc.mask_detect_lung(orc.mask_test_slices_bone(4, 192, 160, 2))]
    blobs = np.zeros((6, 192, 160), np.uint8)
    for b in range(4):
        for _ in range(3 + 4 * b):
            y, x, r = rng.integers(10, 180), rng.integers(10, 150), rng.integers(1, 12)
            yy, xx = np.ogrid[:192, :160]
            blobs[b][(yy - y) ** 2 + (xx - x) ** 2 <= r * r] = 1
    blobs[4, 50, 20:90] = 1                                   # one row: collinear
    blobs[5, np.arange(30, 90), np.arange(30, 90)] = 1        # one diagonal: collinear
    masks.append(blobs)
    masks.append(np.zeros((2, 192, 160), np.uint8))
    masks[-1][1, 7, 9] = masks[-1][1, 100, 3] = 1             # two pixels
    m = np.concatenate(masks)
    hull, verts, nv = mg.lung_hull(torch.from_numpy(m).cuda(), return_vertices=True)
    hull, verts, nv = hull.cpu().numpy(), verts.cpu().numpy(), nv.cpu().numpy()
    seen_real = 0
    for z in range(m.shape[0]):
        coords = np.argwhere(m[z] == 1)
        try:
            want = coords[ConvexHull(coords).vertices] if len(coords) >= 3 else None
        except Exception:
            want = None
        if want is None:
            assert nv[z] == 0 and np.array_equal(hull[z], m[z]), z
            continue
        seen_real += 1
        got = verts[z, :nv[z]]
        assert len(got) == len(want), (z, len(got), len(want))
        k = int(np.flatnonzero((want == got[0]).all(1))[0])
        assert np.array_equal(np.roll(want, -k, axis=0), got), z          # same corners, same (counter-clockwise) cyclic order
        ref_mask, ok = orc.mask_convex_hull(m[z])
        assert ok and np.array_equal(hull[z], ref_mask), z                # rasterisation == the oracle's crossings test
    assert seen_real >= 8


def test_detect_mediastinum_and_bone_match_reference_golden(golden_dir):
    """Second half of the row: masks made by the reference's own detect_mediastinum / detect_bone (hull rasterisation through
    the oracle's restatement of matplotlib's contains_points: that step is parity-unpinned), 3-D and per-slice calls."""
    from ducosy_gan_b200 import mask_generator as mg
    g = np.load(os.path.join(golden_dir, "masks.npz"))
    for name in "cd":
        B, H, W, seed = (int(v) for v in g[f"shape_{name}"])
        hu = torch.from_numpy(orc.mask_test_slices_bone(B, H, W, seed)).cuda()
        unpack = lambda key: np.unpackbits(g[key])[: B * H * W].reshape(B, H, W)
        lung = mg.detect_lung(hu)
        med, bone = mg.detect_mediastinum(hu, lung), mg.detect_bone(hu, lung)
        assert np.array_equal(med.cpu().numpy(), unpack(f"mediastinum_{name}")), name
        assert np.array_equal(bone.cpu().numpy(), unpack(f"bone_{name}")), name
        assert np.array_equal(mg.detect_bone(hu[0], lung[0]).cpu().numpy(), unpack(f"bone_{name}")[0])
        masks = mg.generate_anatomical_masks(hu)
        assert list(masks) == ["lung", "mediastinum", "bone", "lung_vessel"]
        assert torch.equal(masks["mediastinum"], med) and torch.equal(masks["bone"], bone)
    # other parameters, training size, against the oracle
    hu = orc.mask_test_slices_bone(4, 512, 512, seed=6)
    dev = torch.from_numpy(hu).cuda()
    lung = orc.mask_detect_lung(hu)
    ldev = torch.from_numpy(lung).cuda()
    assert np.array_equal(mg.detect_mediastinum(dev, ldev, -200, 300).cpu().numpy(), orc.mask_detect_mediastinum(hu, lung, -200, 300))
    assert np.array_equal(mg.detect_bone(dev, ldev, 250, 0.4).cpu().numpy(), orc.mask_detect_bone(hu, lung, 250, 0.4))


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
