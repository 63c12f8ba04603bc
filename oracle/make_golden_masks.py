"""Golden vectors for the anatomical masks (first half): runs the reference's own
modules/mask_generator.py:detect_lung / detect_lung_vessels -- unmodified; `matplotlib.path` (absent here, imported at
mask_generator.py:8, used only by detect_mediastinum / detect_bone) is stubbed -- on seeded HU slices, both through the 3-D
branch and slice by slice through the 2-D branch, and stores the masks (bit-packed) in tests/golden/masks.npz.
usage: python oracle/make_golden_masks.py   (in the container that has /root/reference)      TEST INFRASTRUCTURE ONLY."""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ducosy_oracle as orc  # noqa: E402

class _Path:
    """Stand-in for matplotlib.path.Path (absent here): contains_points is the ORACLE'S restatement of matplotlib's crossings test
    (oracle.path_contains_points), so detect_mediastinum / detect_bone run the reference's own control flow, scipy's ConvexHull /
    label / binary_fill_holes, and an UNPINNED rasterisation of the hull."""

    def __init__(self, vertices):
        self.vertices = np.asarray(vertices)

    def contains_points(self, points):
        return orc.path_contains_points(self.vertices, points)


stub = types.ModuleType("matplotlib")
stub_path = types.ModuleType("matplotlib.path")
stub_path.Path = _Path
stub.path = stub_path
sys.modules.setdefault("matplotlib", stub)
sys.modules.setdefault("matplotlib.path", stub_path)
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_mask_generator", os.path.join(REF, "modules", "mask_generator.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def main():
    out = {}
    for name, (B, H, W, seed) in {"a": (5, 256, 256, 1), "b": (3, 128, 192, 2)}.items():
        hu = orc.mask_test_slices(B, H, W, seed)
        lung = ref.detect_lung(hu.copy())                               # reference, 3-D branch
        vessel = ref.detect_lung_vessels(hu.copy(), lung.copy())
        for z in range(B):                                              # reference, 2-D branch: same answer per slice
            l2 = ref.detect_lung(hu[z].copy())
            assert np.array_equal(l2, lung[z]), (name, z)
            assert np.array_equal(ref.detect_lung_vessels(hu[z].copy(), l2.copy()), vessel[z]), (name, z)
        assert np.array_equal(lung, orc.mask_detect_lung(hu)), name                      # the oracle restatement is pinned here
        assert np.array_equal(vessel, orc.mask_detect_lung_vessels(hu, lung)), name
        assert lung.sum() > 0 and vessel.sum() > 0, "the test slices must exercise both masks"
        out[f"shape_{name}"] = np.array([B, H, W, seed])
        out[f"lung_{name}"] = np.packbits(lung.astype(np.uint8))
        out[f"vessel_{name}"] = np.packbits(vessel.astype(np.uint8))
        print(name, "lung pixels", int(lung.sum()), "vessel pixels", int(vessel.sum()), "per slice", lung.reshape(B, -1).sum(1), vessel.reshape(B, -1).sum(1))
    # second half of the row: the hull-based masks (bone-range structures added to the phantom)
    for name, (B, H, W, seed) in {"c": (5, 256, 256, 3), "d": (3, 160, 224, 4)}.items():
        hu = orc.mask_test_slices_bone(B, H, W, seed)
        lung = ref.detect_lung(hu.copy())
        med = ref.detect_mediastinum(hu.copy(), lung.copy())                 # reference, 3-D branch
        bone = ref.detect_bone(hu.copy(), lung.copy())
        for z in range(B):                                                   # reference, 2-D branch: same answer per slice
            assert np.array_equal(ref.detect_mediastinum(hu[z].copy(), lung[z].copy()), med[z]), (name, z)
            assert np.array_equal(ref.detect_bone(hu[z].copy(), lung[z].copy()), bone[z]), (name, z)
        assert np.array_equal(med, orc.mask_detect_mediastinum(hu, lung)), name
        assert np.array_equal(bone, orc.mask_detect_bone(hu, lung)), name
        cand = (hu >= 200) & (hu > -1000)
        assert med.sum() > 0 and 0 < bone.sum() and (bone.astype(bool) != cand).any(), "the slices must exercise the hull branches"
        full = ref.generate_anatomical_masks(hu.copy())
        assert all(np.array_equal(full[k], v) for k, v in (("lung", lung), ("mediastinum", med), ("bone", bone)))
        out[f"shape_{name}"] = np.array([B, H, W, seed])
        out[f"mediastinum_{name}"] = np.packbits(med.astype(np.uint8))
        out[f"bone_{name}"] = np.packbits(bone.astype(np.uint8))
        print(name, "mediastinum pixels", med.reshape(B, -1).sum(1), "bone pixels", bone.reshape(B, -1).sum(1), "candidates", cand.reshape(B, -1).sum(1))
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)
    print("masks golden ok")


if __name__ == "__main__":
    main()
