"""Golden values for the image-quality metrics: imports the reference's own calculate.py -- unmodified;
its third-party imports that are absent here (pydicom, skimage, matplotlib, seaborn) are stubbed, and
skimage.metrics.structural_similarity is stubbed BY THE ORACLE'S RESTATEMENT, so calculate_ssim's own loop / data_range logic
runs but the SSIM core stays "parity unpinned" (likewise skimage.filters.sobel for calculate_ts) -- and runs calculate_mae /
calculate_psnr / calculate_ssim / calculate_cs / calculate_ed / calculate_emd / calculate_ts / normalize on seeded int16 volumes
and (the basic metrics) on their normalised float64 forms.  calculate_emd runs on the real scipy.stats.wasserstein_distance.  Checks the oracle restatements against
them and stores the values in tests/golden/metrics.npz.
usage: python oracle/make_golden_metrics.py   (in the container that has /root/reference)      TEST INFRASTRUCTURE ONLY."""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ducosy_oracle as orc  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


_stub("pydicom")
_stub("skimage")
_stub("skimage.metrics", structural_similarity=lambda a, b, data_range: orc.skimage_structural_similarity(a, b, data_range))
_stub("skimage.filters", sobel=lambda img: orc.skimage_sobel(img))       # TS core: the oracle's restatement (unpinned)
mpl = _stub("matplotlib", use=lambda *a, **k: None)
mpl.pyplot = _stub("matplotlib.pyplot")
_stub("seaborn")
_stub("lpips")

spec = importlib.util.spec_from_file_location("ref_calculate", os.path.join(REF, "calculate.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = {"a": (4, 128, 128, 3), "b": (3, 96, 160, 7)}


def main():
    out = {}
    warnings.simplefilter("ignore")          # numpy warns about the int16 overflow the reference runs into
    for name, (S, H, W, seed) in CASES.items():
        tgt, pred = orc.metrics_test_volumes(S, H, W, seed)
        tn, pn = ref.normalize(tgt), ref.normalize(pred)
        assert np.array_equal(tn, orc.metric_normalize(tgt))
        out[f"shape_{name}"] = np.array([S, H, W, seed])
        for tag, (x, y) in {"raw": (tgt, pred), "norm": (tn, pn)}.items():
            for metric in ("mae", "psnr", "ssim", "cs", "ed") + (("emd", "ts") if tag == "raw" else ()):   # advanced pairs: raw only (calculate.py:457-463)
                m, lst = getattr(ref, f"calculate_{metric}")(x, y)
                om, olst = getattr(orc, f"metric_{metric}")(x, y)
                assert np.allclose(np.asarray(lst, dtype=np.float64), np.asarray(olst, dtype=np.float64), rtol=1e-12, atol=0, equal_nan=True), (name, tag, metric)
                assert np.isclose(float(m), float(om), rtol=1e-12, atol=0) or (np.isinf(m) and np.isinf(om)), (name, tag, metric)
                out[f"{metric}_{tag}_{name}"] = np.concatenate([[float(m)], np.asarray(lst, dtype=np.float64)])
                print(name, tag, metric, float(m))
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print("metrics golden ok")


if __name__ == "__main__":
    main()
