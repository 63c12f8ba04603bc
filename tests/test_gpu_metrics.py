"""Image-quality metrics on the GPU (pytest -m gpu): ducosy_gan_b200.metrics against the golden values the
reference's calculate.py produced (tests/golden/metrics.npz, oracle/make_golden_metrics.py) and against the oracle live."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ducosy_oracle as orc  # noqa: E402

# float64 reductions in a different (fixed) order than numpy's pairwise sums: 1e-11 relative.  Integer sums (MAE / PSNR of int16
# volumes, including numpy's int16 wrap-around) are exact, so those match to the last bit of the final division / log10.
RTOL = 1e-11


def _check(got, want, rtol=RTOL):
    m, lst = got
    got = np.concatenate([[m], np.asarray(lst, dtype=np.float64)])
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=rtol, atol=0), (got, want)


@pytest.mark.parametrize("name", ["a", "b"])
def test_metrics_match_reference_golden(golden_dir, name):
    from ducosy_gan_b200 import metrics
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    S, H, W, seed = (int(v) for v in g[f"shape_{name}"])
    tgt, pred = orc.metrics_test_volumes(S, H, W, seed)
    t, p = torch.from_numpy(tgt).cuda(), torch.from_numpy(pred).cuda()
    for tag, (x, y) in {"raw": (t, p), "norm": (metrics.normalize(t), metrics.normalize(p))}.items():
        for metric in ("mae", "psnr", "ssim", "cs", "ed") + (("emd", "ts") if tag == "raw" else ()):
            _check(getattr(metrics, f"calculate_{metric}")(x, y), g[f"{metric}_{tag}_{name}"])
    # EMD of float volumes (sorted-sample branch) == the histogram branch on the same values
    _check(metrics.calculate_emd(t.double(), p.double()), g[f"emd_raw_{name}"], rtol=1e-9)
    # int16 MAE is integer arithmetic end to end: bit-exact
    assert metrics.calculate_mae(t, p)[0] == g[f"mae_raw_{name}"][0]
    assert np.array_equal(metrics.normalize(t).cpu().numpy(), orc.metric_normalize(tgt))


def test_volume_metrics_full_size_vs_oracle():
    """One pass over a 6 x 512 x 512 pair (the evaluation of a synthesized volume) against the oracle; identical volumes give
    inf PSNR, SSIM 1, CS 1, ED 0."""
    from ducosy_gan_b200 import metrics
    tgt, pred = orc.metrics_test_volumes(6, 512, 512, 11)
    out = metrics.volume_metrics(tgt, pred)
    warnings.simplefilter("ignore")
    tn, pn = orc.metric_normalize(tgt), orc.metric_normalize(pred)
    for key, fn, (x, y) in [("emd", orc.metric_emd, (tgt, pred)), ("ts", orc.metric_ts, (tgt, pred)), ("mae", orc.metric_mae, (tgt, pred)), ("psnr", orc.metric_psnr, (tgt, pred)), ("ssim", orc.metric_ssim, (tgt, pred)),
                            ("cs", orc.metric_cs, (tgt, pred)), ("ed", orc.metric_ed, (tgt, pred)), ("mae_norm", orc.metric_mae, (tn, pn)),
                            ("psnr_norm", orc.metric_psnr, (tn, pn)), ("ssim_norm", orc.metric_ssim, (tn, pn))]:
        m, lst = fn(x, y)
        _check(out[key], np.concatenate([[float(m)], np.asarray(lst, dtype=np.float64)]))
    same = metrics.volume_metrics(tgt, tgt)
    assert same["psnr"][0] == float("inf") and same["mae"][0] == 0.0 and same["ed"][0] == 0.0
    assert abs(same["ssim"][0] - 1.0) < 1e-12 and abs(same["cs"][0] - 1.0) < 1e-12


def test_metrics_float32_and_shape_errors():
    from ducosy_gan_b200 import metrics
    rng = np.random.Generator(np.random.PCG64(1))
    a, b = rng.normal(size=(2, 64, 80)).astype(np.float32), rng.normal(size=(2, 64, 80)).astype(np.float32)
    m, lst = metrics.calculate_mae(a, b)
    ref = np.abs(a.astype(np.float64) - b.astype(np.float64))
    assert abs(m - ref.mean()) < 1e-12 and np.allclose(lst, ref.mean(axis=(1, 2)), rtol=1e-12)
    with pytest.raises(ValueError):
        metrics.calculate_mae(a, b[:1])
    with pytest.raises(TypeError):
        metrics.calculate_mae(a.astype(np.int32), b.astype(np.int32))
