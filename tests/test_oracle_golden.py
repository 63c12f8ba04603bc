"""Pins oracle/ducosy_oracle.py against vectors produced by the reference itself
(oracle/make_golden.py, run in the authoring container against /root/reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import ducosy_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _x(seed, shape):
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).uniform(-1, 1, size=shape).astype(np.float32))


@pytest.mark.parametrize("name", ["gen_c1_b2_cbam_64", "gen_c3_b1_plain_32", "gen_c2_b1_cbam_128"])
def test_generator_matches_reference(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    cin, nb, cbam = int(g["cin"]), int(g["blocks"]), bool(g["cbam"])
    astd = float(g["attn_std"])
    shapes = orc.generator_param_shapes(cin, nb, cbam)
    # state_dict layout is the reference's (keys recorded from the reference module)
    assert list(shapes.keys()) == [str(k) for k in g["keys"]]
    sd = orc.make_state_dict(shapes, int(g["wseed"]), attn_std=None if astd < 0 else astd)
    x = _x(int(g["xseed"]), (int(g["B"]), cin, int(g["H"]), int(g["W"])))
    with torch.no_grad():
        y = orc.generator_forward(sd, x, nb, cbam).numpy()
    # same fp32 ops, different composition order of instance-norm => tiny fp32 noise only
    np.testing.assert_allclose(y, g["y"], rtol=0, atol=2e-5)


def test_generator_full_size_matches_reference(golden_dir):
    g = _load(golden_dir, "gen_full_512.npz")
    sd = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), int(g["wseed"]), attn_std=float(g["attn_std"]))
    px = orc.synthetic_volume(1, 512, 512, seed=int(g["vseed"]))[0]
    x = torch.from_numpy(orc.hu_window(px, 1.0, -1024.0, *orc.SOFT_HU).astype(np.float32))[None, None]
    with torch.no_grad():
        y = orc.generator_forward(sd, x)[0, 0].numpy()
    np.testing.assert_allclose(y[::8, ::8], g["y_sub"], rtol=0, atol=5e-5)
    assert abs(float(np.abs(y).astype(np.float64).sum()) - float(g["y_abs_sum"])) < 1.0


def test_discriminator_matches_reference(golden_dir):
    g = _load(golden_dir, "disc_64.npz")
    shapes = orc.discriminator_param_shapes(1)
    assert list(shapes.keys()) == [str(k) for k in g["keys"]]
    assert int(g["n_conv_g"]) == 51 and int(g["n_conv_d"]) == 5
    sd = orc.make_state_dict(shapes, int(g["wseed"]))
    with torch.no_grad():
        y = orc.discriminator_forward(sd, _x(int(g["xseed"]), (2, 1, 64, 64))).numpy()
    np.testing.assert_allclose(y, g["y"], rtol=0, atol=2e-5)


def test_hu_window_and_dewindow_bit_exact(golden_dir):
    g = _load(golden_dir, "hu_window.npz")
    px = g["px"]
    for ci in range(3):
        slope, intercept = float(g[f"slope_{ci}"]), float(g[f"intercept_{ci}"])
        ws = orc.hu_window(px, slope, intercept, -150, 250)
        wl = orc.hu_window(px, slope, intercept, -1000, -150)
        assert ws.dtype == g[f"win_soft_{ci}"].dtype
        assert np.array_equal(ws, g[f"win_soft_{ci}"])
        assert np.array_equal(wl, g[f"win_lung_{ci}"])
        assert np.array_equal(orc.soft_squeeze_window(px, slope, intercept, -150, 250), g[f"sq_soft_{ci}"])
        assert np.array_equal(orc.soft_squeeze_window(px, slope, intercept, -1000, -150), g[f"sq_lung_{ci}"])
        assert np.array_equal(ws, g[f"lin_soft_{ci}"])
        y = g[f"y_{ci}"][0, 0]
        ps = orc.dewindow_to_stored(y, slope, intercept, -150, 250, np.int16)
        pl = orc.dewindow_to_stored(y, slope, intercept, -1000, -150, np.int16)
        assert np.array_equal(ps, g[f"post_soft_{ci}"])
        assert np.array_equal(pl, g[f"post_lung_{ci}"])
        assert np.array_equal(orc.apply_windowing(g[f"y_{ci}"], -150, 250, 40, 400), g[f"disp_soft_{ci}"])
        assert np.array_equal(orc.apply_windowing(g[f"y_{ci}"], -1000, -150, -600, 1500), g[f"disp_lung_{ci}"])


def test_composite_bit_exact(golden_dir):
    g = _load(golden_dir, "composite.npz")
    for ci in range(3):
        merged, sm, lm = orc.composite(g[f"raw_{ci}"], g[f"soft_px_{ci}"], g[f"lung_px_{ci}"],
                                       float(g[f"slope_{ci}"]), float(g[f"intercept_{ci}"]))
        assert np.array_equal(merged, g[f"merged_{ci}"])
        assert np.array_equal(sm, g[f"soft_mask_{ci}"])
        assert np.array_equal(lm, g[f"lung_mask_{ci}"])
    assert np.array_equal(orc.stored_to_hu(g["edge_px"], 1, 0), g["hu_notags"])


def test_composite_edge_table():
    # px 874/873/24/23/1274/1275 @ slope 1, intercept -1024 => HU -150/-151/-1000/-1001/250/251
    px = np.array([[874, 873, 24, 23, 1274, 1275]], np.int16)
    soft = np.full_like(px, 111)
    lung = np.full_like(px, 222)
    merged, sm, lm = orc.composite(px, soft, lung, 1.0, -1024.0)
    assert sm.tolist() == [[True, False, False, False, True, False]]
    assert lm.tolist() == [[True, True, True, False, False, False]]
    assert merged.tolist() == [[222, 222, 222, 23, 111, 1275]]   # lung wins at -150; outside both keeps raw
    assert orc.dewindow_to_stored(np.array([0.0]), 1.0, -1024.0, -150, 250).tolist() == [1074]
    assert np.array([873.575, -0.7], np.float32).astype(np.int16).tolist() == [873, 0]


def test_threshold_candidates_bit_exact(golden_dir):
    g = _load(golden_dir, "thresholds.npz")
    body, lung, bone = orc.threshold_candidates(g["hu"])
    assert np.array_equal(body, g["body"])
    assert np.array_equal(lung, g["lung"])
    assert np.array_equal(bone, g["bone"])


def test_losses_match_reference(golden_dir):
    g = _load(golden_dir, "losses.npz")
    p, t, s = (_x(int(k), (2, 1, 64, 64)) for k in g["seeds"])
    assert abs(orc.gradient_loss(p, t).item() - float(g["grad"])) < 1e-6
    assert abs(orc.contrast_attention_loss(p, t, s).item() - float(g["att"])) < 1e-6
    assert abs(orc.contrast_region_loss(p, t, s).item() - float(g["region"])) < 1e-6
    assert abs(orc.contrast_edge_loss(p, t, s).item() - float(g["edge"])) < 1e-6


def test_ssim_unpinned_restatement_sanity():
    # PARITY UNPINNED (pytorch_msssim absent): only self-consistency is checked.
    x = _x(1, (2, 1, 32, 32))
    assert abs(orc.ssim(x, x).item() - 1.0) < 1e-6
    assert orc.ssim(x, -x).item() < 0.99
    assert orc.ssim(x, 0.5 * x).item() < 1.0


def test_postprocess_volume_matches_reference_golden(golden_dir):
    """the oracle restatement against outputs of the reference's own generate.py:254-263 +
    modules/postprocess.py (tests/golden/postprocess.npz, made by oracle/make_golden_postprocess.py) -- bit-exact."""
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    for name in "abc":
        S, H, W, seed = (int(v) for v in g[f"shape_{name}"])
        vol = orc.postprocess_test_volume(S, H, W, seed)
        assert np.array_equal(orc.postprocess_volume(vol), g[f"out_{name}"]), name


def test_cyclegan_step_matches_reference_loop_body(golden_dir):
    """oracle.cyclegan_step against the losses the reference's OWN loop body (modules/trainer.py:448-525,
    executed from source by oracle/make_golden_trainstep.py on the reference's modules, criteria and optimisers) logs over
    two consecutive iterations -- pins the composition (loss mix, update order, detach points), not just the parts.
    SSIM inside it is the parity-unpinned restatement on both sides."""
    g = _load(golden_dir, "train_step.npz")
    cin, blocks, cbam, B, n = int(g["cin"]), int(g["blocks"]), bool(g["cbam"]), int(g["B"]), int(g["size"])
    gen = torch.Generator().manual_seed(int(g["batch_seed"]))
    smooth = lambda t: torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0
    A = smooth(torch.rand(B, 1, n, n, generator=gen) * 2 - 1).clamp(-1, 1)
    Bt = smooth(torch.rand(B, 1, n, n, generator=gen) * 2 - 1).clamp(-1, 1)
    M = (torch.rand(B, cin - 1, n, n, generator=gen) < 0.1).float()
    mk = lambda shapes, seed: {k: v.clone().requires_grad_(True) for k, v in orc.make_state_dict(shapes, int(seed)).items()}
    gs, ds = orc.generator_param_shapes(cin, blocks, cbam), orc.discriminator_param_shapes(1)
    sds = (mk(gs, g["seeds"][0]), mk(gs, g["seeds"][1]), mk(ds, g["seeds"][2]), mk(ds, g["seeds"][3]))
    adam = lambda ps: torch.optim.Adam(ps, lr=2e-4, betas=(0.5, 0.999))
    opts = (adam(list(sds[0].values()) + list(sds[1].values())), adam(list(sds[2].values())), adam(list(sds[3].values())))
    names = {"G": "loss_G", "GAN": "loss_GAN", "cycle": "loss_cycle", "id": "loss_id", "grad_cycle": "loss_grad_cycle",
             "grad_id": "loss_grad_id", "ssim": "loss_ssim", "contrast_attention": "loss_contrast_attention",
             "contrast_region": "loss_contrast_region", "contrast_edge": "loss_contrast_edge", "D_A": "loss_D_A", "D_B": "loss_D_B"}
    for it in range(len(g["loss_G"])):
        out = orc.cyclegan_step(sds, opts, A, Bt, M, blocks, cbam)
        for k, ref_name in names.items():
            ref = float(g[ref_name][it])
            # iteration 0 is pure forward arithmetic (1e-5); afterwards Adam's first updates are +-lr whatever the gradient's
            # size, so fp32 noise on near-zero gradient entries flips a few of them (measured 3e-4 on one term): 2e-3
            tol = 1e-5 if it == 0 else 2e-3
            assert abs(out[k] - ref) <= tol * abs(ref) + 1e-6, (it, k, out[k], ref)


def test_oracle_masks_match_reference_golden(golden_dir):
    """SURVEY 8f N2 (first half): the oracle's restatement of modules/mask_generator.py:detect_lung / detect_lung_vessels against
    the masks the reference's own functions produced (oracle/make_golden_masks.py)."""
    g = np.load(os.path.join(golden_dir, "masks.npz"))
    for name in "ab":
        B, H, W, seed = (int(v) for v in g[f"shape_{name}"])
        hu = orc.mask_test_slices(B, H, W, seed)
        unpack = lambda key: np.unpackbits(g[key])[: B * H * W].reshape(B, H, W)
        lung = orc.mask_detect_lung(hu)
        assert np.array_equal(lung, unpack(f"lung_{name}"))
        assert np.array_equal(orc.mask_detect_lung_vessels(hu, lung), unpack(f"vessel_{name}"))


def test_oracle_hull_masks_match_reference_golden(golden_dir):
    """SURVEY 8f N2 (second half): the oracle's detect_mediastinum / detect_bone restatements against masks the reference's own
    functions produced with scipy's ConvexHull / label / binary_fill_holes and the oracle's contains_points (matplotlib absent:
    that step is parity-unpinned); plus the crossings test itself on a square, where only boundary conventions can differ."""
    g = np.load(os.path.join(golden_dir, "masks.npz"))
    for name in "cd":
        B, H, W, seed = (int(v) for v in g[f"shape_{name}"])
        hu = orc.mask_test_slices_bone(B, H, W, seed)
        unpack = lambda key: np.unpackbits(g[key])[: B * H * W].reshape(B, H, W)
        lung = orc.mask_detect_lung(hu)
        assert np.array_equal(orc.mask_detect_mediastinum(hu, lung), unpack(f"mediastinum_{name}"))
        assert np.array_equal(orc.mask_detect_bone(hu, lung), unpack(f"bone_{name}"))
    sq = np.array([[2, 2], [6, 2], [6, 6], [2, 6]])                     # counter-clockwise square
    ys, xs = np.mgrid[:9, :9]
    inside = orc.path_contains_points(sq, np.vstack((ys.ravel(), xs.ravel())).T).reshape(9, 9)
    assert inside[3:6, 3:6].all() and not inside[:2].any() and not inside[7:].any() and not inside[:, :2].any() and not inside[:, 7:].any()


def test_oracle_metrics_match_reference_golden(golden_dir):
    """the oracle's restatements of calculate.py:232-271,360-381 against values the reference's own functions
    produced (oracle/make_golden_metrics.py).  SSIM's core is skimage (absent): parity unpinned, see the oracle docstring."""
    import warnings
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    warnings.simplefilter("ignore")
    for name in ("a", "b"):
        S, H, W, seed = (int(v) for v in g[f"shape_{name}"])
        tgt, pred = orc.metrics_test_volumes(S, H, W, seed)
        pairs = {"raw": (tgt, pred), "norm": (orc.metric_normalize(tgt), orc.metric_normalize(pred))}
        for tag, (x, y) in pairs.items():
            for metric in ("mae", "psnr", "ssim", "cs", "ed") + (("emd", "ts") if tag == "raw" else ()):
                m, lst = getattr(orc, f"metric_{metric}")(x, y)
                got = np.concatenate([[float(m)], np.asarray(lst, dtype=np.float64)])
                assert np.allclose(got, g[f"{metric}_{tag}_{name}"], rtol=1e-12, atol=0), (name, tag, metric)
