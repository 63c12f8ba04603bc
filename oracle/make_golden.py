"""Generate tests/golden/*.npz by running the REFERENCE itself (authoring container only).

Run:  python oracle/make_golden.py      (needs /root/reference; not available on the GPU box)

The reference modules are imported from /root/reference unmodified.  Third-party packages that
are absent in this image (pydicom, pytorch_msssim, matplotlib.path) are stubbed in sys.modules --
none of them is touched by the arithmetic that is recorded here.  The composite and the HU
threshold candidates live inside functions that also do DICOM I/O / scipy morphology, so their
statements (generate.py:140-145,218-237; mask_generator.py:14-20,179-183) are executed verbatim
from the reference source via exec() on the cited line ranges.

Weights come from oracle.ducosy_oracle.make_state_dict (numpy PCG64 keyed by name), loaded
strictly into the reference modules -- which also pins the state_dict key/shape layout.
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

# ---- stubs for absent third-party modules (never used by the recorded arithmetic) ----
for name in ("pydicom", "pytorch_msssim", "matplotlib", "matplotlib.path"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            sys.modules[name] = m
sys.modules["pytorch_msssim"].SSIM = object
sys.modules["matplotlib.path"].Path = object
sys.modules["matplotlib"].path = sys.modules["matplotlib.path"]

from modules.model import Generator, Discriminator, weights_init_normal  # noqa: E402  (reference)
import modules.preprocess as ref_pre  # noqa: E402  (reference)
import modules.trainer as ref_trainer  # noqa: E402  (reference)

from oracle import ducosy_oracle as orc  # noqa: E402


class FakeDcm:
    """Stand-in for a pydicom dataset: only the attributes the reference arithmetic reads."""

    def __init__(self, px, slope, intercept):
        self.pixel_array = px
        self.RescaleSlope = slope
        self.RescaleIntercept = intercept
        self.Rows, self.Columns = px.shape

    def __contains__(self, key):
        return hasattr(self, key)


def ref_lines(relpath, first, last):
    with open(os.path.join(REF, relpath)) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def rng_input(seed, shape):
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).uniform(-1, 1, size=shape).astype(np.float32))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())

    # ------------------------------------------------------------------ generator
    gen_cases = [
        # name, Cin, blocks, cbam, B, H, W, weight seed, input seed, attn_std
        ("gen_c1_b2_cbam_64", 1, 2, True, 1, 64, 64, 11, 101, 0.2),
        ("gen_c3_b1_plain_32", 3, 1, False, 2, 32, 32, 12, 102, None),
        ("gen_c2_b1_cbam_128", 2, 1, True, 1, 128, 128, 13, 103, 0.2),
    ]
    for name, cin, nb, cbam, B, H, W, wseed, xseed, astd in gen_cases:
        G = Generator(input_channels=cin, num_residual_blocks=nb, use_cbam=cbam)
        sd = orc.make_state_dict(orc.generator_param_shapes(cin, nb, cbam), wseed, attn_std=astd)
        G.load_state_dict(sd, strict=True)
        G.eval()
        x = rng_input(xseed, (B, cin, H, W))
        with torch.no_grad():
            y = G(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy(), cin=cin, blocks=nb, cbam=cbam,
                            B=B, H=H, W=W, wseed=wseed, xseed=xseed, attn_std=-1.0 if astd is None else astd,
                            keys=np.array(list(G.state_dict().keys())))
        print(name, tuple(y.shape), float(y.abs().mean()))

    # full-size (config 1/2 shape): 9 blocks, 512x512, Cin=1; keep a strided subset + float64 checksum
    G = Generator(input_channels=1, num_residual_blocks=9)
    sd = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), 1234, attn_std=0.2)
    G.load_state_dict(sd, strict=True)
    G.eval()
    px = orc.synthetic_volume(1, 512, 512, seed=5)[0]
    x = torch.from_numpy(orc.hu_window(px, 1.0, -1024.0, *orc.SOFT_HU).astype(np.float32))[None, None]
    with torch.no_grad():
        y = G(x)[0, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "gen_full_512.npz"), y_sub=y[::8, ::8].copy(),
                        y_sum=np.float64(y.astype(np.float64).sum()), y_abs_sum=np.float64(np.abs(y).astype(np.float64).sum()),
                        wseed=1234, vseed=5, attn_std=0.2)
    print("gen_full_512", y.shape, float(np.abs(y).mean()))

    # weights_init_normal touches exactly the Conv* modules (51 in G, 5 in D)
    n_conv_g = sum(1 for m in Generator(3).modules() if m.__class__.__name__.find("Conv") != -1)
    n_conv_d = sum(1 for m in Discriminator(1).modules() if m.__class__.__name__.find("Conv") != -1)

    # ------------------------------------------------------------------ discriminator
    D = Discriminator(1)
    sdd = orc.make_state_dict(orc.discriminator_param_shapes(1), 21)
    D.load_state_dict(sdd, strict=True)
    xd = rng_input(201, (2, 1, 64, 64))
    with torch.no_grad():
        yd = D(xd)
    np.savez_compressed(os.path.join(OUT, "disc_64.npz"), y=yd.numpy(), wseed=21, xseed=201,
                        keys=np.array(list(D.state_dict().keys())), n_conv_g=n_conv_g, n_conv_d=n_conv_d)
    print("disc_64", tuple(yd.shape))

    # ------------------------------------------------------------------ HU windowing / de-windowing
    edge_px = np.array([874, 873, 24, 23, 1274, 1275, 0, 1, 2499, 1024, 1023, 1224, 1223], dtype=np.int16)
    rnd_px = orc.synthetic_volume(1, 32, 32, seed=3)[0]
    rnd_px.flat[:edge_px.size] = edge_px
    cases = [(1.0, -1024.0), (2.0, -1000.0), (0.5, -512.25)]
    hu_out = {"px": rnd_px}
    for ci, (slope, intercept) in enumerate(cases):
        dcm = FakeDcm(rnd_px, slope, intercept)
        ref_pre.pydicom.dcmread = lambda path, _d=dcm: _d
        st, lu, _ = ref_pre.preprocess_dicom("x.dcm", -150, 250, -1000, -150)
        hu_out[f"win_soft_{ci}"] = st.numpy()[0]
        hu_out[f"win_lung_{ci}"] = lu.numpy()[0]
        # apply_hu_transform (training-side soft squeezing) and the linear variant
        hu_out[f"sq_soft_{ci}"] = ref_pre.apply_hu_transform(dcm, -150, 250, True)
        hu_out[f"sq_lung_{ci}"] = ref_pre.apply_hu_transform(dcm, -1000, -150, True)
        hu_out[f"lin_soft_{ci}"] = ref_pre.apply_hu_transform(dcm, -150, 250, False)
        # postprocess_tensor on a seeded tanh-range tensor (incl. exact +-1 and values that truncate)
        yt = rng_input(300 + ci, (1, 1, 32, 32))
        yt.view(-1)[:6] = torch.tensor([-1.0, 1.0, 0.0, -0.9965, 0.33333334, 0.9999999])
        hu_out[f"y_{ci}"] = yt.numpy()
        hu_out[f"post_soft_{ci}"] = ref_pre.postprocess_tensor(yt, dcm, -150, 250)
        hu_out[f"post_lung_{ci}"] = ref_pre.postprocess_tensor(yt, dcm, -1000, -150)
        # apply_windowing (display windowing, reference argmanager.py:126-127,143-144 window settings)
        hu_out[f"disp_soft_{ci}"] = ref_pre.apply_windowing(yt, types.SimpleNamespace(hu_min=-150, hu_max=250, window_center=40, window_width=400)).numpy()
        hu_out[f"disp_lung_{ci}"] = ref_pre.apply_windowing(yt, types.SimpleNamespace(hu_min=-1000, hu_max=-150, window_center=-600, window_width=1500)).numpy()
        hu_out[f"slope_{ci}"] = slope
        hu_out[f"intercept_{ci}"] = intercept
    np.savez_compressed(os.path.join(OUT, "hu_window.npz"), **hu_out)
    print("hu_window ok")

    # ------------------------------------------------------------------ composite (generate.py verbatim lines)
    get_hu_src = ref_lines("generate.py", 140, 145)
    comp_src = "\n".join([
        ref_lines("generate.py", 218, 218),     # merged = raw.copy()
        ref_lines("generate.py", 221, 221),     # raw_hu_array = get_hu_array(raw_dcm)
        ref_lines("generate.py", 224, 227),     # soft mask
        ref_lines("generate.py", 230, 233),     # lung mask
        ref_lines("generate.py", 236, 237),     # overwrites
    ])
    comp_out = {}
    for ci, (slope, intercept) in enumerate(cases):
        raw = orc.synthetic_volume(1, 48, 48, seed=40 + ci)[0]
        raw.flat[:edge_px.size] = edge_px
        g = np.random.Generator(np.random.PCG64(50 + ci))
        soft_px = g.integers(-2000, 4000, size=raw.shape, dtype=np.int16)
        lung_px = g.integers(-2000, 4000, size=raw.shape, dtype=np.int16)
        env = {"np": np, "raw_dcm": FakeDcm(raw, slope, intercept), "raw_pixel_array": raw,
               "soft_tissue_pixel_array": soft_px, "lung_pixel_array": lung_px,
               "soft_tissue_args": types.SimpleNamespace(hu_min=-150, hu_max=250),
               "lung_args": types.SimpleNamespace(hu_min=-1000, hu_max=-150)}
        exec(get_hu_src, env)
        exec(comp_src, env)
        comp_out.update({f"raw_{ci}": raw, f"soft_px_{ci}": soft_px, f"lung_px_{ci}": lung_px,
                         f"merged_{ci}": env["merged_pixel_array"], f"soft_mask_{ci}": env["soft_tissue_mask"],
                         f"lung_mask_{ci}": env["lung_mask"], f"slope_{ci}": slope, f"intercept_{ci}": intercept})
    # tags absent -> slope 1 / intercept 0 defaults (generate.py:142-143)
    class NoTags:
        def __init__(self, px): self.pixel_array = px
        def __contains__(self, key): return False
    env = {"np": np}
    exec(get_hu_src, env)
    comp_out["hu_notags"] = env["get_hu_array"](NoTags(edge_px))
    comp_out["edge_px"] = edge_px
    np.savez_compressed(os.path.join(OUT, "composite.npz"), **comp_out)
    print("composite ok")

    # ------------------------------------------------------------------ HU threshold candidates (mask_generator verbatim lines)
    hu = (orc.synthetic_volume(1, 40, 40, seed=60)[0].astype(np.float32) - 1024.0)
    hu.flat[:8] = np.array([-1000, -1000.5, -999.5, -300, -299.5, 200, 199.5, 3000], np.float32)
    env = {"np": np, "hu_volume": hu, "lung_lower": -1000, "lung_upper": -300, "bone_threshold": 200}
    exec(ref_lines("modules/mask_generator.py", 14, 20), env)
    lung_cand = env["lung_mask"].copy()
    body = env["body_mask"].copy()
    exec(ref_lines("modules/mask_generator.py", 179, 183), env)
    np.savez_compressed(os.path.join(OUT, "thresholds.npz"), hu=hu, body=body, lung=lung_cand,
                        bone=env["all_bone_candidate"])
    print("thresholds ok")

    # ------------------------------------------------------------------ losses (reference classes, unmodified)
    p, t, s = rng_input(401, (2, 1, 64, 64)), rng_input(402, (2, 1, 64, 64)), rng_input(403, (2, 1, 64, 64))
    losses = {
        "grad": ref_trainer.GradientLoss()(p, t).item(),
        "att": ref_trainer.ContrastAttentionLoss(sigma=0.15, min_weight=1.0, max_weight=3.0, blur_kernel=7)(p, t, s).item(),
        "region": ref_trainer.ContrastRegionLoss(threshold=0.15, weight=1.5)(p, t, s).item(),
        "edge": ref_trainer.ContrastEdgeLoss()(p, t, s).item(),
    }
    np.savez_compressed(os.path.join(OUT, "losses.npz"), seeds=np.array([401, 402, 403]), **losses)
    print("losses", losses)


if __name__ == "__main__":
    main()
