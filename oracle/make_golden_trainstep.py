"""Golden vector for the CycleGAN optimisation step: executes the REFERENCE's own statements --
modules/trainer.py:347-362 (criteria, optimisers) and :448-525 (the loop body) via exec() on the cited line ranges -- on
the reference's own Generator / Discriminator modules, for two consecutive iterations on a small seeded batch, and stores
every logged loss in tests/golden/train_step.npz.  tests/test_oracle_golden.py replays oracle.cyclegan_step against it.

pytorch_msssim is absent (and unpinned in the reference): the SSIM criterion the loop calls is a stub that evaluates the
oracle's restatement (PARITY UNPINNED for that one term, as everywhere else).  TEST INFRASTRUCTURE ONLY.
usage: python oracle/make_golden_trainstep.py      (authoring container, needs /root/reference)"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
from oracle import ducosy_oracle as orc  # noqa: E402

for name in ("pydicom", "pytorch_msssim", "matplotlib", "matplotlib.path"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)


class _SSIMStub(torch.nn.Module):          # stands in for pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=1)
    def __init__(self, data_range=1.0, size_average=True, channel=1):
        super().__init__()
        self.data_range = data_range

    def forward(self, x, y):
        return orc.ssim(x, y, self.data_range)


sys.modules["pytorch_msssim"].SSIM = _SSIMStub
sys.modules["matplotlib.path"].Path = object
sys.modules["matplotlib"].path = sys.modules["matplotlib.path"]

from modules.model import Discriminator, Generator  # noqa: E402  (reference)
import modules.trainer as ref_trainer  # noqa: E402  (reference)


def ref_lines(relpath, first, last):
    with open(os.path.join(REF, relpath)) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


CFG = dict(cin=2, blocks=1, cbam=True, B=1, size=256, seeds=(11, 12, 13, 14), batch_seed=5, steps=2)


def make_batch(cfg):
    g = torch.Generator().manual_seed(cfg["batch_seed"])
    smooth = lambda t: torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0
    n = cfg["size"]
    A = smooth(torch.rand(cfg["B"], 1, n, n, generator=g) * 2 - 1).clamp(-1, 1)
    Bt = smooth(torch.rand(cfg["B"], 1, n, n, generator=g) * 2 - 1).clamp(-1, 1)
    M = (torch.rand(cfg["B"], cfg["cin"] - 1, n, n, generator=g) < 0.1).float()
    return A, Bt, M


def main():
    cfg = CFG
    A, Bt, M = make_batch(cfg)
    gshapes = orc.generator_param_shapes(cfg["cin"], cfg["blocks"], cfg["cbam"])
    dshapes = orc.discriminator_param_shapes(1)
    G_A2B, G_B2A = (Generator(cfg["cin"], cfg["blocks"], cfg["cbam"]) for _ in range(2))
    D_A, D_B = Discriminator(1), Discriminator(1)
    for m, shapes, seed in ((G_A2B, gshapes, cfg["seeds"][0]), (G_B2A, gshapes, cfg["seeds"][1]),
                            (D_A, dshapes, cfg["seeds"][2]), (D_B, dshapes, cfg["seeds"][3])):
        m.load_state_dict(orc.make_state_dict(shapes, seed), strict=True)
    args = types.SimpleNamespace(lr=2e-4, img_size=cfg["size"], lambda_cyc=10.0, lambda_id=5.0)
    env = {"torch": torch, "args": args, "device": torch.device("cpu"), "G_A2B": G_A2B, "G_B2A": G_B2A, "D_A": D_A, "D_B": D_B,
           "GradientLoss": ref_trainer.GradientLoss, "SSIM": _SSIMStub, "ContrastAttentionLoss": ref_trainer.ContrastAttentionLoss,
           "ContrastRegionLoss": ref_trainer.ContrastRegionLoss, "ContrastEdgeLoss": ref_trainer.ContrastEdgeLoss,
           "batch": {"A": A, "B": Bt, "masks": M}}
    exec(ref_lines("modules/trainer.py", 347, 362), env)          # criteria + the three Adam optimisers
    body = ref_lines("modules/trainer.py", 448, 525)              # real_A, real_B = ... through optimizer_D_B.step()
    names = ["loss_G", "loss_GAN", "loss_cycle", "loss_id", "loss_grad_cycle", "loss_grad_id", "loss_ssim",
             "loss_contrast_attention", "loss_contrast_region", "loss_contrast_edge", "loss_D_A", "loss_D_B"]
    hist = {n: [] for n in names}
    for _ in range(cfg["steps"]):
        exec(body, env)
        for n in names:
            hist[n].append(float(env[n].detach()))
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), cin=cfg["cin"], blocks=cfg["blocks"], cbam=cfg["cbam"], B=cfg["B"],
                        size=cfg["size"], seeds=np.array(cfg["seeds"]), batch_seed=cfg["batch_seed"],
                        **{n: np.array(v, dtype=np.float64) for n, v in hist.items()})
    print({n: v for n, v in hist.items()})


if __name__ == "__main__":
    main()
