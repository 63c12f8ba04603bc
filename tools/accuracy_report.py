"""Accuracy report of the CUDA generator against the fp32 oracle (run on the GPU box).
Writes gpurun_out/accuracy.json.  Also evaluates a rounding-aware oracle (activations / weights rounded to the
16-bit operand type at the points where the CUDA path stores them) to separate quantisation noise from bugs."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ducosy_oracle as orc  # noqa: E402


rounded_oracle = lambda sd, x, nb, cbam, dt: orc.generator_forward_rounded(sd, x, nb, cbam, dt)


def stats(a, b):
    d = (a - b).abs().flatten().double()
    return {"max": d.max().item(), "mean": d.mean().item(), "rms": d.pow(2).mean().sqrt().item(),
            "p999": torch.quantile(d[:: max(1, d.numel() // 1_000_000)], 0.999).item()}


def main():
    from ducosy_gan_b200.modules.model import Generator
    out = []
    cases = [(1, 0, True, 1, 128, 128), (1, 1, False, 1, 128, 128), (1, 2, True, 1, 128, 128), (1, 9, True, 1, 256, 256),
             (1, 9, True, 1, 512, 512)]
    for prec, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16), ("fp16x2", None)):
        os.environ["DUCOSY_PRECISION"] = prec
        for cin, nb, cbam, B, H, W in cases:
            sd = orc.make_state_dict(orc.generator_param_shapes(cin, nb, cbam), 1234, attn_std=0.2)
            G = Generator(cin, nb, cbam)
            G.load_state_dict(sd)
            G = G.cuda().eval()
            px = orc.synthetic_volume(B, H, W, seed=5)
            x = torch.from_numpy(orc.hu_window(px, 1.0, -1024.0, *orc.SOFT_HU).astype(np.float32))[:, None]
            with torch.no_grad():
                y = G(x.cuda()).cpu()
                ref = orc.generator_forward(sd, x, nb, cbam)
                rref = rounded_oracle(sd, x, nb, cbam, dt) if dt is not None else None
            rec = {"precision": prec, "cin": cin, "blocks": nb, "cbam": cbam, "H": H, "W": W, "cuda_vs_fp32_oracle": stats(y, ref)}
            if rref is not None:
                rec.update({"cuda_vs_rounded_oracle": stats(y, rref), "rounded_vs_fp32_oracle": stats(rref, ref)})
            else:
                # split-operand arm: the fp32 oracle's own rounding is of the same order as the error, so also compare both
                # against the oracle evaluated in float64
                with torch.no_grad():
                    ref64 = orc.generator_forward({k: v.double() for k, v in sd.items()}, x.double(), nb, cbam)
                rec.update({"cuda_vs_fp64_oracle": stats(y.double(), ref64), "fp32_oracle_vs_fp64_oracle": stats(ref.double(), ref64)})
                for name, (lo, hi) in (("soft", orc.SOFT_HU), ("lung", orc.LUNG_HU)):
                    a = orc.dewindow_to_stored(y[:, 0].numpy(), 1.0, -1024.0, lo, hi).astype(np.int32)
                    b = orc.dewindow_to_stored(ref[:, 0].numpy(), 1.0, -1024.0, lo, hi).astype(np.int32)
                    rec[f"dewindowed_{name}_max_abs_diff_stored_units"] = int(np.abs(a - b).max())
                    rec[f"dewindowed_{name}_fraction_differing"] = float((a != b).mean())
            rec["max_err_HU_soft"] = rec["cuda_vs_fp32_oracle"]["max"] * 200
            rec["mean_err_HU_soft"] = rec["cuda_vs_fp32_oracle"]["mean"] * 200
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "accuracy.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
