"""BASELINE.json config 5: CBAM residual-block throughput sweep, batch 1..64 at 256 channels x 128x128,
fp16 vs bf16 operands (fp32 accumulation in TMEM in both cases).  One block = model.py:68-87:
x + CBAM(IN(conv(refpad(ReLU(IN(conv(refpad(x)))))))) through the C-ABI kernels.  Writes gpurun_out/resblock_sweep.json."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ducosy_gan_b200 import ops

GF_PER_SAMPLE = 38.65


def block(x_pad, w1, w2, fc0, fc2, wsa):
    B, Hp, Wp, C = x_pad.shape
    H, W = Hp - 2, Wp - 2
    y1, p1 = ops.conv2d_nhwc(x_pad, w1, 3, 3, 1)
    sc, sh = ops.in_finalize(p1, H * W)
    mid = ops.in_apply_pad(y1, sc, sh, 1, ops.PAD_REFLECT, ops.ACT_RELU)
    y2, p2 = ops.conv2d_nhwc(mid, w2, 3, 3, 1)
    sc, sh = ops.in_finalize(p2, H * W, fc0, fc2)
    sa = ops.cbam_spatial_conv(ops.cbam_pool(y2, sc, sh), wsa)
    return ops.residual_apply_pad(y2, sc, sh, sa, x_pad, 1, 1, ops.PAD_REFLECT)


def main():
    out = []
    for dt, name in ((torch.float16, "fp16"), (torch.bfloat16, "bf16")):
        w1 = ops.pack_conv_weight(torch.randn(256, 256, 3, 3, device="cuda") * 0.02, dt)
        w2 = ops.pack_conv_weight(torch.randn(256, 256, 3, 3, device="cuda") * 0.02, dt)
        fc0 = (torch.randn(16, 256, device="cuda") * 0.1).contiguous()
        fc2 = (torch.randn(256, 16, device="cuda") * 0.1).contiguous()
        wsa = torch.randn(1, 2, 7, 7, device="cuda") * 0.1
        for B in (1, 2, 4, 8, 16, 32, 64):
            x = torch.randn(B, 130, 130, 256, device="cuda").to(dt)
            for _ in range(3):
                block(x, w1, w2, fc0, fc2, wsa)
            torch.cuda.synchronize()
            iters = max(4, 64 // B)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                block(x, w1, w2, fc0, fc2, wsa)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            rec = {"operands": name, "accumulate": "fp32 (TMEM)", "batch": B, "us_per_block": round(us, 1),
                   "samples_per_s": round(B / us * 1e6, 1), "conv_tflops": round(B * GF_PER_SAMPLE * 1e3 / us, 1)}
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del x
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "resblock_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
