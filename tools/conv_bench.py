"""Micro-benchmark of the residual-block convolution (3x3, 256->256, 128x128) through the C ABI.
usage: python tools/conv_bench.py [B] ; env DUCOSY_CONV_CTA_GROUP=1|2"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ducosy_gan_b200 import ops

def run(B, stats, iters=30):
    dt = torch.float16
    x = torch.randn((B, 130, 130, 256), device="cuda").to(dt)
    w = ops.pack_conv_weight(torch.randn((256, 256, 3, 3), device="cuda") * 0.02, dt)
    for _ in range(5):
        ops.conv2d_nhwc(x, w, 3, 3, 1, want_stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv2d_nhwc(x, w, 3, 3, 1, want_stats=stats)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tf = 2.0 * B * 128 * 128 * 256 * 2304 / us / 1e6
    return us, tf

if __name__ == "__main__":
    Bs = [int(a) for a in sys.argv[1:]] or [10]
    for B in Bs:
        for stats in (True, False):
            us, tf = run(B, stats)
            print(json.dumps({"B": B, "stats": stats, "cg": os.environ.get("DUCOSY_CONV_CTA_GROUP", "2"),
                              "us": round(us, 1), "tflops": round(tf, 1)}), flush=True)
