"""BASELINE.json config 5: CBAM residual-block throughput sweep, batch 1..64 at 256 channels x 128x128,
fp16 vs bf16 operands (fp32 accumulation in TMEM in both cases), and the split-operand arm fp16x2 ((hi, lo) fp16 pairs, three
tensor-core products per tap: the fp32-class / <= 1 HU mode).  One block = model.py:68-87:
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


def block_split(x_pad, w1, w2, fc0, fc2, wsa):
    """The same block in split-operand mode through the C ABI: buffers hold 2*C 16-bit channels per pixel (hi plane, lo plane)."""
    from ducosy_gan_b200._lib import F16X2, call, ptr, stream_ptr
    B, Hp, Wp, C2 = x_pad.shape
    H, W, C = Hp - 2, Wp - 2, C2 // 2
    dev = x_pad.device
    e = lambda *shape, dt=torch.float16: torch.empty(shape, dtype=dt, device=dev)
    y, mid, out = e(B, H, W, C2), e(B, Hp, Wp, C2), e(B, Hp, Wp, C2)
    part = e(B, H * W // 128, 3, C, dt=torch.float32)
    sc, sh, chmax = (e(B, C, dt=torch.float32) for _ in range(3))
    pooled, sa = e(B, H, W, 2, dt=torch.float32), e(B, H, W, dt=torch.float32)
    st = stream_ptr()
    call("ducosy_conv2d_nhwc", ptr(x_pad), ptr(w1), ptr(y), ptr(part), None, 0, B, Hp, Wp, C, C, 3, 3, 1, F16X2, st)
    call("ducosy_in_finalize", ptr(part), H * W // 128, H * W, ptr(sc), ptr(sh), None, None, None, B, C, st)
    call("ducosy_in_apply_pad", ptr(y), ptr(sc), ptr(sh), ptr(mid), B, H, W, C, 1, ops.PAD_REFLECT, ops.ACT_RELU, F16X2, st)
    call("ducosy_conv2d_nhwc", ptr(mid), ptr(w2), ptr(y), ptr(part), None, 0, B, Hp, Wp, C, C, 3, 3, 1, F16X2, st)
    call("ducosy_in_finalize", ptr(part), H * W // 128, H * W, ptr(sc), ptr(sh), ptr(fc0), ptr(fc2), ptr(chmax), B, C, st)
    call("ducosy_cbam_pool", ptr(y), ptr(sc), ptr(sh), ptr(pooled), B, H, W, C, F16X2, st)
    call("ducosy_cbam_spatial_conv", ptr(pooled), ptr(wsa), ptr(sa), B, H, W, st)
    call("ducosy_residual_apply_pad", ptr(y), ptr(sc), ptr(sh), ptr(sa), ptr(x_pad), 1, ptr(out), B, H, W, C, 1, ops.PAD_REFLECT, F16X2, st)
    return out


def pack_split(w):
    from ducosy_gan_b200._lib import F16X2, call, ptr, stream_ptr
    out = torch.empty((w.shape[0], 2 * 9 * w.shape[1]), dtype=torch.float16, device=w.device)
    call("ducosy_pack_conv_weight", ptr(w.contiguous()), ptr(out), w.shape[0], w.shape[1], 3, 3, F16X2, stream_ptr())
    return out


def main():
    out = []
    for dt, name in ((torch.float16, "fp16"), (torch.bfloat16, "bf16"), (torch.float16, "fp16x2")):
        split = name == "fp16x2"
        run = block_split if split else block
        pk = pack_split if split else (lambda w: ops.pack_conv_weight(w, dt))
        w1 = pk(torch.randn(256, 256, 3, 3, device="cuda") * 0.02)
        w2 = pk(torch.randn(256, 256, 3, 3, device="cuda") * 0.02)
        fc0 = (torch.randn(16, 256, device="cuda") * 0.1).contiguous()
        fc2 = (torch.randn(256, 16, device="cuda") * 0.1).contiguous()
        wsa = (torch.randn(1, 2, 7, 7, device="cuda") * 0.1).contiguous()
        for B in (1, 2, 4, 8, 16, 32, 64):
            x = torch.randn(B, 130, 130, 256, device="cuda").to(dt)
            if split:   # (hi, lo) planes; a random lo plane of the right magnitude times like a real one
                x = torch.cat([x, (torch.randn(B, 130, 130, 256, device="cuda") * 2e-4).to(dt)], dim=-1).contiguous()
            for _ in range(3):
                run(x, w1, w2, fc0, fc2, wsa)
            torch.cuda.synchronize()
            iters = max(4, 64 // B)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                run(x, w1, w2, fc0, fc2, wsa)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            rec = {"operands": name, "accumulate": "fp32 (TMEM)" + (", 3 products per tap" if split else ""), "batch": B, "us_per_block": round(us, 1),
                   "samples_per_s": round(B / us * 1e6, 1), "conv_tflops": round(B * GF_PER_SAMPLE * 1e3 / us, 1)}
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del x
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "resblock_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
