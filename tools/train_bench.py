"""CycleGAN train-step throughput (BASELINE config 4: batch 8, 512x512, soft-tissue generators with 2 mask channels).
usage: python tools/train_bench.py [--batch 8] [--cin 3] [--steps 5] [--warmup 2] [--profile] [--out profiles/x.json]"""
import argparse
import re
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ducosy_gan_b200.trainer import CycleGANStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--cin", type=int, default=3)
ap.add_argument("--blocks", type=int, default=9)
ap.add_argument("--no-cbam", action="store_true")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--graph", action="store_true", help="replay the whole step from a CUDA graph (what bench.py times)")
ap.add_argument("--out", default="")
a = ap.parse_args()

if a.graph:
    from ducosy_gan_b200.data_parallel import DataParallelCycleGANStep, GraphedCycleGANStep  # noqa: E402
    step = DataParallelCycleGANStep(a.cin, a.blocks, not a.no_cbam, seed=1234, capturable=True)
else:
    step = CycleGANStep(a.cin, a.blocks, not a.no_cbam, seed=1234)
g = torch.Generator().manual_seed(2)
B = a.batch
real_A = (torch.rand(B, 1, 512, 512, generator=g) * 2 - 1).cuda()
real_B = (torch.rand(B, 1, 512, 512, generator=g) * 2 - 1).cuda()
masks = (torch.rand(B, a.cin - 1, 512, 512, generator=g) < 0.1).float().cuda() if a.cin > 1 else None
if a.graph:
    graphed = GraphedCycleGANStep(step, real_A, real_B, masks, warmup=max(a.warmup, 2))
    step_fn = lambda: graphed(real_A, real_B, masks)
    step_fn()
else:
    step_fn = lambda: step.step(real_A, real_B, masks)
    for _ in range(a.warmup):
        out = step_fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = step_fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
# nominal conv work per step: 6 generator passes x 3 (fwd + 2x bwd) + 6 discriminator passes x 3
gflop = {1: 447.82, 2: 449.46, 3: 451.11}[a.cin] * a.blocks / 9 if a.blocks != 9 else {1: 447.82, 2: 449.46, 3: 451.11}[a.cin]
tflop = (6 * B * gflop * 3 + 6 * B * 13.04 * 3) / 1e3
res = {"config": f"CycleGAN step, batch {B}, Cin {a.cin}, {a.blocks} blocks, cbam {not a.no_cbam}, 512x512", "ms_per_step": ms,
       "steps_per_s": 1e3 / ms, "samples_per_s": B * 1e3 / ms, "nominal_tflop_per_step": tflop, "tflops_nominal": tflop / ms * 1e3 / 1e3,
       "losses": {k: float(v) for k, v in out.items()}, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
print(json.dumps(res, indent=1))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
if a.profile:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step_fn()
        torch.cuda.synchronize()
    agg = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            m = re.search(r"ducosy::(?:\(anonymous namespace\)::)?(\w+)", ev.name)
            n = m.group(1) if m else ev.name.split("<")[0].split("(")[0]
            d = agg.setdefault(n, [0, 0.0])
            d[0] += 1
            d[1] += ev.device_time
    if os.environ.get("LIST_KERNEL"):
        evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA
                      and os.environ["LIST_KERNEL"] in ev.name), key=lambda e: e.time_range.start)
        print(os.environ["LIST_KERNEL"], "durations (us), launch order:", [round(e.device_time) for e in evs][:120])
    tot = sum(v[1] for v in agg.values())
    print(f"kernel time total {tot / 1e3:.2f} ms")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{c:<5d} {n[:90]}")
