"""Kernel timeline of the dual-HU synthesizer via CUPTI (torch.profiler): how much the two generator streams overlap.
Writes gpurun_out/timeline.json (summary) -- run on the GPU box."""
import json, os, sys, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ducosy_gan_b200.synthesis import DualHUSynthesizer

def main():
    dev = torch.device("cuda", 0)
    B = int(os.environ.get("DUCOSY_BATCH_SLICES", "10"))
    soft, lung = bench.make_models(dev)
    synth = DualHUSynthesizer(soft, lung, batch_slices=B, device=dev)
    vol = torch.from_numpy(bench.synthetic_volume(0)[: 6 * B]).to(dev)
    out = torch.empty_like(vol)
    for _ in range(3):
        synth.synthesize_device(vol, 1.0, -1024.0, out=out)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        synth.synthesize_device(vol, 1.0, -1024.0, out=out)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    wall = t1 - t0
    total = sum(k[1] - k[0] for k in ks)
    # union of busy intervals
    busy, cur_s, cur_e = 0.0, None, None
    for s, e, _ in ks:
        if cur_e is None or s > cur_e:
            if cur_e is not None: busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    per = collections.defaultdict(float); cnt = collections.Counter()
    for s, e, n in ks:
        import re
        mm = re.search(r"(\w+_kernel)(<[^(]*>)?", n)
        n = (mm.group(1) + (mm.group(2) or "")) if mm else n[:60]
        per[n] += e - s; cnt[n] += 1
    summary = {"kernels": len(ks), "wall_us": wall, "sum_kernel_us": total, "busy_union_us": busy,
               "idle_us": wall - busy, "overlap_factor": total / busy,
               "per_kernel_us": {k: [round(v, 1), cnt[k], round(v / cnt[k], 1)] for k, v in sorted(per.items(), key=lambda kv: -kv[1])}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(summary, open(os.path.join(ROOT, "gpurun_out", "timeline.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "per_kernel_us"}))
    for k, v in summary["per_kernel_us"].items():
        print(f"{v[0]:10.1f} us n={v[1]:4d} avg={v[2]:8.1f}  {k[:80]}")

if __name__ == "__main__":
    main()
