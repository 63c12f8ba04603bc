"""Where a data-parallel CycleGAN step's time goes on the GPU timeline (run under torchrun, one process per GPU):
global batch = WORLD_SIZE x --per-rank-batch, the step replayed from its CUDA graph under torch.profiler (CUPTI kernel records);
rank 0 reports, per step: wall time, the union of all kernel intervals (GPU busy), idle time, summed kernel time split into NCCL /
this library / torch, and the NCCL share of the busy time.   torchrun --nproc-per-node N tools/train_timeline.py --out X.json"""
import argparse, json, os, re, sys, tempfile
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--per-rank-batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--out", default="")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from ducosy_gan_b200.data_parallel import DataParallelCycleGANStep, GraphedCycleGANStep  # noqa: E402

step = DataParallelCycleGANStep(3, 9, True, seed=1234, device=dev, capturable=True)
g = torch.Generator().manual_seed(2 + rank)
b = a.per_rank_batch
real_A = (torch.rand(b, 1, 512, 512, generator=g) * 2 - 1).to(dev)
real_B = (torch.rand(b, 1, 512, 512, generator=g) * 2 - 1).to(dev)
masks = (torch.rand(b, 2, 512, 512, generator=g) < 0.1).float().to(dev)
graphed = GraphedCycleGANStep(step, real_A, real_B, masks, warmup=2)
for _ in range(3):
    graphed(real_A, real_B, masks)
torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[dev.index])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    graphed(real_A, real_B, masks)
e1.record()
torch.cuda.synchronize()
ms_plain = e0.elapsed_time(e1) / a.steps

from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        graphed(real_A, real_B, masks)
    torch.cuda.synchronize()
if rank == 0:
    path = os.path.join(tempfile.gettempdir(), "ducosy_timeline_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    busy, cur_s, cur_e = 0.0, None, None
    for e in ev:                                   # union of the kernel intervals over all streams
        s, t = e["ts"], e["ts"] + e["dur"]
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, t
        else:
            cur_e = max(cur_e, t)
    busy += cur_e - cur_s
    cat = {"nccl": 0.0, "ducosy": 0.0, "torch_and_other": 0.0}
    count = {"nccl": 0, "ducosy": 0, "torch_and_other": 0}
    for e in ev:
        n = e["name"]
        k = "nccl" if "nccl" in n.lower() else ("ducosy" if "ducosy" in n else "torch_and_other")
        cat[k] += e["dur"]
        count[k] += 1
    n = a.steps
    res = {"what": "GPU timeline of the graph-replayed data-parallel CycleGAN step on rank 0 (torch.profiler / CUPTI kernel records)",
           "world_size": world, "per_rank_batch": b, "global_batch": b * world, "steps_profiled": n,
           "ms_per_step_without_profiler": ms_plain,
           "ms_per_step_span_under_profiler": (t1 - t0) / 1e3 / n,
           "gpu_busy_ms_per_step (union of kernel intervals, all streams)": busy / 1e3 / n,
           "gpu_idle_ms_per_step": ((t1 - t0) - busy) / 1e3 / n,
           "summed_kernel_ms_per_step": {k: v / 1e3 / n for k, v in cat.items()},
           "launches_per_step": {k: v / n for k, v in count.items()},
           "nccl_share_of_busy_time": cat["nccl"] / busy}
    print(json.dumps(res, indent=1))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
graphed.close()
if world > 1:
    dist.barrier(device_ids=[dev.index])
    dist.destroy_process_group()
