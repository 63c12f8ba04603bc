#!/usr/bin/env python
"""bench.py -- dual-HU synthesis throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): soft-tissue + lung A2B generators (input_channels=1, 9 CBAM residual
blocks, seeded random weights) + de-window + complementary composite over a synthetic 300-slice 512x512
NCCT volume.  One step = one pass over the volume.  With N ranks every rank synthesizes its own 300-slice
volume (independent slices, no data-path collective) => weak scaling; value = all slices / max-over-ranks time.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, H, W = 300, 512, 512
SLOPE, INTERCEPT = 1.0, -1024.0
GFLOP_PER_SLICE = 2 * 447.82            # two generators, Cin = 1, nominal conv FLOPs
METRIC = "dual_hu_synth_slices_per_sec"
UNIT = "slices/s"
WORKLOAD = "dual-HU generate: soft-tissue + lung Generator(Cin=1, 9 CBAM blocks) + composite, 300x512x512 int16 volume per GPU"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_models(device, seed=1234):
    """Seeded random-init generators (no shipped checkpoints: reference .gitignore:220-221)."""
    from ducosy_gan_b200.modules.model import Generator, weights_init_normal
    torch.manual_seed(seed)
    soft = Generator(input_channels=1, num_residual_blocks=9)      # generate.py:29-30
    lung = Generator(input_channels=1, num_residual_blocks=9)
    soft.apply(weights_init_normal)
    lung.apply(weights_init_normal)
    return soft.to(device).eval(), lung.to(device).eval()


def synthetic_volume(seed):
    g = np.random.Generator(np.random.PCG64(seed))
    return g.integers(0, 2500, size=(S, H, W), dtype=np.int16)     # HU in [-1024, 1475] at slope 1 / intercept -1024


def time_dominant_kernel(device, batch, iters=20):
    """CUDA-event timing of the dominant kernel alone: the 3x3 256->256 residual-block convolution
    (tcgen05 implicit GEMM) at the batch the synthesizer uses.  Returns (avg seconds, FLOPs per launch)."""
    from ducosy_gan_b200 import ops
    dt = torch.float16 if os.environ.get("DUCOSY_PRECISION", "fp16") == "fp16" else torch.bfloat16
    x = torch.randn((batch, 130, 130, 256), device=device).to(dt)
    w = ops.pack_conv_weight(torch.randn((256, 256, 3, 3), device=device) * 0.02, dt)
    for _ in range(3):
        ops.conv2d_nhwc(x, w, 3, 3, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv2d_nhwc(x, w, 3, 3, 1)
    e1.record()
    torch.cuda.synchronize()
    flops = 2.0 * batch * 128 * 128 * 256 * 256 * 9
    return e0.elapsed_time(e1) / 1e3 / iters, flops


def cpu_baseline_sample(n_slices=3):
    """Oracle (CPU restatement of the reference path) timed on the host cores on a bounded sample."""
    from oracle import ducosy_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    sd_s = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), 1234)
    sd_l = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), 1235)
    vol = synthetic_volume(0)[: n_slices + 1]
    orc.dual_hu_synthesize(vol[:1], SLOPE, INTERCEPT, sd_s, sd_l)          # warm-up slice
    t0 = time.perf_counter()
    orc.dual_hu_synthesize(vol[1:], SLOPE, INTERCEPT, sd_s, sd_l)
    dt = time.perf_counter() - t0
    return {"value": n_slices / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_slices} of the {S} slices (512x512), fp32 torch-CPU oracle of the same path, after 1 warm-up slice"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot travel to the box)."""
    if rank != 0:
        return
    from oracle import ducosy_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    sd_s = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), 1234)
    sd_l = orc.make_state_dict(orc.generator_param_shapes(1, 9, True), 1235)
    n = 2                                                                    # slices per step (bounded sample)
    vol = synthetic_volume(0)[:n]
    for _ in range(min(args.warmup, 1)):
        orc.dual_hu_synthesize(vol[:1], SLOPE, INTERCEPT, sd_s, sd_l)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.dual_hu_synthesize(vol, SLOPE, INTERCEPT, sd_s, sd_l)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    cores = torch.get_num_threads()
    sample = f"{n} slices (512x512) per step of the {S}-slice volume; fp32 torch-CPU oracle port of the reference path"
    _emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def train_metric(device, rank, world, steps, warmup=2):
    """Second half of BASELINE.json's metric: CycleGAN train steps/s (config 4: batch 8, 512x512, soft-tissue generators
    with two mask channels, all losses, three Adam steps).  The global batch of 8 is sharded over the ranks with a NCCL
    gradient all-reduce (strong scaling).  Every step copies its batch from pinned host memory and reads the loss back."""
    import torch.distributed as dist
    from ducosy_gan_b200.data_parallel import DataParallelCycleGANStep, GraphedCycleGANStep, shard_batch
    B, cin = 8, 3
    if world > B:
        return None
    lo, hi = shard_batch(B, rank, world)
    g = torch.Generator().manual_seed(2)
    host = [(torch.rand(B, 1, H, W, generator=g) * 2 - 1)[lo:hi].contiguous().pin_memory() for _ in range(2)]
    host.append((torch.rand(B, cin - 1, H, W, generator=g) < 0.1).float()[lo:hi].contiguous().pin_memory())
    step = DataParallelCycleGANStep(cin, 9, True, seed=1234, device=device, capturable=True)
    # the whole step (~3000 launches, 3 all-reduces) is captured once in a CUDA graph; the warm-up steps are the capture's
    graphed = GraphedCycleGANStep(step, *(t.to(device) for t in host), warmup=max(warmup, 2))

    def one():
        return graphed(*host)["G"].item()      # host -> static device buffers (pinned, async), replay, loss read-back

    one()
    if world > 1:
        dist.barrier(device_ids=[device.index])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    graphed.close()      # the graph holds captured NCCL collectives: it must be gone before the process group is torn down
    tflop = (6 * B * 451.11 * 3 + 6 * B * 13.04 * 3) / 1e3      # nominal conv work, backward = 2x forward
    return {"metric": "cyclegan_train_steps_per_s", "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": steps,
            "warmup": max(warmup, 2) + 1, "global_batch": B, "scaling": "strong", "loss_G": loss, "cuda_graph": True,
            "config": "G_A2B/G_B2A (Cin 3, 9 CBAM blocks) + D_A/D_B, 512x512, all 9 loss terms, 3 fused Adam steps; "
                      f"batch 8 sharded x{world}" + (", NCCL gradient all-reduce" if world > 1 else ""),
            "h2d_bytes_per_step": sum(t.numel() * 4 for t in host), "d2h_bytes_per_step": 4,
            "nominal_tflop_per_step": tflop, "achieved_tflops_nominal": tflop / ms * 1e3 / world,
            "frac_of_sustained_peak_per_gpu": tflop / ms * 1e3 / world / peaks()["tf_sustained"]}


def cpu_train_baseline():
    """The oracle's restatement of the reference loop body (pinned to modules/trainer.py:448-525 by tests/golden/train_step.npz)
    on the host cores: one optimisation step at batch 1 (batch 8 needs ~100 GB of autograd state)."""
    from oracle import ducosy_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    mk = lambda shapes, seed: {k: v.clone().requires_grad_(True) for k, v in orc.make_state_dict(shapes, seed).items()}
    gsh, dsh = orc.generator_param_shapes(3, 9, True), orc.discriminator_param_shapes(1)
    sds = (mk(gsh, 1), mk(gsh, 2), mk(dsh, 3), mk(dsh, 4))
    adam = lambda ps: torch.optim.Adam(ps, lr=2e-4, betas=(0.5, 0.999))
    opts = (adam(list(sds[0].values()) + list(sds[1].values())), adam(list(sds[2].values())), adam(list(sds[3].values())))
    g = torch.Generator().manual_seed(2)
    a, b = (torch.rand(1, 1, H, W, generator=g) * 2 - 1 for _ in range(2))
    m = (torch.rand(1, 2, H, W, generator=g) < 0.1).float()
    t0 = time.perf_counter()
    orc.cyclegan_step(sds, opts, a, b, m, 9, True)
    sec = time.perf_counter() - t0
    return {"value": 1.0 / sec, "unit": "steps/s at batch 1", "samples_per_s": 1.0 / sec, "cores": os.cpu_count(), "kind": "port",
            "sample": "one optimisation step at batch 1 (512x512, Cin 3, 9 CBAM blocks), fp32 torch-CPU oracle of trainer.py:448-525"}


_REAL_STDOUT_FD = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath (NCCL prints its version banner to stdout at the first
    communicator) must not add to it: file descriptor 1 points at stderr until the result line is written."""
    global _REAL_STDOUT_FD
    sys.stdout.flush()
    _REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(_REAL_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-slices", type=int, default=int(os.environ.get("DUCOSY_BATCH_SLICES", "30")))
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--train-steps", type=int, default=3, help="timed CycleGAN steps for the 'train' object (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    _quiet_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from ducosy_gan_b200.synthesis import DualHUSynthesizer
    soft, lung = make_models(device)
    synth = DualHUSynthesizer(soft, lung, batch_slices=args.batch_slices, device=device)
    host_vol = torch.from_numpy(synthetic_volume(rank)).pin_memory()
    host_out = torch.empty_like(host_vol).pin_memory()
    dev_vol = host_vol.to(device)
    dev_out = torch.empty_like(dev_vol)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        sec = torch.tensor([e0.elapsed_time(e1) / 1e3], device=device)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return float(sec.item())

    dev_step = lambda: synth.synthesize_device(dev_vol, SLOPE, INTERCEPT, out=dev_out)
    e2e_step = lambda: synth.synthesize_volume(host_vol, SLOPE, INTERCEPT, out_host=host_out)

    for _ in range(args.warmup):
        dev_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sec = timed(dev_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    e2e_step()
    sec_e2e = timed(e2e_step, args.steps)

    chunks = (S + args.batch_slices - 1) // args.batch_slices
    launches = synth.launches_per_chunk() * chunks * args.steps
    train = None
    if args.train_steps > 0:
        del dev_out, dev_vol
        synth = None
        torch.cuda.empty_cache()
        try:
            train = train_metric(device, rank, world, args.train_steps)
        except Exception as exc:          # the synthesis line above is the contract; a failing extra must not take it down
            train = {"metric": "cyclegan_train_steps_per_s", "error": f"{type(exc).__name__}: {exc}"[:300]}

    value = world * S * args.steps / sec
    e2e_value = world * S * args.steps / sec_e2e
    if rank == 0:
        pk = peaks()
        ksec, kflops = time_dominant_kernel(device, args.batch_slices)
        achieved = kflops / ksec / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": os.environ.get("DUCOSY_PRECISION", "fp16") + " operands, fp32 accumulate",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "slices_per_gpu": S, "batch_slices": args.batch_slices,
                       "weights": "random init (weights_init_normal, seed 1234)",
                       "l2": "inputs larger than L2 (157 MB volume, >2 GB of activations per batch)",
                       "parallelism": f"slice-sharded x{world}, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": S * H * W * 2, "d2h_bytes_per_step": S * H * W * 2},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel<256> (3x3 256->256 residual-block conv)",
                         "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"],
                         "traffic": 4.85e8 if args.batch_slices == 30 else None,
                         "traffic_note": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture at "
                                         "batch 30 (profiles/r01_final_conv256_kernel.txt); algorithmic 5.13e8",
                         "peak_source": pk["source"] + ", burst figure (kernel timed alone)",
                         "path_frac_of_sustained": value / world * GFLOP_PER_SLICE / 1e3 / pk["tf_sustained"]},
        }
        if train is not None:
            line["train"] = train
        if world == 1 and not args.skip_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
            if train is not None and "error" not in train:
                try:
                    train["cpu_baseline"] = cpu_train_baseline()
                except Exception as exc:
                    train["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        _emit(line)
    if world > 1:
        os.dup2(2, 1)      # NCCL teardown chatter, if any, after the result line
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
