"""Data-parallel CycleGAN step check + timing (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/train_dp_check.py
(1) generator / discriminator gradients after the all-reduce against the single-process step on the whole batch
    (rank 0 computes that reference with the same kernels); (2) steps/s of the sharded step at global batch 8."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ducosy_gan_b200.data_parallel import DataParallelCycleGANStep, GraphedCycleGANStep, shard_batch  # noqa: E402
from ducosy_gan_b200.trainer import CycleGANStep  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
Cin, blocks = 3, int(os.environ.get("DP_BLOCKS", 9))


def batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    smooth = lambda t: torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0
    a = smooth(torch.rand(B, 1, 512, 512, generator=g) * 2 - 1).clamp(-1, 1).cuda()
    b = smooth(torch.rand(B, 1, 512, 512, generator=g) * 2 - 1).clamp(-1, 1).cuda()
    m = (torch.rand(B, Cin - 1, 512, 512, generator=g) < 0.1).float().cuda()
    return a, b, m


res = {"world": world}
# ---- (1) gradient parity, global batch = 2 * world
Bg = 2 * world
A, Bt, M = batch(Bg, 5)
lo, hi = shard_batch(Bg, rank, world)
dp = DataParallelCycleGANStep(Cin, blocks, True, seed=77)
dp.bucket_G.zero()
loss_G, terms, _, _ = dp.generator_losses(A[lo:hi], Bt[lo:hi], M[lo:hi])
loss_G.backward()
dp.bucket_G.all_reduce_mean()
torch.cuda.synchronize()
if rank == 0:
    single = CycleGANStep(Cin, blocks, True, seed=77)
    l1, t1, _, _ = single.generator_losses(A, Bt, M)
    l1.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in list(single.G_A2B.parameters()) + list(single.G_B2A.parameters())])
    got = dp.bucket_G.flat
    res["grad_rel_l2_vs_single_process"] = ((got - ref).norm() / ref.norm()).item()
    res["grad_cosine"] = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    res["loss_terms_dp_rank0"] = {k: float(v) for k, v in terms.items()}
    res["loss_terms_single"] = {k: float(v) for k, v in t1.items()}
    del single
del dp
torch.cuda.empty_cache()
dist.barrier()
# ---- (2) timing at global batch 8
A, Bt, M = batch(8, 2)
lo, hi = shard_batch(8, rank, world)
use_graph = os.environ.get("DP_GRAPH", "1") == "1"
dp = DataParallelCycleGANStep(Cin, blocks, True, seed=1234, capturable=use_graph)
if use_graph:
    graphed = GraphedCycleGANStep(dp, A[lo:hi].contiguous(), Bt[lo:hi].contiguous(), M[lo:hi].contiguous(), warmup=2)
    run_step = lambda: graphed(A[lo:hi], Bt[lo:hi], M[lo:hi])
else:
    run_step = lambda: dp.step(A[lo:hi], Bt[lo:hi], M[lo:hi])
for _ in range(2):
    run_step()
steps = 5
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = run_step()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    res.update(global_batch=8, cuda_graph=use_graph, ms_per_step=t.item(), steps_per_s=1e3 / t.item(), samples_per_s=8e3 / t.item(),
               final_losses={k: float(v) for k, v in out.items()})
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/train_dp{world}.json", "w"), indent=1)
if use_graph:
    graphed.close()
dist.barrier()
dist.destroy_process_group()
