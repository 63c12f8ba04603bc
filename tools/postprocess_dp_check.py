"""z-sharded post-composite smoothing under torchrun (one rank per GPU): the halo-exchange + min/max all-reduce version must
be bit-identical to the single-process result on the whole volume (rank 0 computes that with the same kernels and also
checks it against the scipy oracle).  Also runs the whole sharded synthesis call with postprocess=True once."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ducosy_gan_b200.postprocess import postprocess_volume, postprocess_volume_sharded  # noqa: E402
from ducosy_gan_b200.synthesis import DualHUSynthesizer, shard_range  # noqa: E402
from oracle import ducosy_oracle as orc  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
S, H, W = 10 * world, 128, 192          # equal shards (all_gather needs equal shapes)
vol = orc.postprocess_test_volume(S, H, W, seed=9)
lo, hi = shard_range(S, rank, world)
mine = postprocess_volume_sharded(torch.from_numpy(vol[lo:hi]).cuda().contiguous())
parts = [torch.empty((shard_range(S, r, world)[1] - shard_range(S, r, world)[0], H, W), dtype=torch.int16, device="cuda") for r in range(world)]
dist.all_gather([p.view(torch.uint8) for p in parts], mine.view(torch.uint8))
res = {"world": world, "volume": [S, H, W]}
if rank == 0:
    whole = torch.cat(parts).cpu().numpy()
    single = postprocess_volume(torch.from_numpy(vol).cuda()).cpu().numpy()
    res["sharded_equals_single_process"] = bool(np.array_equal(whole, single))
    res["single_process_equals_scipy_oracle"] = bool(np.array_equal(single, orc.postprocess_volume(vol)))
# whole path: sharded synthesis + smoothing (small generators to keep it quick)
from ducosy_gan_b200.modules.model import Generator, weights_init_normal  # noqa: E402
torch.manual_seed(3)
gs, gl = Generator(1, 2).cuda().apply(weights_init_normal), Generator(1, 2).cuda().apply(weights_init_normal)
synth = DualHUSynthesizer(gs, gl, batch_slices=4)
raw = orc.synthetic_volume(8 * world, 128, 128, seed=2)
lo, hi = shard_range(raw.shape[0], rank, world)
with torch.no_grad():
    out_local = synth.synthesize_device(torch.from_numpy(raw[lo:hi]).cuda(), postprocess=True)
    parts = [torch.empty_like(out_local) for _ in range(world)]
    dist.all_gather([p.view(torch.uint8) for p in parts], out_local.view(torch.uint8))
    if rank == 0:
        full = synth.synthesize_device(torch.from_numpy(raw).cuda())
        ref = postprocess_volume(full)
        res["sharded_synthesis_with_smoothing_equals_single_process"] = bool(torch.equal(torch.cat(parts), ref))
if rank == 0:
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/postprocess_dp{world}.json", "w"), indent=1)
dist.destroy_process_group()
