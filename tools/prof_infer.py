"""Profiling driver for ncu (--profile-from-start off): one warm chunk, then ONE chunk of the dual-HU synthesis between
cudaProfilerStart / Stop.  AB_BATCH slices per chunk (default 30); DUCOSY_SINGLE_STREAM=1 keeps the two generators in
launch order (soft-tissue generator first)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ducosy_gan_b200.synthesis import DualHUSynthesizer  # noqa: E402

dev = torch.device("cuda", 0)
soft, lung = bench.make_models(dev)
B = int(os.environ.get("AB_BATCH", "30"))
synth = DualHUSynthesizer(soft, lung, batch_slices=B, device=dev)
vol = torch.from_numpy(bench.synthetic_volume(0)[:B]).to(dev)
out = torch.empty_like(vol)
synth.synthesize_device(vol, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
synth.synthesize_device(vol, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("checksum", int(out.to(torch.int64).sum()))
