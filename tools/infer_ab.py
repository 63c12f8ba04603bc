"""A/B of experiment switches on the device-resident synthesis throughput (one process per setting; the switches are read once).
usage: python tools/infer_ab.py   -> JSON {setting: slices/s}"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
import bench
from ducosy_gan_b200.synthesis import DualHUSynthesizer
dev = torch.device("cuda", 0)
soft, lung = bench.make_models(dev)
B = int(os.environ.get("AB_BATCH", "30"))
S = int(os.environ.get("AB_SLICES", "300"))
synth = DualHUSynthesizer(soft, lung, batch_slices=B, device=dev)
vol = torch.from_numpy(bench.synthetic_volume(0)[:S]).to(dev)
out = torch.empty_like(vol)
for _ in range(2): synth.synthesize_device(vol, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = int(os.environ.get("AB_STEPS", "3"))
for _ in range(n): synth.synthesize_device(vol, out=out)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"slices_per_s": S * n / (e0.elapsed_time(e1) / 1e3), "checksum": int(out.to(torch.int64).sum())}))
''' % ROOT

settings = [dict(), dict(DUCOSY_FUSED_FINALIZE="1"), dict(DUCOSY_FUSED_SPATIAL="0"), dict(DUCOSY_FUSED_FINALIZE="1", DUCOSY_FUSED_SPATIAL="0")]
if os.environ.get("AB_ONLY_DEFAULT"):
    settings = settings[:1]
extra = [s.split("=") for s in sys.argv[1:]]
res = {}
for st in settings:
    env = dict(os.environ, **st, **{k: v for k, v in extra})
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    key = ",".join(f"{k}={v}" for k, v in st.items()) or "default"
    try:
        res[key] = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        res[key] = {"error": (out.stderr or out.stdout)[-400:]}
print(json.dumps(res, indent=1))
