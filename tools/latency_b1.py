"""Batch-1 drop-in latency: ``Generator(x[1,1,512,512])`` called slice by slice exactly as reference generate.py:89-102 does
(soft-tissue model, then lung model, then both outputs brought to the host for postprocess_tensor), with the modules'
CUDA-graph replay on (default) and off (DUCOSY_FORWARD_GRAPH=0), plus the back-to-back rate without the per-slice host sync."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
soft, lung = bench.make_models(dev)
xs = [(torch.rand(1, 1, 512, 512) * 2 - 1).pin_memory() for _ in range(8)]
res = {}


def generate_py_pattern(n):
    """per slice: .to(device) x2, two forwards, .cpu() x2 (generate.py:94-102, preprocess.py:96)"""
    t0 = time.perf_counter()
    for i in range(n):
        x = xs[i % len(xs)]
        a, b = x.to(dev), x.to(dev)
        ys, yl = soft(a), lung(b)
        ys.cpu(), yl.cpu()
    return (time.perf_counter() - t0) / n * 1e3


def back_to_back(n):
    x = xs[0].to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        ys, yl = soft(x), lung(x)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


with torch.no_grad():
    for mode, flag in (("eager", "0"), ("graph", "1")):
        os.environ["DUCOSY_FORWARD_GRAPH"] = flag
        generate_py_pattern(5)
        back_to_back(5)
        res[f"{mode}_ms_per_slice_generate_py_pattern"] = generate_py_pattern(100)
        res[f"{mode}_ms_per_slice_back_to_back"] = back_to_back(100)
    # equality of the two paths on a fresh input
    x = xs[3].to(dev)
    os.environ["DUCOSY_FORWARD_GRAPH"] = "0"
    e = soft(x)
    os.environ["DUCOSY_FORWARD_GRAPH"] = "1"
    res["graph_output_equals_eager"] = bool(torch.equal(soft(x), e))
res["slices_per_s_generate_py_pattern"] = 1e3 / res["graph_ms_per_slice_generate_py_pattern"]
res["note"] = "both generators per slice; the reference's own figure is 100-200 ms per slice on an RTX 4090 including DICOM I/O (README.md:506-509)"
print(json.dumps(res, indent=1))
