"""Batch-1 drop-in latency: Generator(x[1,1,512,512]) called slice by slice as generate.py:89-102 does,
eager and replayed from a CUDA graph."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
soft, lung = bench.make_models(dev)
x = torch.rand(1, 1, 512, 512, device=dev) * 2 - 1
with torch.no_grad():
    for _ in range(5):
        soft(x); lung(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        ys = soft(x); yl = lung(x)
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t0) / n * 1e3
    # CUDA graph of both forwards
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        soft(x); lung(x)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            ys = soft(x); yl = lung(x)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        g.replay()
    torch.cuda.synchronize()
    graph_ms = (time.perf_counter() - t0) / n * 1e3
print(json.dumps({"eager_ms_per_slice_both_generators": eager_ms, "graph_ms_per_slice_both_generators": graph_ms,
                  "eager_slices_per_s": 1e3 / eager_ms, "graph_slices_per_s": 1e3 / graph_ms}))
