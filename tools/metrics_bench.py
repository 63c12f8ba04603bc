"""Times ducosy_gan_b200.metrics.volume_metrics (calculate.py's MAE / PSNR / SSIM / CS / ED, raw + normalised) on a 300 x 512 x 512
int16 pair resident on the GPU, next to the oracle (numpy / scipy restatement of the same functions) on a 6-slice sample."""
import json, os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ducosy_gan_b200 import metrics
from oracle import ducosy_oracle as orc
S = int(sys.argv[1]) if len(sys.argv) > 1 else 300
tgt, pred = orc.metrics_test_volumes(6, 512, 512, 11)
reps = (S + 5) // 6
t = torch.from_numpy(np.concatenate([tgt] * reps)[:S]).cuda()
p = torch.from_numpy(np.concatenate([pred] * reps)[:S]).cuda()
for _ in range(2):
    metrics.volume_metrics(t, p)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    out = metrics.volume_metrics(t, p)
torch.cuda.synchronize()
gpu_s = (time.perf_counter() - t0) / n
warnings.simplefilter("ignore")
t0 = time.perf_counter()
tn, pn = orc.metric_normalize(tgt), orc.metric_normalize(pred)
for fn, (x, y) in [(orc.metric_mae, (tgt, pred)), (orc.metric_psnr, (tgt, pred)), (orc.metric_ssim, (tgt, pred)), (orc.metric_cs, (tgt, pred)),
                   (orc.metric_ed, (tgt, pred)), (orc.metric_mae, (tn, pn)), (orc.metric_psnr, (tn, pn)), (orc.metric_ssim, (tn, pn))]:
    fn(x, y)
cpu_s_6 = time.perf_counter() - t0
print(json.dumps({"slices": S, "gpu_ms_per_volume": gpu_s * 1e3, "gpu_slices_per_s": S / gpu_s, "cpu_oracle_slices_per_s": 6 / cpu_s_6,
                  "psnr": out["psnr"][0], "ssim": out["ssim"][0]}))
