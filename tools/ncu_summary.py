"""Summarise ncu outputs into small text files for profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.txt   (needs ncu on PATH)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("ducosy::<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# per-kernel device time over {sum(cnt.values())} consecutive launches (ncu gpu__time_duration.sum, serialised, cold cache)")
    print(f"# total {T:.1f} us")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d} avg={v / cnt[k]:8.1f} us  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full, {len(rows) - 2} launches of {rows[2][ki][:100]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} [{units[i]}]  " + "  ".join(r[i] for r in rows[2:]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
