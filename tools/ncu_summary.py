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


def rawcsv(path):
    """Table of the first launch of every distinct kernel in an exported raw page (ncu -i X.ncu-rep --page raw --csv)."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = lambda k: hdr.index(k) if k in hdr else None
    ki = col("Kernel Name")
    want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB read"), ("dram__bytes_write.sum", "MB written"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % active"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid")]

    def val(r, k):
        i = col(k)
        if i is None or r[i] == "":
            return float("nan")
        v = float(r[i].replace(",", ""))
        u = units[i]
        if k.startswith("gpu__time"):
            return v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        if "bytes" in k:
            return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        return v

    print("# first launch of every distinct (kernel, grid) in " + path.split("/")[-1] + " -- ncu --set full --clock-control none, one 30-slice chunk")
    print("# " + " | ".join(f"{n:>14s}" for _, n in want) + " | TB/s (read+write) | kernel")
    seen = set()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("ducosy::<unnamed>::", "").replace("void ", "")
        key = (name, r[col("launch__grid_size")] if col("launch__grid_size") is not None else "")
        if key in seen:
            continue
        seen.add(key)
        vals = [val(r, k) for k, _ in want]
        tbs = (vals[1] + vals[2]) / vals[0] if vals[0] else float("nan")
        print("  " + " | ".join(f"{v:14.1f}" for v in vals) + f" | {tbs:17.2f} | {name}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "rawcsv": rawcsv}[sys.argv[1]](sys.argv[2])
