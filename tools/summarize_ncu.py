#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small text summaries kept
under profiles/ (the .ncu-rep files themselves are scratch).

  launches CSV  (ncu --metrics gpu__time_duration.sum --clock-control none --csv)
      python tools/summarize_ncu.py launches gpurun_out/launches_x.csv > profiles/x.md
  full report   (ncu --set full --clock-control none --import-source on -o rep)
      python tools/summarize_ncu.py full gpurun_out/rep.ncu-rep > profiles/y.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("amgb::", "")


def launches(path):
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit == "ns" else (v if unit == "us" else v * 1e3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
        total += us
    print(f"source: {path}; {sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms of kernel time "
          f"(serialised, cold cache, no clock control: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share | avg us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {100 * t / total:.1f}% | {t / c:.1f} | {mx:.1f} |")


WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "smsp cycles"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short sb"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch"),
]


def full(path):
    if path.endswith(".csv"):   # already exported on the GPU box (ncu -i rep --page raw --csv)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True,
                             check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), lbl) for m, lbl in WANT if m in hdr]
    ki = hdr.index("Kernel Name")
    print(f"source: {path} (ncu --set full --clock-control none; one replayed launch per row)\n")
    print("| # | kernel | " + " | ".join(f"{lbl} [{units[i]}]" if units[i] else lbl for i, lbl in cols) + " |")
    print("|---|---|" + "---:|" * len(cols))
    for n, r in enumerate(rows[2:]):
        name = re.sub(r"\(.*$", "", re.sub(r"^void ", "", r[ki]))
        vals = []
        for i, _ in cols:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {n} | `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
