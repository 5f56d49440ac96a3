#!/usr/bin/env python
"""Per-instruction view of an exported ncu source page (ncu -i rep --page source --csv
--print-kernel-base function [| gzip]): executed warp instructions and stall samples per SASS
line, grouped per kernel.  usage: ncu_source_hot.py file.csv[.gz] [kernel-substring] [top N | all]"""
import csv, gzip, io, sys
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
mode = sys.argv[3] if len(sys.argv) > 3 else "40"
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for n, b in enumerate(blocks):
    if want not in b["name"] and want != str(n):
        continue
    h = b["hdr"]
    ia, isrc, iex, ismp, ithr = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
    rows = b["rows"]
    tot = sum(int(r[iex]) for r in rows)
    tots = sum(int(r[ismp]) for r in rows)
    print(f"== kernel {n}: {b['name']}: {len(rows)} SASS lines, {tot:,} warp instructions, {tots:,} samples")
    if mode == "all":
        for k, r in enumerate(rows):
            print(f"{k:5d} {int(r[iex]):>12,} {100*int(r[iex])/tot:5.2f}% smp {100*int(r[ismp])/max(tots,1):5.2f}% thr {r[ithr]:>5} {r[isrc].strip()}")
    else:
        top = sorted(range(len(rows)), key=lambda k: -int(rows[k][ismp]))[:int(mode)]
        for k in sorted(top):
            r = rows[k]
            print(f"{k:5d} {int(r[iex]):>12,} {100*int(r[iex])/tot:5.2f}% smp {100*int(r[ismp])/max(tots,1):5.2f}% thr {r[ithr]:>5} {r[isrc].strip()}")
