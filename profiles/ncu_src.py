#!/usr/bin/env python3
"""Aggregate `ncu --page source --print-source cuda,sass --csv` per CUDA source line for the first captured launch.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python profiles/ncu_src.py src.csv [launch] [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
agg = defaultdict(lambda: [0.0, 0.0, 0.0, ""])  # samples, inst, thread inst
fpath = None
hdr = None
seen = defaultdict(int)
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        seen[fpath] += 1
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or seen[fpath] - 1 != launch:
        continue
    try:
        line = int(r[0])
        smp = float(r[hdr["# Samples"]] or 0)
        ins = float(r[hdr["Instructions Executed"]] or 0)
        tin = float(r[hdr["Thread Instructions Executed"]] or 0)
    except (ValueError, IndexError):
        continue
    a = agg[(fpath, line)]
    a[0] += smp
    a[1] += ins
    a[2] += tin
    if r[1].strip():
        a[3] = r[1].strip()[:100]
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"launch {launch}: samples {ts:.0f}  warp-instructions {ti:.0f}  avg active threads {sum(a[2] for a in agg.values())/ti:.1f}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]/ts*100:5.1f}% smp {a[1]/ti*100:5.1f}% ins thr {a[2]/max(a[1],1):4.1f}  {f}:{l}  {a[3]}")
