#!/usr/bin/env python3
"""Per-kernel summary of one tick's movement launches from `ncu --metrics ... --csv --log-file X.csv` (long format: one row per
launch and metric).  Writes the JSON that bench.py reports as roofline.traffic.
usage: python profiles/ncu_tick.py X.csv OUT.json "<where the capture came from>" """
import csv
import json
import sys
from collections import OrderedDict, defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = {h: i for i, h in enumerate(rows[0])}
launches = OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(r[hdr["ID"]], {"kernel": r[hdr["Kernel Name"]]})
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    unit = r[hdr["Metric Unit"]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(unit, 1.0)
    d[r[hdr["Metric Name"]]] = v * scale
by = defaultdict(list)
for d in launches.values():
    by[d["kernel"]].append(d)
out = {"source": sys.argv[3], "launches": len(launches)}
tot = [d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in launches.values()]
out["dram_bytes_per_launch_mean"] = sum(tot) / len(tot)
out["dram_bytes_per_tick"] = sum(tot)
out["warp_instructions_per_tick"] = sum(d["smsp__inst_executed.sum"] for d in launches.values())
out["by_kernel"] = {}
for k, v in by.items():
    n = len(v)
    out["by_kernel"][k] = {
        "launches": n,
        "ms_mean_under_ncu": sum(d["gpu__time_duration.sum"] for d in v) / n * 1e3,
        "dram_bytes_mean": sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in v) / n,
        "warp_instructions_mean": sum(d["smsp__inst_executed.sum"] for d in v) / n,
        "active_lanes_mean": sum(d["smsp__thread_inst_executed_per_inst_executed.ratio"] for d in v) / n,
        "issue_active_pct_mean": sum(d["smsp__issue_active.avg.pct_of_peak_sustained_active"] for d in v) / n,
        "stalled_no_instruction_per_issue_mean": sum(d.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 0.0) for d in v) / n,
    }
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["by_kernel"], indent=1))
