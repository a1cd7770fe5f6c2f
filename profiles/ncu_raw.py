#!/usr/bin/env python3
"""Print the metrics we track from `ncu --page raw --csv` output, one column per captured launch.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python profiles/ncu_raw.py raw.csv [stride]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"] + [
    f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in
    ("long_scoreboard", "wait", "branch_resolving", "no_instruction", "short_scoreboard", "lg_throttle", "not_selected",
     "math_pipe_throttle", "barrier", "dispatch_stall", "mio_throttle", "selected")]
for w in WANT:
    if w in idx:
        print(f"{w:75s} {units[idx[w]]:12s}", [r[idx[w]] for r in rows[2:]][::stride])
