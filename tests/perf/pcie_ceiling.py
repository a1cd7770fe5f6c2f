#!/usr/bin/env python3
"""What this box's PCIe link gives a pinned host buffer: host -> device alone, device -> host alone, both at once.
The e2e figure of bench.py (fss_tick_host: the whole grid up and down every step) is bounded by the last line.
usage: python tests/perf/pcie_ceiling.py [GiB per direction]   (prints one JSON line)"""
import json
import sys

import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30))
h_up = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_dn = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=5):
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1, f0, f1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        if up:
            with torch.cuda.stream(s_up):
                e0.record()
                d_up.copy_(h_up, non_blocking=True)
                e1.record()
        if down:
            with torch.cuda.stream(s_dn):
                f0.record()
                h_dn.copy_(d_dn, non_blocking=True)
                f1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1) if up else 0.0, f0.elapsed_time(f1) if down else 0.0)
        best = max(best, n / ms / 1e6)
    return best


run(True, True, 1)
out = {"bytes_per_direction": n, "h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True),
       "both_at_once_gbs_per_direction": run(True, True),
       "note": "pinned host memory, one cudaMemcpyAsync per direction on its own stream, best of 5; GB/s = 1e9 bytes/s"}
print(json.dumps(out))
