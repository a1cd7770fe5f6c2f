#!/usr/bin/env python3
"""Latency of ONE World::tick at the size the game itself runs (tick zone 768 x 512 inside a 1024 x 768 world, Game.cpp:450-452,
:2216) and at BASELINE configs[0]'s size (1920 x 1088 zone): 24 plain kernel launches per tick against the same 24 launches
replayed from a CUDA graph (VERDICT round 1, weak point 10).  The graph bakes the tick counter (the RNG key) into its kernel
parameters, so a real caller would update the nodes' parameters per tick; here it only measures what launch overhead costs.

  python tests/perf/bench_small.py [--ticks 200]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import helpers  # noqa: E402
from fallingsandsurvival_b200 import world, worlds  # noqa: E402


def measure(name, cells, zone, ticks, particles):
    g = world.World(cells, zone, defs=helpers.full_table(), particles=particles)
    s = torch.cuda.Stream()
    g.set_stream(s.cuda_stream)
    for t in range(20):  # warm-up (first-use allocations happen here, not inside the capture)
        g.tick(t)
    g.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    t0 = time.perf_counter()
    for t in range(20, 20 + ticks):
        g.tick(t)
    e1.record(s)
    g.sync()
    plain_wall = (time.perf_counter() - t0) / ticks * 1e3
    plain_dev = e0.elapsed_time(e1) / ticks
    # one tick at a time, as a game loop calls it: launch, then wait for the result
    t0 = time.perf_counter()
    for t in range(50):
        g.tick(1000 + t)
        g.sync()
    sync_wall = (time.perf_counter() - t0) / 50 * 1e3
    out = {"scene": name, "zone": list(zone[2:]), "plain_ms_per_tick_device": plain_dev, "plain_ms_per_tick_wall": plain_wall,
           "tick_then_sync_ms_wall": sync_wall}
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            g.tick(5000)
        g.sync()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(ticks):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        out["graph_ms_per_tick_device"] = e0.elapsed_time(e1) / ticks
    except Exception as e:  # capture refused: say so instead of a number
        out["graph_ms_per_tick_device"] = None
        out["graph_error"] = str(e)[:200]
    g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ticks", type=int, default=200)
    args = ap.parse_args()
    res = []
    cells = worlds.synthetic_world(1024, 768, worlds.MIX_C2, seed=0xF55, defs=helpers.full_table(), border=16)
    res.append(measure("game window, c2 fill", cells, (128, 128, 768, 512), args.ticks, False))
    cells, zone = worlds.material_test_world(helpers.full_table())
    res.append(measure("configs[0]: MaterialTestGenerator world", cells, zone, args.ticks, True))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
