#!/usr/bin/env python3
"""The dirty -> pixels job of Game::tick (Game.cpp:2579-2650) on the device next to the reference's loop on the host.

  python tests/perf/bench_render.py [--size 8192] [--ticks 6] [--cpu-size 2048]

Device: c3 fill, World::tick with FSS_TRACK_DIRTY | FSS_TRACK_FLOW (and without, for the cost of tracking), then
fss_render_dirty timed with CUDA events (it ends with a host read of the counters).  CPU: the reference's own loop
(oracle/_ref) on one host thread, as the game runs it.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--ticks", type=int, default=6)
    ap.add_argument("--cpu-size", type=int, default=2048)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import helpers
    from fallingsandsurvival_b200 import materials as M, world, worlds
    defs = helpers.full_table()
    with open(os.path.join(helpers.GOLDEN, "materials_looks.json")) as f:
        lk = json.load(f)
    alpha, emit = np.array(lk["alpha"], dtype=np.uint8), np.array(lk["emit_color"], dtype=np.uint32)
    S = args.size
    zone = (16, 16, S, S)
    out = {"metric": "cells_per_second (render job)", "size": S, "fill": "c3"}
    tick_ms = {}
    for tracked in (False, True):
        g = world.World(None, zone, defs=defs, size=(S + 32, S + 32), track_dirty=tracked, track_flow=tracked, early_out=False)
        g.fill_synthetic(worlds.MIX_C3, border=16, border_material=M.COBBLE_STONE)
        stream = torch.cuda.Stream()
        g.set_stream(stream.cuda_stream)
        if tracked:
            g.set_material_looks(alpha, emit)
        rows = []
        with torch.cuda.stream(stream):
            ms_tick = []
            for t in range(args.ticks):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record(stream)
                g.tick(t)
                e1.record(stream)
                if tracked:
                    moving, info = g.render_dirty()
                e2.record(stream)
                e2.synchronize()
                ms_tick.append(e0.elapsed_time(e1))
                if tracked:
                    rows.append(dict(tick=t, dirty_cells=int(moving.sum()), info=list(info), render_ms=e1.elapsed_time(e2)))
        tick_ms["tracked" if tracked else "plain"] = ms_tick
        if tracked:
            out["device"] = rows
            cells = (S + 32) * (S + 32)
            last = rows[-1]
            # bytes the job has to move: the dirty plane (1 B per stored cell, read) + per dirty cell mf, colour (8 B read),
            # dirty reset (1 B), main + emission pixel (8 B written); liquids add flow state (16 B read, 20 B written)
            out["render_bytes_model"] = cells + last["dirty_cells"] * 17
            out["render_gbs_model"] = out["render_bytes_model"] / (last["render_ms"] * 1e-3) / 1e9
            out["render_cells_per_s"] = cells / (last["render_ms"] * 1e-3)
        g.close()
    out["tick_ms_no_early_out"] = tick_ms
    if not args.no_cpu:
        from oracle import refsim
        if refsim.available():
            C = args.cpu_size
            cells = worlds.synthetic_world(C + 32, C + 32, worlds.MIX_C3, seed=0xF55, defs=defs)
            r = refsim.RefWorld(cells, (16, 16, C, C), particles=False, threads=os.cpu_count())
            cpu = []
            for t in range(args.ticks):
                r.tick(t)
                moving, info, dt = r.render_dirty()
                cpu.append(dict(tick=t, dirty_cells_mod_65536=int(moving.astype(np.uint64).sum()), ms=dt * 1e3))
            out["cpu_reference"] = {"size": C, "threads": 1, "rows": cpu, "cells_per_s": (C + 32) * (C + 32) / (cpu[-1]["ms"] * 1e-3)}
            r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
