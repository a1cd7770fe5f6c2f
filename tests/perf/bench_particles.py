#!/usr/bin/env python3
"""World::tickParticles on the device (fss_tick_particles) next to the serial loop on the host.

  python tests/perf/bench_particles.py [--size 4096] [--ticks 12] [--cpu-size 1024]

Each step = World::tick (which feeds the list) + World::tickParticles.  The timed figure is the particle step alone (CUDA
events around fss_tick_particles, which ends with a host read of the list length).  CPU: the reference's own lines
(oracle/_ref) when built, else the plain-C restatement, single thread like the reference's remove_if loop.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--ticks", type=int, default=12)
    ap.add_argument("--cpu-size", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    from fallingsandsurvival_b200 import materials as M, world, worlds
    import helpers
    defs = helpers.full_table()
    mix = worlds.MIX_C2
    S = args.size
    zone = (16, 16, S, S)
    g = world.World(None, zone, defs=defs, size=(S + 32, S + 32), particles=True, particle_capacity=1 << 26)
    g.fill_synthetic(mix, border=16, border_material=M.COBBLE_STONE)
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    rows = []
    with torch.cuda.stream(stream):
        for t in range(args.ticks):
            g.tick(t)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.tick_particles()
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            rows.append(dict(tick=t, live_after=g.particles_count(), rounds=g.particle_rounds()[0], passes=g.particle_rounds()[1], ms=ms))
    g.close()
    # particles processed in a step = those alive after the previous step + the ones this tick spawned; the list length
    # before the step is not returned by the ABI, so count what the serial loop visits on the host run below for the rate
    out = {"metric": "particle_steps_per_second", "device": rows, "size": S}
    if not args.no_cpu:
        from oracle import oraclesim, refsim
        C = args.cpu_size
        cells = worlds.synthetic_world(C + 32, C + 32, mix, seed=0xF55, defs=defs)
        czone = (16, 16, C, C)
        use_ref = refsim.available()
        w = refsim.RefWorld(cells, czone, particles=True, threads=os.cpu_count()) if use_ref else \
            oraclesim.OracleWorld(cells, czone, defs=defs, particles=True)
        if use_ref:
            w.keep_particles()
        gg = world.World(cells, czone, defs=defs, particles=True, particle_capacity=1 << 24)
        cpu = []
        for t in range(args.ticks):
            w.tick(t)
            gg.tick(t)
            before = len(w.particle_list()) if use_ref else None
            t0 = time.perf_counter()
            w.tick_particles()
            dt = time.perf_counter() - t0
            gg.sync()
            t0 = time.perf_counter()
            gg.tick_particles()  # ends with a host read of the list length: wall clock covers all of it
            gdt = time.perf_counter() - t0
            n_after = len(w.particle_list())
            assert gg.particles_count() == n_after
            cpu.append(dict(tick=t, visited=before, live_after=n_after, cpu_ms=dt * 1e3, gpu_ms=gdt * 1e3, rounds=gg.particle_rounds()[0], passes=gg.particle_rounds()[1]))
        same = gg.particle_list().tobytes() == w.particle_list().tobytes()
        out["same_world"] = {"size": C, "kind": "reference" if use_ref else "port", "rows": cpu, "lists_identical": bool(same)}
        gg.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
