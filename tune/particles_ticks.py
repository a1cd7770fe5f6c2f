#!/usr/bin/env python3
"""a few World::tick + fss_tick_particles on a c2 world (for ncu captures of the particle kernels at later ticks)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import time
import helpers
from fallingsandsurvival_b200 import materials as M, world, worlds
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = world.World(None, (16, 16, S, S), defs=helpers.full_table(), size=(S + 32, S + 32), particles=True, particle_capacity=1 << 24)
g.fill_synthetic(worlds.MIX_C2, border=16, border_material=M.COBBLE_STONE)
for t in range(T):
    g.tick(t)
    g.sync()
    t0 = time.perf_counter()
    g.tick_particles()
    print("tick", t, "tick_particles ms %.3f" % ((time.perf_counter() - t0) * 1e3), "rounds/passes", g.particle_rounds(), "left", g.particles_count())
g.close()
