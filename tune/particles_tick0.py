#!/usr/bin/env python3
"""one World::tick + one fss_tick_particles on a c2 world (for an ncu launch list of the particle kernels)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import time
import helpers
from fallingsandsurvival_b200 import materials as M, world, worlds
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = world.World(None, (16, 16, S, S), defs=helpers.full_table(), size=(S + 32, S + 32), particles=True, particle_capacity=1 << 24)
g.fill_synthetic(worlds.MIX_C2, border=16, border_material=M.COBBLE_STONE)
g.tick(0)
g.sync()
t0 = time.perf_counter()
g.tick_particles()
print("tick_particles ms", (time.perf_counter() - t0) * 1e3, "rounds/passes", g.particle_rounds(), "left", g.particles_count())
g.close()
