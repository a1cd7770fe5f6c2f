#!/usr/bin/env python3
"""Mint the golden fixtures from the reference's own tick (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  The reference ships no tests for this path
(SURVEY.md section 4), so these are the known answers everything else is pinned to.

  materials_srand1234.json  Materials::MATERIALS after srand(1234); Materials::init()
  tick_hashes.json          {case -> digest}: fss_hash-style digest of the whole grid after N ticks
  micro_*.npz               tiny before/after scenes per rule family (full cell records)
  particle_scene.json       World::tick + World::tickParticles on helpers.particle_scene(): per tick the list length and
                            sha256 of the particle list and of the grid

Run:  python tests/golden/make_golden.py      (needs oracle/_ref)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oraclesim, refsim  # noqa: E402

CASES = [
    # name, size, mix, seed, ticks, particles
    ("c2_160x160_s1_t40", (160, 160), "c2", 1, 40, False),
    ("c3_160x160_s2_t40", (160, 160), "c3", 2, 40, False),
    ("c3_particles_160x128_s3_t30", (160, 128), "c3", 3, 30, True),
    ("sand_192x160_s4_t60", (192, 160), "sand", 4, 60, False),
    ("sand_particles_128x160_s5_t30", (128, 160), "sand", 5, 30, True),
    ("liquids_160x160_s6_t40", (160, 160), "liquids", 6, 40, False),
    ("gas_160x128_s7_t40", (160, 128), "gas", 7, 40, False),
    ("fire_128x128_s8_t60", (128, 128), "fire", 8, 60, False),
    ("randmats_160x160_s9_t40", (160, 160), "randmats", 9, 40, False),
    ("settled_256x96_s10_t20", (256, 96), "settled", 10, 20, False),
    ("c3_320x224_s11_t100", (320, 224), "c3", 11, 100, False),
]


def export_materials():
    rows = []
    for i, (m, name, inter) in enumerate(refsim.materials()):
        rows.append(dict(id=i, name=name, physics_type=m.physics_type, iterations=m.iterations,
                         slipperyness=m.slipperyness, density=float(m.density), add_temp=m.add_temp,
                         conduction_self=float(m.conduction_self), conduction_other=float(m.conduction_other),
                         color=m.color, flags=m.flags, interact=inter,
                         reactions=[[m.reactions[j].type, m.reactions[j].data1, m.reactions[j].data2]
                                    for j in range(m.n_reactions)]))
    with open(os.path.join(HERE, "materials_srand1234.json"), "w") as f:
        json.dump(rows, f, indent=1)


def main():
    assert refsim.available(), "build oracle/_ref first: python oracle/build_ref.py"
    export_materials()
    import helpers
    hashes = {}
    for name, size, mix, seed, ticks, particles in CASES:
        cells, zone = helpers.make_world(size, mix, seed)
        r = refsim.RefWorld(cells, zone, particles=particles, seed=0xF55)
        helpers.run_ticks(r, ticks)
        out = r.cells()
        hashes[name] = dict(size=list(size), mix=mix, seed=seed, ticks=ticks, particles=particles,
                            digest=f"{oraclesim.hash_cells(out):016x}",
                            histogram={str(k): int(v) for k, v in zip(*np.unique(out['material'], return_counts=True))})
        if name.startswith(("fire", "gas", "sand_particles")):
            np.savez_compressed(os.path.join(HERE, f"micro_{name}.npz"), before=cells, after=out, zone=np.array(zone))
        r.close()
        print(name, hashes[name]["digest"])
    with open(os.path.join(HERE, "tick_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    particle_scene(helpers)
    render_scene(helpers)
    c1_full_size(helpers)


PARTICLE_SCENE_TICKS = 16
C1_CHECKPOINTS = (100, 500, 1000)


def c1_full_size(helpers):
    """c1_material_test_1984x1152.json: BASELINE.json configs[0] at its stated size (worlds.material_test_world: the
    MaterialTestGenerator layout, 1920 x 1088 tick zone), World::tick with tickTemperature every 4th tick as Game::tick runs
    them (Game.cpp:2759), particles emitted, 1000 ticks on the reference's own lines (about 20 minutes on 8 cores)."""
    from fallingsandsurvival_b200 import worlds
    cells, zone = worlds.material_test_world(helpers.full_table())
    r = refsim.RefWorld(cells, zone, particles=True, seed=0xF55)
    out = dict(size=list(worlds.C1_SIZE), zone=list(zone), checkpoints={})
    spawned = 0
    for t in range(max(C1_CHECKPOINTS)):
        r.tick(t)
        if t % 4 == 2:
            r.tick_temperature()
        if t + 1 in C1_CHECKPOINTS:
            spawned += len(r.drain_particles()[1])
            c = r.cells()
            out["checkpoints"][str(t + 1)] = dict(
                digest=f"{oraclesim.hash_cells(c):016x}", spawn_records=spawned,
                histogram={str(k): int(v) for k, v in zip(*np.unique(c['material'], return_counts=True))})
            print("c1", t + 1, out["checkpoints"][str(t + 1)]["digest"], flush=True)
    r.close()
    with open(os.path.join(HERE, "c1_material_test_1984x1152.json"), "w") as f:
        json.dump(out, f, indent=1)


def render_scene(helpers):
    """materials_looks.json (Material::alpha / emitColor) and render_scene.json: World::tick + World::tickParticles + the
    dirty -> pixels job of Game::tick (Game.cpp:2584-2650) every third tick, all three being the reference's own lines"""
    import hashlib
    alpha, emit = refsim.material_looks()
    with open(os.path.join(HERE, "materials_looks.json"), "w") as f:
        json.dump(dict(alpha=[int(v) for v in alpha], emit_color=[int(v) for v in emit]), f)
    cells, zone = helpers.make_world((264, 240), "c3", 4002)
    r = refsim.RefWorld(cells, zone, particles=True, seed=0xF55)
    r.keep_particles()
    rows = []
    for t in range(18):
        r.tick(t)
        r.tick_particles()
        row = dict(tick=t, dirty=hashlib.sha256(r.dirty().tobytes()).hexdigest(), n_dirty=int(r.dirty().sum()))
        if t % 3 == 1:
            moving, info, _ = r.render_dirty()
            row.update(info=list(info), moving=[int(v) for v in moving],
                       layers=[hashlib.sha256(r.pixels[k].tobytes()).hexdigest() for k in range(4)])
        rows.append(row)
    r.close()
    with open(os.path.join(HERE, "render_scene.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print("render_scene", rows[-1]["n_dirty"], rows[-2]["layers"][0][:16])


def particle_scene(helpers):
    import hashlib
    cells, zone, parts = helpers.particle_scene()
    r = refsim.RefWorld(cells, zone, particles=True, seed=0xF55)
    r.keep_particles()
    r.add_particles(parts)
    rows = []
    for t in range(PARTICLE_SCENE_TICKS):
        r.tick(t)
        r.tick_particles()
        pl = r.particle_list()
        rows.append(dict(tick=t, n=int(len(pl)), particles=hashlib.sha256(pl.tobytes()).hexdigest(),
                         cells=hashlib.sha256(r.cells().tobytes()).hexdigest(),
                         bounced=int((pl["vy"] == -4).sum()), in_object=int((pl["in_object_state"] == 1).sum())))
    r.close()
    with open(os.path.join(HERE, "particle_scene.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print("particle_scene", rows[-1]["n"], rows[-1]["particles"][:16])


if __name__ == "__main__":
    if "--c1" in sys.argv:  # only the full-size configs[0] record (the slow one)
        import helpers
        c1_full_size(helpers)
    else:
        main()
