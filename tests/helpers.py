"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

from fallingsandsurvival_b200 import abi, worlds
from fallingsandsurvival_b200 import materials as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def full_table():
    """31 named materials + the 10 `Mat_N` the reference generated under srand(1234)
    (tests/golden/materials_srand1234.json, exported from oracle/_ref by make_golden.py)."""
    defs = M.named_materials()
    with open(os.path.join(GOLDEN, "materials_srand1234.json")) as f:
        rows = json.load(f)
    for r in rows[31:]:
        defs.append(M.MaterialDef(r["id"], r["name"], r["physics_type"], r["slipperyness"], 255, r["density"],
                                  r["iterations"], color=r["color"]))
    return defs


MIXES = {
    "c2": worlds.MIX_C2,
    "c3": worlds.MIX_C3,
    "sand": [(M.TEST_SAND, 30), (M.COBBLE_STONE, 8), (M.DIRT, 5), (M.SOFT_DIRT_SAND, 5), (M.GRASS, 4)],
    "liquids": [(M.WATER, 25), (M.LAVA, 8), (M.TEST_LIQUID, 6), (M.GOLD_MOLTEN, 4), (M.COBBLE_STONE, 12)],
    "gas": [(M.STEAM, 25), (M.GAS, 10), (M.COBBLE_STONE, 15), (M.TEST_SAND, 5)],
    "fire": [(M.FIRE, 10), (M.SOFT_DIRT, 30), (M.CLOUD, 10), (M.TEST_SAND, 10)],
    "randmats": [(31 + i, 6) for i in range(10)] + [(M.STONE, 10)],
    "settled": [(M.COBBLE_STONE, 60), (M.TEST_SAND, 1)],
    "pockets": [(M.COBBLE_STONE, 68), (M.WATER, 22), (M.LAVA, 4)],  # liquid trapped in stone: settles within a few ticks
}


def make_world(size, mix, seed, defs=None):
    w, h = size
    cells = worlds.synthetic_world(w, h, MIXES[mix] if isinstance(mix, str) else mix, seed=seed, defs=defs or full_table())
    if mix in ("c3", "liquids", "fire"):
        # give the temperature stencil and the reactions something to chew on
        hot = cells["material"] == M.GOLD_ORE
        cells["temperature"][hot] = 600
    return cells, worlds.default_zone(w, h)


def run_ticks(world, ticks, temperature=True):
    for t in range(ticks):
        world.tick(t)
        if temperature and t % 4 == 2:  # Game.cpp:2759 `tickTime % 4 == 2`
            world.tick_temperature()


def diff_report(a, b, limit=5):
    bad = np.argwhere(a != b)
    lines = [f"{len(bad)} cells differ"]
    for y, x in bad[:limit]:
        lines.append(f"  ({x},{y}): {a[y, x]} vs {b[y, x]}")
    return "\n".join(lines)


def particle_scene(seed=3, size=(256, 224)):
    """A world plus hand-made particles that drive every branch of World::tickParticles (world.cpp:2171-2311):
    target force, `phase`, temporary lifetimes, the in-object states, landing in the start cell, the 32 x 32 landing
    search (free cell / same liquid / nothing found -> bounce), leaving the world, frozen outside the tick zone, and
    columns of particles that land on top of each other in one tick."""
    rng = np.random.default_rng(seed)
    w, h = size
    cells, zone = make_world(size, [(M.TEST_SAND, 6), (M.WATER, 5), (M.COBBLE_STONE, 4), (M.FIRE, 1), (M.SOFT_DIRT, 2)], seed=seed)
    defs = full_table()

    def stamp(x0, y0, x1, y1, mat):
        cells[y0:y1, x0:x1] = _flat_cells(mat, (y1 - y0, x1 - x0), defs)

    stamp(40, 40, 100, 100, M.COBBLE_STONE)   # bigger than the landing search: particles inside it bounce
    stamp(120, 40, 150, 70, M.OBJECT)         # PhysicsType::OBJECT
    stamp(120, 70, 150, 74, M.COBBLE_STONE)
    stamp(170, 60, 230, 64, M.WATER)          # a pool to pour into
    stamp(170, 64, 230, 100, M.COBBLE_STONE)
    stamp(24, 180, 232, 184, M.COBBLE_STONE)  # a floor
    cells[150:180, 24:232] = _flat_cells(M.AIR, (30, 208), defs)  # open air above the floor

    parts = []

    def add(x, y, vx, vy, mat=M.TEST_SAND, ay=0.1, **kw):
        p = np.zeros((), dtype=abi.LIVE_PARTICLE_DTYPE)
        p["x"], p["y"], p["vx"], p["vy"], p["ay"] = x, y, vx, vy, ay
        p["fade_time"] = 60
        t = _flat_cells(mat, (1, 1), defs)[0, 0]
        t["color"] = 0x123456 + len(parts)
        t["fluid_amount"] = kw.pop("amount", 2.0)
        p["tile"] = t
        for k, v in kw.items():
            p[k] = v
        parts.append(p)

    for k in range(12):                                   # columns that land on each other, on the floor
        for c in range(6):
            add(60 + 3 * c + 0.5, 150 + 2.5 * k, 0.0, 2.0 + 0.25 * c)
    for k in range(8):                                    # the same start cell for several particles
        add(130.25, 160.5, 0.1 * k - 0.3, 3.0)
    for k in range(10):                                   # inside the big stone block: nothing free within reach
        add(70 + k * 0.7, 70 + k * 0.3, 0.3, 0.4)
    for k in range(10):                                   # inside stone, near its edge: the search finds air
        add(42.5 + k, 43.5, 0.2, 0.3, mat=M.WATER if k % 2 else M.TEST_SAND)
    for k in range(10):                                   # water started inside the pool's stone: pours into the pool
        add(175.5 + 4 * k, 66.5, 0.05, -0.2, mat=M.WATER, amount=0.25 + 0.125 * k)
    for k in range(6):                                    # born inside an object, leaves it, then hits stone
        add(125.5 + 3 * k, 60.5 + k, 0.0, 1.5)
    for k in range(4):                                    # born in air, flies into the object: state 2, lands at once
        add(125.5 + 4 * k, 30.5, 0.0, 4.0)
    for k in range(6):                                    # attracted to a point
        add(30.0 + 5 * k, 160.0, 0.0, 0.0, ay=0.0, target_x=128.0, target_y=120.0 + 10 * k, target_force=0.4)
    for k in range(4):                                    # phase: goes through everything
        add(50.0 + 10 * k, 30.0, 0.5, 3.0, phase=1)
    for k in range(6):                                    # temporary: dies on contact or when its lifetime is over
        add(200.0 + 3 * k, 150.0 + k, 0.2, -0.5, ay=0.01, temporary=1, lifetime=2 + 3 * k, fade_time=10)
    add(5.5, 100.0, -2.0, 0.0)                            # outside the tick zone: frozen
    add(w - 20.5, 100.0, 30.0, 0.0)                       # leaves the world in one tick
    add(100.5, 20.5, 0.0, -30.0)
    for _ in range(200):                                  # a random crowd over the open area
        add(rng.uniform(30, 226), rng.uniform(120, 176), rng.uniform(-1.5, 1.5), rng.uniform(-1, 4),
            mat=M.WATER if rng.random() < 0.4 else M.TEST_SAND, amount=float(rng.uniform(0.1, 2.0)))
    return cells, zone, np.array(parts, dtype=abi.LIVE_PARTICLE_DTYPE)


def _flat_cells(mat, shape, defs):
    """cells of one material as the MaterialInstance constructor leaves them (colour = Material::color)"""
    c = np.zeros(shape, dtype=abi.CELL_DTYPE)
    c["material"] = mat
    c["color"] = defs[mat].color
    c["fluid_amount"] = 2.0
    return c


def random_particles(seed, n, size):
    """particles of every kind: fast ones (paths longer than the 64 cells that the device lists one by one), target forces,
    `phase`, temporary ones, liquids of three materials, start cells anywhere (inside solids, outside the zone and the world)"""
    rng = np.random.default_rng(seed)
    parts = np.zeros(n, dtype=abi.LIVE_PARTICLE_DTYPE)
    parts["x"] = rng.uniform(-4, size[0] + 4, n).astype(np.float32)
    parts["y"] = rng.uniform(-4, size[1] + 4, n).astype(np.float32)
    speed = np.where(rng.random(n) < 0.1, 100.0, 4.0)
    parts["vx"] = (rng.uniform(-1, 1, n) * speed).astype(np.float32)
    parts["vy"] = (rng.uniform(-1, 1, n) * speed).astype(np.float32)
    parts["ay"] = np.float32(0.1)
    parts["ax"] = (rng.uniform(-0.05, 0.05, n) * (rng.random(n) < 0.2)).astype(np.float32)
    force = rng.random(n) < 0.1
    parts["target_x"] = rng.uniform(0, size[0], n).astype(np.float32)
    parts["target_y"] = rng.uniform(0, size[1], n).astype(np.float32)
    parts["target_force"] = np.where(force, rng.uniform(0.1, 0.8, n), 0).astype(np.float32)
    parts["phase"] = rng.random(n) < 0.05
    parts["temporary"] = rng.random(n) < 0.15
    parts["lifetime"] = rng.integers(0, 12, n)
    parts["fade_time"] = 60
    parts["in_object_state"] = rng.integers(0, 3, n)
    parts["tile"]["material"] = rng.choice([M.TEST_SAND, M.WATER, M.LAVA, M.TEST_LIQUID, M.GOLD_ORE, M.FIRE, M.COBBLE_STONE], n)
    parts["tile"]["color"] = rng.integers(0, 1 << 24, n)
    parts["tile"]["fluid_amount"] = rng.uniform(0.01, 3.0, n).astype(np.float32)
    parts["tile"]["temperature"] = rng.integers(-100, 100, n)
    return parts


def thin_liquid_world(seed=6, size=(200, 176)):
    """liquids whose amounts sit around FLUID_MinValue (0.0005): the `amount < MinValue` exits of the SOUP rule
    (world.cpp:1346-1350 and the four `remainingValue < FLUID_MinValue` exits, :1414 ... :1518) and the deletion in scan 2
    (:1812-1815) all come into play, which a world of full cells (amount 2.0) never reaches"""
    rng = np.random.default_rng(seed)
    cells, zone = make_world(size, "liquids", seed=seed)
    soup = np.isin(cells["material"], [M.WATER, M.LAVA, M.TEST_LIQUID, M.GOLD_MOLTEN])
    r = rng.random(cells.shape)
    amt = cells["fluid_amount"].copy()
    amt[soup & (r < 0.15)] = 0.0
    thin = soup & (r >= 0.15) & (r < 0.6)
    amt[thin] = rng.uniform(0.0, 0.004, cells.shape).astype(np.float32)[thin]
    mid = soup & (r >= 0.6) & (r < 0.8)
    amt[mid] = rng.uniform(0.0, 0.6, cells.shape).astype(np.float32)[mid]
    cells["fluid_amount"] = amt
    return cells, zone


def fat_drops_world(seed=8, size=(200, 176)):
    """liquid cells holding 8 to 20 units in open air: a falling cell then leaves the grid as n = amount / 4 particles of
    amount / n each (world.cpp:1355-1371), which is the only way to get more than one spawn record per cell visit"""
    rng = np.random.default_rng(seed)
    cells, zone = make_world(size, [(M.WATER, 6), (M.LAVA, 2), (M.COBBLE_STONE, 6), (M.TEST_SAND, 4)], seed=seed)
    soup = np.isin(cells["material"], [M.WATER, M.LAVA])
    amt = cells["fluid_amount"].copy()
    amt[soup] = rng.uniform(0.5, 20.0, cells.shape).astype(np.float32)[soup]
    cells["fluid_amount"] = amt
    return cells, zone
