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
