"""The host-side material table and Tiles::create recipes against the reference itself."""
import json
import os

import numpy as np
import pytest

from fallingsandsurvival_b200 import abi
from fallingsandsurvival_b200 import materials as M
from helpers import GOLDEN, full_table


def _golden_rows():
    with open(os.path.join(GOLDEN, "materials_srand1234.json")) as f:
        return json.load(f)


def test_named_materials_match_golden_export():
    """materials.py rows 0..30 == Materials::MATERIALS exported from the reference (Materials.cpp:4-199)."""
    rows = _golden_rows()
    defs = M.named_materials()
    assert len(defs) == 31 and len(rows) == 41
    for d, r in zip(defs, rows):
        s = d.to_struct()
        assert d.name == r["name"], d.id
        assert s.physics_type == r["physics_type"], d.id
        assert s.iterations == r["iterations"], d.id
        assert s.slipperyness == r["slipperyness"], d.id
        assert s.density == np.float32(r["density"]), d.id
        assert s.add_temp == r["add_temp"], d.id
        assert s.conduction_self == np.float32(r["conduction_self"]), d.id
        assert s.conduction_other == np.float32(r["conduction_other"]), d.id
        assert s.color == r["color"], d.id
        assert s.flags == r["flags"], d.id
        assert [[s.reactions[j].type, s.reactions[j].data1, s.reactions[j].data2] for j in range(s.n_reactions)] == r["reactions"]
        assert r["interact"] == 0  # Materials.cpp:128,:144: no material has interactions enabled


def test_golden_export_is_fresh(have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle import refsim
    rows = _golden_rows()
    for (m, name, inter), r in zip(refsim.materials(), rows):
        assert name == r["name"] and m.physics_type == r["physics_type"] and m.iterations == r["iterations"]
        assert float(m.density) == r["density"] and m.color == r["color"]


def test_create_recipes_match_tiles_create(have_ref):
    """create_cell(recipe) == Tiles::create(mat, x, y) (Tiles.cpp:190-268) for every material, incl. RNG colours."""
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle import refsim
    defs = full_table()
    textures = [M.synthetic_texture(i) for i in range(M.N_TEXTURES)]
    rng = np.random.default_rng(0)
    for d in defs:
        for _ in range(6):
            x, y = int(rng.integers(0, 5000)), int(rng.integers(0, 5000))
            seed, tick, it = int(rng.integers(1 << 31)), int(rng.integers(1000)), int(rng.integers(6))
            cx, cy = int(rng.integers(0, 5000)), int(rng.integers(0, 5000))
            ref = refsim.create(d.id, x, y, seed, tick, it, cx, cy)
            key = M.rng_cell_key(M.rng_tick_key(seed, tick, it), cx, cy)
            mine = M.create_cell(defs, textures, d.id, x, y, key, 0)
            assert ref == mine, (d.id, d.name, ref, mine)


def test_synthetic_texture_matches_shim(have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refsim
    L = refsim.load()
    for t in (0, 4, 15):
        tex = M.synthetic_texture(t)
        for (x, y) in [(0, 0), (127, 127), (5, 77), (100, 3)]:
            assert int(tex[y, x]) == L.fssref_texture_pixel(t, x, y)


def test_struct_sizes():
    import ctypes as C
    assert abi.CELL_DTYPE.itemsize == 20
    assert C.sizeof(abi.Material) == 40 + 12 * abi.MAX_REACTIONS + 28
    assert C.sizeof(abi.Config) == 40
