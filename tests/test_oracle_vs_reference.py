"""Pins the plain-C oracle (oracle/fss_oracle.c) to the reference.

1. against the golden digests minted from the reference's own lines (always runs);
2. cell-by-cell against oracle/_ref itself on fresh seeds (runs where /root/reference was available
   to build oracle/_ref -- in the build container and, because the .so travels, on the GPU box).
"""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, MIXES, diff_report, full_table, make_world, run_ticks
from oracle import oraclesim

with open(os.path.join(GOLDEN, "tick_hashes.json")) as f:
    HASHES = json.load(f)


@pytest.mark.parametrize("name", sorted(HASHES))
def test_oracle_matches_golden_digest(name):
    g = HASHES[name]
    cells, zone = make_world(tuple(g["size"]), g["mix"], g["seed"])
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=g["particles"])
    run_ticks(o, g["ticks"])
    out = o.cells()
    assert f"{oraclesim.hash_cells(out):016x}" == g["digest"]
    hist = {str(k): int(v) for k, v in zip(*np.unique(out["material"], return_counts=True))}
    assert hist == g["histogram"]


@pytest.mark.parametrize("name", ["fire_128x128_s8_t60", "gas_160x128_s7_t40", "sand_particles_128x160_s5_t30"])
def test_oracle_matches_golden_scene(name):
    """full cell records (colour, temperature, amounts) of the committed before/after scenes"""
    g = HASHES[name]
    z = np.load(os.path.join(GOLDEN, f"micro_{name}.npz"))
    o = oraclesim.OracleWorld(z["before"], tuple(int(v) for v in z["zone"]), defs=full_table(), particles=g["particles"])
    run_ticks(o, g["ticks"])
    out = o.cells()
    assert out.tobytes() == z["after"].tobytes(), diff_report(z["after"], out)


@pytest.mark.parametrize("mix", sorted(MIXES))
@pytest.mark.parametrize("particles", [False, True])
def test_oracle_matches_reference_build(mix, particles, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle import refsim
    cells, zone = make_world((200, 176), mix, seed=1000 + len(mix))
    r = refsim.RefWorld(cells, zone, particles=particles)
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=particles)
    for t in range(24):
        r.tick(t)
        o.tick(t)
        if t % 4 == 2:
            r.tick_temperature()
            o.tick_temperature()
        a, b = r.cells(), o.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    if particles:
        # the reference builds Particle objects; compare what left the grid (cell + position), order-free
        xyv, pc, temporary = r.drain_particles()
        mine = o.drain_particles()
        ref_keys = sorted((int(c["material"]), int(x), int(y)) for (x, y, _, _), c in zip(xyv, pc))
        # Particle(tile, x, y-1) for fire sparks (world.cpp:1180), (x, y+1) for sand/liquid (:1291, :1366)
        my_keys = sorted((int(p["cell"]["material"]), int(p["x"]), int(p["y"]) + (-1 if p["site"] == 1180 else 1)) for p in mine)
        assert ref_keys == my_keys
    r.close()
    o.close()


def test_oracle_particle_tick_matches_golden():
    """World::tick + World::tickParticles (world.cpp:2171-2311) on the hand-made scene that drives every branch; the
    golden hashes were minted from the reference's own lines (tests/golden/make_golden.py)"""
    import hashlib
    import json
    from helpers import GOLDEN, particle_scene
    with open(os.path.join(GOLDEN, "particle_scene.json")) as f:
        rows = json.load(f)
    cells, zone, parts = particle_scene()
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    o.add_particles(parts)
    assert rows[0]["bounced"] > 0 and rows[0]["in_object"] > 0
    for row in rows:
        o.tick(row["tick"])
        o.tick_particles()
        pl = o.particle_list()
        assert len(pl) == row["n"], f"tick {row['tick']}"
        assert hashlib.sha256(pl.tobytes()).hexdigest() == row["particles"], f"tick {row['tick']}"
        assert hashlib.sha256(o.cells().tobytes()).hexdigest() == row["cells"], f"tick {row['tick']}"
    o.close()


@pytest.mark.parametrize("mix", ["c3", "sand", "liquids", "fire"])
def test_oracle_particle_tick_matches_reference_build(mix, have_ref):
    """the list (order, positions, velocities, tiles) and the grid, every tick, with World::tick feeding the list"""
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle import refsim
    cells, zone = make_world((200, 176), mix, seed=2000 + len(mix))
    r = refsim.RefWorld(cells, zone, particles=True)
    r.keep_particles()
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    seen = 0
    for t in range(30):
        r.tick(t)
        o.tick(t)
        r.tick_particles()
        o.tick_particles()
        a, b = r.particle_list(), o.particle_list()
        assert len(a) == len(b) and a.tobytes() == b.tobytes(), f"tick {t}: particle lists differ"
        ca, cb = r.cells(), o.cells()
        assert ca.tobytes() == cb.tobytes(), f"tick {t}: " + diff_report(ca, cb)
        seen += len(a)
    assert seen > 0
    r.close()
    o.close()


@pytest.mark.parametrize("seed,n", [(1, 200), (2, 1500)])
def test_oracle_particle_tick_fuzz_matches_reference_build(seed, n, have_ref):
    """random particles of every kind (helpers.random_particles) in a busy world: list, grid and dirty[] every tick"""
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from helpers import random_particles
    from oracle import refsim
    size = (392, 304)
    cells, zone = make_world(size, "c3", seed=50 + seed)
    parts = random_particles(seed, n, size)
    r = refsim.RefWorld(cells, zone, particles=True)
    r.keep_particles()
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    r.add_particles(parts)
    o.add_particles(parts)
    for t in range(14):
        r.tick(t)
        o.tick(t)
        r.tick_particles()
        o.tick_particles()
        assert r.particle_list().tobytes() == o.particle_list().tobytes(), f"tick {t}"
        assert r.cells().tobytes() == o.cells().tobytes(), f"tick {t}"
        assert r.dirty().tobytes() == o.dirty().tobytes(), f"tick {t}"
        if t == 6:
            r.add_particles(parts[: n // 2])
            o.add_particles(parts[: n // 2])
    r.close()
    o.close()


def test_oracle_thin_liquids_match_reference_build(have_ref):
    """amounts around FLUID_MinValue: every early exit of the SOUP rule, every tick"""
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from helpers import thin_liquid_world
    from oracle import refsim
    cells, zone = thin_liquid_world()
    r = refsim.RefWorld(cells, zone, particles=True)
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    for t in range(30):
        r.tick(t)
        o.tick(t)
        a, b = r.cells(), o.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
        assert r.dirty().tobytes() == o.dirty().tobytes(), f"tick {t}: dirty differs"
    r.close()
    o.close()


def test_oracle_fat_drops_match_reference_build(have_ref):
    """several particles per falling liquid cell (world.cpp:1355-1371): spawn order, the 2j / 2j+1 velocity draws, landing"""
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from helpers import fat_drops_world
    from oracle import refsim
    cells, zone = fat_drops_world()
    r = refsim.RefWorld(cells, zone, particles=True)
    r.keep_particles()
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    most = 0
    for t in range(16):
        r.tick(t)
        o.tick(t)
        r.tick_particles()
        o.tick_particles()
        a, b = r.particle_list(), o.particle_list()
        assert a.tobytes() == b.tobytes(), f"tick {t}: particle lists differ ({len(a)} vs {len(b)})"
        assert r.cells().tobytes() == o.cells().tobytes(), f"tick {t}"
        most = max(most, len(b))
    assert most > 100
    r.close()
    o.close()


def _looks():
    with open(os.path.join(GOLDEN, "materials_looks.json")) as f:
        d = json.load(f)
    return np.array(d["alpha"], dtype=np.uint8), np.array(d["emit_color"], dtype=np.uint32)


def test_oracle_dirty_and_pixels_match_golden():
    """World::dirty as World::tick / tickParticles set it, and the dirty -> pixels job of Game::tick (Game.cpp:2584-2650);
    golden hashes minted from the reference's own lines (tests/golden/make_golden.py::render_scene)"""
    import hashlib
    with open(os.path.join(GOLDEN, "render_scene.json")) as f:
        rows = json.load(f)
    cells, zone = make_world((264, 240), "c3", 4002)
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    o.set_material_looks(*_looks())
    for row in rows:
        o.tick(row["tick"])
        o.tick_particles()
        d = o.dirty()
        assert int(d.sum()) == row["n_dirty"] and hashlib.sha256(d.tobytes()).hexdigest() == row["dirty"], f"tick {row['tick']}"
        if "layers" in row:
            moving, info = o.render_dirty()
            assert list(info) == row["info"]
            assert [int(v) & 0xFFFF for v in moving] == row["moving"]  # Game::movingTiles is uint16_t
            assert [hashlib.sha256(o.pixels[k].tobytes()).hexdigest() for k in range(4)] == row["layers"], f"tick {row['tick']}"
    o.close()


@pytest.mark.parametrize("mix", ["c3", "sand", "liquids", "gas", "fire"])
def test_oracle_dirty_and_pixels_match_reference_build(mix, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle import refsim
    cells, zone = make_world((200, 176), mix, seed=77)
    r = refsim.RefWorld(cells, zone, particles=True)
    r.keep_particles()
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    o.set_material_looks(*refsim.material_looks())
    for t in range(20):
        r.tick(t)
        o.tick(t)
        r.tick_particles()
        o.tick_particles()
        if t == 5:
            r.mark_dirty(30, 30, 50, 20)
            o.mark_dirty(30, 30, 50, 20)
        assert r.dirty().tobytes() == o.dirty().tobytes(), f"tick {t}: dirty differs"
        if t % 2 == 1:
            mr, ir, _ = r.render_dirty()
            mo, io = o.render_dirty()
            assert ir == io and np.array_equal(mr.astype(np.uint64), mo[:len(mr)] & np.uint64(0xFFFF))
            assert r.pixels.tobytes() == o.pixels.tobytes(), f"tick {t}: pixel layers differ"
            fa, fb = r.flow(), o.flow()
            assert fa[0].tobytes() == fb[0].tobytes() and fa[1].tobytes() == fb[1].tobytes()
    r.close()
    o.close()


def test_reference_build_is_thread_count_independent(have_ref):
    """the 4x16 checkerboard is race-free: 1, 3 and 8 pool threads give the same grid"""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refsim
    cells, zone = make_world((224, 208), "c3", seed=77)
    outs = []
    for th in (1, 3, 8):
        r = refsim.RefWorld(cells, zone, threads=th, particles=True)
        run_ticks(r, 12)
        outs.append(r.cells().tobytes())
        r.close()
    assert outs[0] == outs[1] == outs[2]


def test_temperature_rows_split_equals_whole():
    cells, zone = make_world((160, 160), "c3", seed=5)
    a = oraclesim.OracleWorld(cells, zone, defs=full_table())
    b = oraclesim.OracleWorld(cells, zone, defs=full_table())
    a.tick_temperature()
    # Jacobi: any row split gives the same answer only if done from the same input -> emulate halo copy
    snap = b.cells()
    b.tick_temperature_rows(0, 80)
    top = b.cells()
    b.view()[...] = snap
    b.tick_temperature_rows(80, 1 << 20)
    bot = b.cells()
    merged = np.concatenate([top[:80], bot[80:]])
    assert merged.tobytes() == a.cells().tobytes()


@pytest.mark.parametrize("mix", ["c2", "liquids", "c3"])
def test_oracle_flow_matches_reference_build(mix, have_ref):
    """World::flowX / flowY, the SOUP rule's side output for the water shader (world.cpp:1405, :1447, :1477, :1509)"""
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle import refsim
    cells, zone = make_world((200, 176), mix, seed=31)
    r = refsim.RefWorld(cells, zone)
    o = oraclesim.OracleWorld(cells, zone, defs=full_table())
    for t in range(10):
        r.tick(t)
        o.tick(t)
    (rx, ry), (ox, oy) = r.flow(), o.flow()
    assert np.abs(rx).sum() > 0 and np.abs(ry).sum() > 0
    assert rx.tobytes() == ox.tobytes() and ry.tobytes() == oy.tobytes()
    r.close()
    o.close()


def test_oracle_matches_c1_full_size_golden():
    """BASELINE.json configs[0] at its stated size (1984 x 1152 MaterialTestGenerator world, worlds.material_test_world): the
    plain-C restatement against what the reference's own lines produced offline (tests/golden/make_golden.py --c1), at the
    first checkpoint (100 ticks; the GPU test runs all 1000)."""
    from fallingsandsurvival_b200 import worlds
    with open(os.path.join(GOLDEN, "c1_material_test_1984x1152.json")) as f:
        gold = json.load(f)
    cells, zone = worlds.material_test_world(full_table())
    o = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    run_ticks(o, 100)
    want = gold["checkpoints"]["100"]
    c = o.cells()
    assert f"{oraclesim.hash_cells(c):016x}" == want["digest"]
    assert len(o.drain_particles()) == want["spawn_records"]
    assert {str(k): int(v) for k, v in zip(*np.unique(c["material"], return_counts=True))} == want["histogram"]
    o.close()
