"""Parity of the CUDA path (through the C ABI) against the oracle: bit-exact cells, particles, digests.

All tests need a B200 (`-m gpu`).  The oracle (oracle/fss_oracle.c) is itself pinned to the reference's own
lines by test_oracle_vs_reference.py; the golden digests / scenes in tests/golden were minted from the
reference build, so `test_gpu_matches_golden_*` compare the GPU to the reference directly.
"""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, MIXES, diff_report, full_table, make_world, run_ticks
from fallingsandsurvival_b200 import abi, worlds
from fallingsandsurvival_b200 import materials as M

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "tick_hashes.json")) as f:
    HASHES = json.load(f)


def gpu_world(cells, zone, **kw):
    from fallingsandsurvival_b200 import world
    return world.World(cells, zone, defs=full_table(), **kw)


def oracle_world(cells, zone, **kw):
    from oracle import oraclesim
    return oraclesim.OracleWorld(cells, zone, defs=full_table(), **kw)


def test_library_loaded_is_the_cuda_one():
    from fallingsandsurvival_b200 import world
    L = world.load()
    assert L.fss_abi_version() == abi.FSS_ABI_VERSION
    with open("/proc/self/maps") as f:
        assert "libfss_b200.so" in f.read()


@pytest.mark.parametrize("size", [(64, 64), (200, 176), (517, 301), (1031, 96)])
def test_upload_download_roundtrip(size):
    """ragged widths: not a multiple of the 256-column storage block, rectangles off the block grid"""
    cells, zone = make_world(size, "c3", seed=3)
    rng = np.random.default_rng(1)
    cells["moved"] = rng.integers(0, 2, cells.shape)
    cells["settle_count"] = rng.integers(0, 256, cells.shape)
    cells["fluid_amount_diff"] = rng.random(cells.shape, dtype=np.float32)
    g = gpu_world(cells, None)
    assert g.cells().tobytes() == cells.tobytes()
    # a rectangle in the middle
    w, h = size
    x0, y0, rw, rh = 5, 7, w - 11, h - 20
    assert g.download(x0, y0, rw, rh).tobytes() == np.ascontiguousarray(cells[y0:y0 + rh, x0:x0 + rw]).tobytes()
    patch = cells[y0:y0 + rh, x0:x0 + rw].copy()
    patch["material"] = M.STONE
    g.upload(x0, y0, patch)
    cells[y0:y0 + rh, x0:x0 + rw] = patch
    assert g.cells().tobytes() == cells.tobytes()
    from oracle import oraclesim
    assert g.hash() == oraclesim.hash_cells(cells)
    hist = g.histogram()
    ids, cnt = np.unique(cells["material"], return_counts=True)
    assert {int(i): int(c) for i, c in zip(ids, cnt)} == {i: int(c) for i, c in enumerate(hist) if c}
    g.close()


@pytest.mark.parametrize("mix", sorted(MIXES))
@pytest.mark.parametrize("particles", [False, True])
def test_gpu_matches_oracle_every_tick(mix, particles):
    cells, zone = make_world((200, 176), mix, seed=1000 + len(mix))
    g = gpu_world(cells, zone, particles=particles)
    o = oracle_world(cells, zone, particles=particles)
    for t in range(24):
        g.tick(t)
        o.tick(t)
        if t % 4 == 2:
            g.tick_temperature()
            o.tick_temperature()
        a, b = o.cells(), g.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    if particles:
        po, pg = o.drain_particles(), g.drain_particles()
        assert len(po) == len(pg)
        assert po.tobytes() == pg.tobytes()
    g.close()
    o.close()


def test_gpu_matches_oracle_substep_by_substep():
    """the finest granularity the ABI offers: one (iter, tk) phase at a time"""
    cells, zone = make_world((264, 208), "c3", seed=99)
    g = gpu_world(cells, zone, particles=True)
    o = oracle_world(cells, zone, particles=True)
    for t in range(3):
        for it in range(6):
            for tk in range(4):
                g.substep(t, it, tk)
                o.substep_rows(t, it, tk)
                a, b = o.cells(), g.cells()
                assert a.tobytes() == b.tobytes(), f"tick {t} iter {it} tk {tk}: " + diff_report(a, b)
    g.close()
    o.close()


@pytest.mark.parametrize("name", sorted(HASHES))
def test_gpu_matches_golden_digest(name):
    """digests minted from the reference's own lines (tests/golden/make_golden.py)"""
    gd = HASHES[name]
    cells, zone = make_world(tuple(gd["size"]), gd["mix"], gd["seed"])
    g = gpu_world(cells, zone, particles=gd["particles"])
    run_ticks(g, gd["ticks"])
    assert f"{g.hash():016x}" == gd["digest"]
    hist = {str(i): int(c) for i, c in enumerate(g.histogram()) if c}
    assert hist == gd["histogram"]
    g.close()


@pytest.mark.parametrize("name", ["fire_128x128_s8_t60", "gas_160x128_s7_t40", "sand_particles_128x160_s5_t30"])
def test_gpu_matches_golden_scene(name):
    gd = HASHES[name]
    z = np.load(os.path.join(GOLDEN, f"micro_{name}.npz"))
    g = gpu_world(z["before"], tuple(int(v) for v in z["zone"]), particles=gd["particles"])
    run_ticks(g, gd["ticks"])
    out = g.cells()
    assert out.tobytes() == z["after"].tobytes(), diff_report(z["after"], out)
    g.close()


@pytest.mark.parametrize("mix,size,ticks", [("c3", (1040, 784), 40), ("c2", (2080, 1056), 16), ("liquids", (784, 528), 60)])
def test_gpu_matches_oracle_larger(mix, size, ticks):
    cells, zone = make_world(size, mix, seed=4242)
    g = gpu_world(cells, zone, particles=(mix == "c3"))
    o = oracle_world(cells, zone, particles=(mix == "c3"))
    run_ticks(g, ticks)
    run_ticks(o, ticks)
    a, b = o.cells(), g.cells()
    assert a.tobytes() == b.tobytes(), diff_report(a, b)
    from oracle import oraclesim
    assert g.hash() == oraclesim.hash_cells(a)
    g.close()
    o.close()


@pytest.mark.parametrize("size,zone", [((264, 112), None), ((1288, 208), None), ((520, 400), (24, 48, 456, 288)), ((776, 336), (264, 16, 256, 64)),
                                       ((300, 176), (16, 32, 264, 96))])
def test_temperature_geometries(size, zone):
    """k_temperature (rows arrive as bulk copies into a three-stage ring): CTAs that own one 256-column block instead of two (an
    odd number of blocks per row), zones that start inside a CTA's columns, row bands of every length modulo the ring's three
    stages, CTAs with no column in the zone at all"""
    cells, dzone = make_world(size, "c3", seed=66)
    zone = zone or dzone
    rng = np.random.default_rng(1)
    cells["temperature"] = rng.integers(-1100, 1100, cells.shape).astype(np.int32) * (rng.random(cells.shape) < 0.7)
    g = gpu_world(cells, zone)
    o = oracle_world(cells, zone)
    for k in range(5):
        g.tick_temperature()
        o.tick_temperature()
        a, b = o.cells(), g.cells()
        assert a.tobytes() == b.tobytes(), f"pass {k}: " + diff_report(a["temperature"], b["temperature"])
    g.close()
    o.close()


def test_temperature_only_many_passes():
    cells, zone = make_world((520, 272), "c3", seed=6)
    rng = np.random.default_rng(0)
    cells["temperature"] = rng.integers(-1100, 1100, cells.shape).astype(np.int32) * (rng.random(cells.shape) < 0.6)
    g = gpu_world(cells, zone)
    o = oracle_world(cells, zone)
    for k in range(25):
        g.tick_temperature()
        o.tick_temperature()
        a, b = o.cells(), g.cells()
        assert a.tobytes() == b.tobytes(), f"pass {k}: " + diff_report(a["temperature"], b["temperature"])
    g.close()
    o.close()


@pytest.mark.parametrize("n_strips,particles", [(2, True), (3, True), (2, False), (3, False)])
def test_strips_equal_whole_world(n_strips, particles):
    """row strips with ghost rows exchanged after tk 1 / tk 3 give the same grid as one world (without particle emission the
    sub-steps run on the register-window kernels, fire through its pre-pass)"""
    from fallingsandsurvival_b200 import world
    cells, zone = make_world((264, 336), "c3", seed=11)
    whole = gpu_world(cells, zone, particles=particles)
    parts = world.StripWorld(cells, zone, n_strips, defs=full_table(), particles=particles)
    for t in range(16):
        whole.tick(t)
        parts.tick(t)
        if t % 4 == 2:
            whole.tick_temperature()
            parts.tick_temperature()
        a, b = whole.cells(), parts.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    assert whole.hash() == parts.hash()
    if particles:
        pa, pb = whole.drain_particles(), parts.drain_particles()
        assert pa.tobytes() == pb.tobytes()
    whole.close()
    parts.close()


def test_device_fill_equals_host_fill():
    from fallingsandsurvival_b200 import world
    size = (776, 300)
    for mix in (worlds.MIX_C2, worlds.MIX_C3):
        host = worlds.synthetic_world(size[0], size[1], mix, seed=0xF55, defs=full_table())
        g = world.World(None, None, defs=full_table(), size=size)
        g.fill_synthetic(mix, seed=0xF55)
        assert g.cells().tobytes() == host.tobytes()
        g.close()


def test_invariants_without_particles():
    """with particles off and no fire / steam / reactive material, SAND, SOLID and GAS cell counts
    are conserved and the liquid total never grows"""
    mix = [(M.TEST_SAND, 25), (M.WATER, 20), (M.COBBLE_STONE, 10), (M.GAS, 5)]
    cells, zone = make_world((520, 400), mix, seed=8)
    g = gpu_world(cells, zone)
    h0, f0 = g.histogram(), g.fluid_totals()
    prev = f0[M.WATER]
    for t in range(30):
        g.tick(t)
        f = g.fluid_totals()[M.WATER]
        assert f <= prev + 1e-3
        prev = f
    h1 = g.histogram()
    for m in (M.TEST_SAND, M.COBBLE_STONE, M.GAS, M.SOLID):
        assert h0[m] == h1[m], m
    assert prev > 0.9 * f0[M.WATER]
    g.close()


def test_error_paths():
    from fallingsandsurvival_b200 import world
    cells, zone = make_world((200, 176), "c2", seed=1)
    g = gpu_world(cells, None)
    with pytest.raises(world.FssError) as e:
        g.tick(0)  # no tick zone yet
    assert e.value.code == -5
    with pytest.raises(world.FssError) as e:
        g.set_tick_zone(17, 16, 160, 128)
    assert e.value.code == -2
    with pytest.raises(world.FssError) as e:
        g.set_tick_zone(16, 16, 184, 128)  # reaches into the margin
    assert e.value.code == -2
    with pytest.raises(world.FssError) as e:
        g.download(0, 0, 201, 10)
    assert e.value.code == -6
    g.close()
    with pytest.raises(world.FssError):
        world.World(None, None, size=(8, 8))


def test_nccl_strips_equal_single_gpu():
    """one process per GPU over NCCL (needs >= 2 visible GPUs; the 1-GPU box covers the same schedule through
    test_strips_equal_whole_world's peer copies)"""
    import subprocess
    import sys

    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    n = 2 if n < 4 else 4
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_nccl_strips.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


@pytest.mark.parametrize("mix", ["settled", "c3", "liquids", "sand"])
def test_early_out_is_result_preserving(mix):
    """skipping chunks verified settled (the dirty-rect replacement) must not change a single bit: early-out on ==
    early-out off == oracle, over a run long enough for large parts of the world to settle, with host edits
    (uploads) in the middle that must wake the neighbourhood up again"""
    cells, zone = make_world((392, 304), mix, seed=77)
    on = gpu_world(cells, zone, particles=True, early_out=True)
    off = gpu_world(cells, zone, particles=True, early_out=False)
    o = oracle_world(cells, zone, particles=True)
    rng = np.random.default_rng(5)
    for t in range(70):
        for w in (on, off, o):
            w.tick(t)
            if t % 4 == 2:
                w.tick_temperature()
        if t in (30, 50):  # a brush stroke of sand and one of water through the host mirror
            x0, y0 = int(rng.integers(40, 300)), int(rng.integers(40, 200))
            patch = on.download(x0, y0, 24, 12)
            assert patch.tobytes() == off.download(x0, y0, 24, 12).tobytes()
            patch["material"] = M.TEST_SAND if t == 30 else M.WATER
            patch["fluid_amount"], patch["fluid_amount_diff"], patch["moved"], patch["settle_count"] = 2.0, 0.0, 0, 0
            on.upload(x0, y0, patch)
            off.upload(x0, y0, patch)
            o.view()[y0:y0 + 12, x0:x0 + 24] = patch
        if t % 10 == 9 or t in (30, 31, 50, 51):
            a, b, c = o.cells(), on.cells(), off.cells()
            assert a.tobytes() == c.tobytes(), f"tick {t} (early-out off): " + diff_report(a, c)
            assert a.tobytes() == b.tobytes(), f"tick {t} (early-out on): " + diff_report(a, b)
    assert on.drain_particles().tobytes() == o.drain_particles().tobytes()
    for w in (on, off, o):
        w.close()


def test_early_out_sparse_world_long_run():
    """the configs[4] shape in small: stone floor, sparse sand, a few active boxes; 200 ticks, on == off"""
    from fallingsandsurvival_b200 import world
    size = (1056, 784)
    zone = worlds.default_zone(*size)
    ws = []
    for eo in (True, False):
        g = world.World(None, zone, defs=full_table(), size=size, early_out=eo)
        g.fill_synthetic(worlds.MIX_C2, seed=3, solid_from_row=int(size[1] * 0.3), n_boxes=3, box_size=128,
                         sparse_material=M.TEST_SAND, sparse_per_mille=2)
        ws.append(g)
    assert ws[0].hash() == ws[1].hash()
    for t in range(200):
        for g in ws:
            g.tick(t)
        if t % 50 == 49:
            assert ws[0].hash() == ws[1].hash(), t
    assert ws[0].cells().tobytes() == ws[1].cells().tobytes()
    for g in ws:
        g.close()


def material_test_world(size=(520, 400)):
    """A small version of BASELINE configs[0]: the MaterialTestGenerator layout (cobble band, walled containers,
    MaterialTestGenerator.cpp:32-48) with painted blobs of sand, water, steam and fire on soft dirt."""
    w, h = size
    cells = worlds.synthetic_world(w, h, [], seed=1)
    defs = full_table()
    key = worlds.cell_key_grid(1, w, h)
    mat = np.zeros((h, w), dtype=np.uint16)
    mat[int(h * 0.78):int(h * 0.9), :] = M.COBBLE_STONE
    for k, m in enumerate([31, 33, 35, 38]):  # containers filled with one random `Mat_N` each
        x0 = 40 + k * 110
        mat[200:290, x0:x0 + 4] = M.COBBLE_STONE
        mat[200:290, x0 + 76:x0 + 80] = M.COBBLE_STONE
        mat[286:290, x0:x0 + 80] = M.COBBLE_STONE
        mat[210:270, x0 + 4:x0 + 76] = m
    mat[40:100, 60:130] = M.TEST_SAND
    mat[40:100, 180:260] = M.WATER
    mat[120:170, 300:360] = M.STEAM
    mat[100:110, 380:470] = M.SOFT_DIRT
    mat[96:100, 390:460] = M.FIRE
    mat[30:36, 330:400] = M.GOLD_ORE
    mat[:16, :] = mat[-16:, :] = M.SOLID
    mat[:, :16] = mat[:, -16:] = M.SOLID
    cells = worlds.cells_of_material(defs, [M.synthetic_texture(i) for i in range(M.N_TEXTURES)], mat, key)
    cells["temperature"][mat == M.GOLD_ORE] = 700
    return cells, worlds.default_zone(w, h)


def test_thousand_ticks_bit_exact():
    """the north star's acceptance line: bit-exact grids against the reference schedule after 1000 ticks"""
    cells, zone = material_test_world()
    g = gpu_world(cells, zone, particles=True)
    o = oracle_world(cells, zone, particles=True)
    for t in range(1000):
        g.tick(t)
        o.tick(t)
        if t % 4 == 2:
            g.tick_temperature()
            o.tick_temperature()
        if t in (99, 499, 999):
            a, b = o.cells(), g.cells()
            assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    assert o.drain_particles().tobytes() == g.drain_particles(cap=1 << 22).tobytes()
    g.close()
    o.close()


def test_c1_full_size_thousand_ticks_against_the_reference_build():
    """BASELINE.json configs[0] at its stated size: the MaterialTestGenerator world (worlds.material_test_world, 1984 x 1152,
    tick zone 1920 x 1088), 1000 ticks with tickTemperature every 4th, particles emitted -- against the digests, spawn-record
    counts and material histograms the reference's OWN lines (oracle/_ref) produced offline
    (tests/golden/c1_material_test_1984x1152.json, minted by tests/golden/make_golden.py --c1)."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_material_test_1984x1152.json")) as f:
        gold = json.load(f)
    cells, zone = worlds.material_test_world(full_table())
    assert list(cells.shape[::-1]) == gold["size"] and list(zone) == gold["zone"]
    g = gpu_world(cells, zone, particles=True, particle_capacity=1 << 22)
    spawned = 0
    for t in range(1000):
        g.tick(t)
        if t % 4 == 2:
            g.tick_temperature()
        if t % 20 == 19:
            spawned += len(g.drain_particles(cap=1 << 22))
        if str(t + 1) in gold["checkpoints"]:
            want = gold["checkpoints"][str(t + 1)]
            assert f"{g.hash():016x}" == want["digest"], f"tick {t + 1}: digest differs from the reference build's"
            assert spawned == want["spawn_records"], f"tick {t + 1}: {spawned} spawn records, the reference build pushed {want['spawn_records']}"
            hist = g.histogram()
            assert {str(m): int(c) for m, c in enumerate(hist) if c} == want["histogram"]
    g.close()


@pytest.mark.parametrize("mix", ["c2", "c3"])
def test_full_size_kernel_families_agree(mix, monkeypatch):
    """c2 / c3 (BASELINE configs[1] / [2]: + fire, lava, steam, gold ore) at 16384 x 16384, where the oracle is out of reach: the
    register-window kernels (general walk, fire through its pre-pass) and the lock-step kernels (FSS_NO_WINDOW) are two
    implementations of the same rules and must leave the same grid, tick after tick, temperature included"""
    import torch
    from fallingsandsurvival_b200 import world
    free, _ = torch.cuda.mem_get_info()
    S = 16384 if free > 60e9 else 4096
    size = (S + 32, S + 32)
    zone = (16, 16, S, S)
    a = world.World(None, zone, defs=full_table(), size=size)
    b = world.World(None, zone, defs=full_table(), size=size)
    for w in (a, b):
        w.fill_synthetic(MIXES[mix], seed=0xF55)
    assert a.hash() == b.hash()
    for t in range(4):
        a.tick(t)
        a.sync()
        monkeypatch.setenv("FSS_NO_WINDOW", "1")
        b.tick(t)
        b.sync()
        monkeypatch.delenv("FSS_NO_WINDOW")
        if t == 2:
            a.tick_temperature()
            b.tick_temperature()
        assert a.hash() == b.hash(), f"tick {t}: window kernels {a.hash():016x} != lock-step kernels {b.hash():016x}"
    la, lb = a.stats().kernel_launches, b.stats().kernel_launches
    assert mix == "c2" or la > lb  # (c3: the window path launched its fire pre-passes: the two worlds did take different kernels)
    a.close()
    b.close()


def test_full_size_tick_host_equals_resident_ticks():
    """fss_tick_host at the benchmark's size (16384 x 16384: 16 bands of 1024 rows, 5.4 GB each way per tick): the grid that comes
    back to the host, and the one left on the device, equal the grid of a world that stayed resident"""
    import torch
    from fallingsandsurvival_b200 import world
    free, _ = torch.cuda.mem_get_info()
    S = 16384 if free > 60e9 else 4096
    size = (S + 32, S + 32)
    zone = (16, 16, S, S)
    a = world.World(None, zone, defs=full_table(), size=size)
    b = world.World(None, zone, defs=full_table(), size=size)
    for w in (a, b):
        w.fill_synthetic(worlds.MIX_C2, seed=0xF55)
    host = torch.empty((size[1], size[0], 20), dtype=torch.uint8, pin_memory=True).numpy().view(abi.CELL_DTYPE).reshape(size[1], size[0])
    a.download(0, 0, size[0], size[1], out=host)
    for t in range(2):
        a.tick_host(host, t)
        b.tick(t)
        assert a.hash() == b.hash(), f"tick {t}: device grid after fss_tick_host differs from the resident world"
    back = torch.empty((size[1], size[0], 20), dtype=torch.uint8, pin_memory=True).numpy().view(abi.CELL_DTYPE).reshape(size[1], size[0])
    b.download(0, 0, size[0], size[1], out=back)
    assert np.array_equal(host.view(np.uint8), back.view(np.uint8)), "the grid fss_tick_host handed back differs from the resident world's"
    a.close()
    b.close()


def test_full_size_mostly_settled_world_early_out():
    """BASELINE configs[4] at its stated size (32768 x 32768, the lower nine tenths stone with pockets of the c2 mix, sparse sand
    above): the settled-chunk flags and the compacted worklist skip most of the world; the grid must equal the one every
    chunk visit produced (FSS_NO_EARLY_OUT), tick after tick"""
    import torch
    from fallingsandsurvival_b200 import world
    free, _ = torch.cuda.mem_get_info()
    S = 32768 if free > 90e9 else 4096
    size = (S + 32, S + 32)
    zone = (16, 16, S, S)
    on = world.World(None, zone, defs=full_table(), size=size, early_out=True)
    off = world.World(None, zone, defs=full_table(), size=size, early_out=False)
    for w in (on, off):
        w.fill_synthetic(worlds.MIX_C2, seed=0xF55, border=16, solid_from_row=int(size[1] * 0.1), n_boxes=64 if S >= 4096 else 4, box_size=256,
                         sparse_material=M.TEST_SAND, sparse_per_mille=1)
    for t in range(8):
        on.tick(t)
        off.tick(t)
        assert on.hash() == off.hash(), f"tick {t}"
    assert on.stats().kernel_launches > off.stats().kernel_launches  # (the worklist kernels ran)
    on.close()
    off.close()


def test_full_size_properties():
    """at BASELINE's 16384 x 16384 the oracle is out of reach; check what does not need it: the digest of a 2-strip
    run equals the undivided run, early-out on equals off, and SAND / SOLID counts are conserved"""
    import torch
    from fallingsandsurvival_b200 import world
    free, _ = torch.cuda.mem_get_info()
    S = 16384 if free > 60e9 else 4096
    size = (S + 32, S + 32)
    zone = (16, 16, S, S)
    digests, hists = [], []
    for mode in ("whole", "off", "strips"):
        if mode == "strips":
            from fallingsandsurvival_b200 import strips
            parts = [world.World(None, zone, defs=full_table(), size=size, strip=s) for s in strips.plan_strips(size[1], zone, 2)]
            ws = parts
        else:
            ws = [world.World(None, zone, defs=full_table(), size=size, early_out=(mode != "off"))]
        for p in ws:
            p.fill_synthetic(worlds.MIX_C2, seed=0xF55)
        if not hists:
            hists.append(ws[0].histogram())
        L = ws[0].L
        for t in range(3):
            if mode == "strips":
                from fallingsandsurvival_b200 import strips
                for op, it, tk in strips.tick_schedule():
                    if op == "substep":
                        for p in ws:
                            p.substep(t, it, tk)
                    else:
                        assert L.fss_exchange_halo_peer(ws[0].h_, ws[1].h_, tk) == 0
            else:
                ws[0].tick(t)
        digests.append(sum(p.hash() for p in ws) & 0xFFFFFFFFFFFFFFFF)
        if mode == "whole":
            hists.append(ws[0].histogram())
        for p in ws:
            p.close()
    assert digests[0] == digests[1] == digests[2], [f"{d:016x}" for d in digests]
    for m in (M.TEST_SAND, M.COBBLE_STONE, M.SOLID):
        assert hists[0][m] == hists[1][m]


@pytest.mark.parametrize("change", [(5, 0), (-3, 7), (128, -64), (-300, 11), (0, -1), (2000, 0)])
def test_shift_grid_matches_reference_loop(change):
    """World::tickChunks (world.cpp:2617-2630): tiles[new] = tiles[old] where both are inside the grid, in an order that
    never reads an already overwritten cell; everything else keeps its content"""
    cells, zone = make_world((517, 301), "c3", seed=13)
    g = gpu_world(cells, zone)
    cx, cy = change
    g.shift_grid(cx, cy)
    h, w = cells.shape
    want = cells.copy()
    ys, xs = np.mgrid[0:h, 0:w]
    oy, ox = ys - cy, xs - cx
    ok = (ox >= 0) & (ox < w) & (oy >= 0) & (oy < h)
    want[ok] = cells[oy[ok], ox[ok]]
    assert g.cells().tobytes() == want.tobytes()
    # a device-resident particle list moves with the grid (world.cpp:2729-2731)
    from helpers import random_particles
    parts = random_particles(4, 300, (517, 301))
    g2 = gpu_world(cells, zone)
    g2.add_particles(parts)
    g2.shift_grid(cx, cy)
    want_p = parts.copy()
    want_p["x"] = parts["x"] + np.float32(cx)
    want_p["y"] = parts["y"] + np.float32(cy)
    assert g2.particle_list().tobytes() == want_p.tobytes()
    assert g2.cells().tobytes() == want.tobytes()
    g2.close()
    # the world keeps ticking correctly afterwards (flags were reset): compare with the oracle on the shifted grid
    o = oracle_world(want, zone)
    for t in range(4):
        g.tick(t)
        o.tick(t)
    assert g.cells().tobytes() == o.cells().tobytes()
    g.close()
    o.close()


@pytest.mark.parametrize("n_strips", [2, 3])
def test_shift_grid_on_strips_equals_whole_world(n_strips):
    """World::tickChunks' shift (world.cpp:2617-2630) on a strip-sharded world: rows that cross a strip boundary come from the
    neighbouring strip (fss_shift_grid_peer here; the NCCL form of the same exchange is driven by tests/run_nccl_strips.py).
    Shifts in all directions, larger than the ghost rows, ticks in between; a shift longer than a strip is refused whole."""
    from fallingsandsurvival_b200 import world
    cells, zone = make_world((264, 400), "c3", seed=12)
    whole = gpu_world(cells, zone, particles=True)
    parts = world.StripWorld(cells, zone, n_strips, defs=full_table(), particles=True)
    t = 0
    for cx, cy in [(3, 5), (-7, -40), (0, 33), (11, -1), (-2, 64), (0, 0), (40, -70)]:
        whole.shift_grid(cx, cy)
        parts.shift_grid(cx, cy)
        a, b = whole.cells(), parts.cells()
        assert a.tobytes() == b.tobytes(), f"shift ({cx}, {cy}): " + diff_report(a, b)
        for _ in range(2):
            whole.tick(t)
            parts.tick(t)
            t += 1
        a, b = whole.cells(), parts.cells()
        assert a.tobytes() == b.tobytes(), f"ticks after shift ({cx}, {cy}): " + diff_report(a, b)
    before = parts.cells()
    with pytest.raises(world.FssError) as e:
        parts.shift_grid(0, 300)
    assert e.value.code == -6
    assert parts.cells().tobytes() == before.tobytes()
    whole.close()
    parts.close()


@pytest.mark.parametrize("table", ["reference", "dense sand", "sand turns to stone"])
def test_fire_paths_match_oracle(table):
    """Fire without particle emission: with the reference's table the fire cells' visits are decided by k_fire_prepass and applied by
    the register-window kernel; a SAND material denser than fire (it displaces fire cells) or a SAND reaction with a SOLID product
    switch that off (fss_abi.cu `fire_static`) and the sub-steps in which fire acts run on the lock-step kernel.  All three against
    the oracle, every tick, with and without the settled-chunk flags."""
    import copy
    from fallingsandsurvival_b200 import world
    from oracle import oraclesim
    defs = copy.deepcopy(full_table())
    if table == "dense sand":
        defs[M.TEST_SAND].density = 30.0
    elif table == "sand turns to stone":
        defs[M.GOLD_ORE].reactions = [(abi.REACT_TEMPERATURE_ABOVE, 512, M.COBBLE_STONE)]
    mix = [(M.FIRE, 6), (M.SOFT_DIRT, 18), (M.COBBLE_STONE, 8), (M.TEST_SAND, 15), (M.GOLD_ORE, 5), (M.WATER, 12), (M.LAVA, 3), (M.STEAM, 3)]
    cells, zone = make_world((328, 272), mix, seed=91, defs=defs)
    cells["temperature"][cells["material"] == M.GOLD_ORE] = 600
    for early_out in (False, True):
        g = world.World(cells, zone, defs=defs, early_out=early_out)
        o = oraclesim.OracleWorld(cells, zone, defs=defs)
        for t in range(20):
            g.tick(t)
            o.tick(t)
            a, b = o.cells(), g.cells()
            assert a.tobytes() == b.tobytes(), f"{table}, early-out {early_out}, tick {t}: " + diff_report(a, b)
        g.close()
        o.close()


def test_kernel_specialisation_follows_the_materials_present():
    """k_substep<K1_LIQUID / K1_NOGAS / K1_GENERAL> are picked per iter from the set of materials that can be in the world.
    Start with water + stone only (every iter runs the liquid-only kernel), then paint sand, then steam and fire through the
    host mirror: each upload must widen the set, and the grid must stay bit-identical to the oracle throughout."""
    cells, zone = make_world((328, 272), [(M.WATER, 30), (M.LAVA, 5), (M.COBBLE_STONE, 15)], seed=41)
    g = gpu_world(cells, zone, particles=True)
    o = oracle_world(cells, zone, particles=True)

    def paint(t, mat, x0, y0, w=40, h=24, temp=None):
        patch = g.download(x0, y0, w, h)
        patch["material"] = mat
        patch["fluid_amount"], patch["fluid_amount_diff"], patch["moved"], patch["settle_count"] = 2.0, 0.0, 0, 0
        if temp is not None:
            patch["temperature"] = temp
        g.upload(x0, y0, patch)
        o.view()[y0:y0 + h, x0:x0 + w] = patch

    for t in range(40):
        if t == 8:
            paint(t, M.TEST_SAND, 60, 40)
        if t == 16:
            paint(t, M.STEAM, 150, 180)
            paint(t, M.GOLD_ORE, 200, 40, temp=700)  # reacts into molten gold (a liquid)
        if t == 24:
            paint(t, M.SOFT_DIRT, 240, 100)
            paint(t, M.FIRE, 244, 96, w=30, h=4)
        g.tick(t)
        o.tick(t)
        if t % 4 == 2:
            g.tick_temperature()
            o.tick_temperature()
        if t % 4 == 3:
            a, b = o.cells(), g.cells()
            assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    assert o.drain_particles().tobytes() == g.drain_particles().tobytes()
    g.close()
    o.close()


def test_plain_c_caller_matches_python_binding(tmp_path):
    """tests/c/abi_smoke.c drives the C ABI directly (create, materials, fill, rect edit, tick, temperature, digest);
    the same sequence through the Python binding, checked against the oracle, must give the same digest"""
    import re
    import subprocess

    from test_abi import build_c_caller
    from fallingsandsurvival_b200 import world
    exe = build_c_caller(tmp_path)
    r = subprocess.run([str(exe), "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c_digest = int(re.search(r"digest ([0-9a-f]{16})", r.stdout).group(1), 16)
    c_particles = int(re.search(r"particles (\d+)", r.stdout).group(1))

    A, SO, SA, LQ, G = abi.PHYS_AIR, abi.PHYS_SOLID, abi.PHYS_SAND, abi.PHYS_SOUP, abi.PHYS_GAS
    D = M.MaterialDef
    defs = [D(0, "air", A, 0, 255, 0.0, 0, conduction_self=0.8, conduction_other=0.8, color=0, create_base=0),
            D(1, "solid", SO, 0, 255, 1.0, 0, create_base=0x808080, color=0),
            D(2, "sand", SA, 20, 255, 10.0, 2, color=0, create_kind=abi.CREATE_RAND, create_base=(220 << 16) + (155 << 8) + 100,
              create_rand_mod=30, create_rand_shift=8, create_rand_site=abi.SITE_TILES + 16),
            D(3, "water", LQ, 0, 255, 1.5, 6, color=0, create_base=0x00B69F, create_temperature=-1023),
            D(4, "steam", G, 0, 255, -1.0, 1, color=0, flags=abi.MAT_IS_STEAM, create_base=0x666666)]
    size, zone = (264, 208), (16, 16, 232, 160)
    g = world.World(None, zone, defs=defs, textures=[], size=size, particles=True, condense_material=3)
    g.fill_synthetic([(2, 25), (3, 20), (1, 10), (4, 5)], seed=7, border=16, border_material=1)
    patch = g.download(100, 60, 20, 10)
    patch["material"], patch["fluid_amount"], patch["fluid_amount_diff"], patch["moved"] = 2, 2.0, 0.0, 0
    g.upload(100, 60, patch)
    from oracle import oraclesim
    o = oraclesim.OracleWorld(g.cells(), zone, defs=defs, textures=[], particles=True, condense_material=3)
    run_ticks(g, 8)
    run_ticks(o, 8)
    assert g.cells().tobytes() == o.cells().tobytes()
    assert g.hash() == c_digest
    assert len(g.drain_particles()) == c_particles
    g.close()
    o.close()


def test_tall_world_beyond_grid_dim_limits():
    """a 66000-row world: rows beyond gridDim.y's 65535 in every kernel that walks rows (fill, upload, download,
    shift, hash), and a couple of ticks against the oracle"""
    from fallingsandsurvival_b200 import world
    size = (264, 66016)
    zone = worlds.default_zone(*size)
    g = world.World(None, zone, defs=full_table(), size=size)
    g.fill_synthetic(worlds.MIX_C2, seed=9)
    host = worlds.synthetic_world(size[0], size[1], worlds.MIX_C2, seed=9, defs=full_table())
    got = g.cells()
    assert got.tobytes() == host.tobytes()
    g2 = gpu_world(host, zone)  # the upload path
    from oracle import oraclesim
    assert g.hash() == g2.hash() == oraclesim.hash_cells(host)
    o = oracle_world(host, zone)
    for t in range(2):
        g.tick(t)
        g2.tick(t)
        o.tick(t)
    a = o.cells()
    assert a.tobytes() == g.cells().tobytes() == g2.cells().tobytes()
    g.shift_grid(0, 65540)
    want = a.copy()
    want[65540:] = a[:size[1] - 65540]
    assert g.cells().tobytes() == want.tobytes()
    for w in (g, g2, o):
        w.close()


@pytest.mark.parametrize("device_particles", [False, True, "render"])
def test_cpp_world_mirror_matches_oracle(tmp_path, device_particles):
    """tests/cpp/world_shim_test.cpp plays a short game script against fss::World (include/fss_world.hpp); the oracle
    replays the same script (same initial grid, same brush stroke) and must produce the same tiles[].  With
    device_particles the script also runs World::tickParticles every tick and throws one particle by hand."""
    import subprocess

    from test_abi import build_cpp_shim_test
    exe = build_cpp_shim_test(tmp_path)
    render = device_particles == "render"  # + World::dirty and the dirty -> pixels job on the device
    r = subprocess.run([str(exe), str(tmp_path)] + (["render"] if render else ["particles"] if device_particles else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    W, H = 328, 240
    initial = np.fromfile(tmp_path / "initial.bin", dtype=abi.CELL_DTYPE).reshape(H, W)
    final = np.fromfile(tmp_path / "final.bin", dtype=abi.CELL_DTYPE).reshape(H, W)
    A, SO, SA, LQ, G = abi.PHYS_AIR, abi.PHYS_SOLID, abi.PHYS_SAND, abi.PHYS_SOUP, abi.PHYS_GAS
    D = M.MaterialDef
    defs = [D(0, "_AIR", A, 0, 255, 0.0, 0, color=0, conduction_self=0.8, conduction_other=0.8),
            D(1, "Cobblestone", SO, 0, 255, 1.0, 0, color=0x808080, conduction_self=0.01, conduction_other=0.4),
            D(2, "Test Sand", SA, 20, 255, 10.0, 2, color=0xDC9B64),
            D(3, "Water", LQ, 0, 255, 1.5, 6, color=0x00B69F, create_temperature=-1023),
            D(4, "Steam", G, 0, 255, -1.0, 1, color=0x666666, flags=abi.MAT_IS_STEAM),
            D(5, "Lava", LQ, 0, 255, 2.0, 1, color=0xFF7C00, add_temp=2, conduction_self=0.5, conduction_other=0.7, create_temperature=1024)]
    from oracle import oraclesim
    o = oraclesim.OracleWorld(initial, (16, 16, W - 32, H - 48), defs=defs, textures=[], particles=True, condense_material=3)
    if render:
        looks_a, looks_e = np.full(6, 255, dtype=np.uint8), np.zeros(6, dtype=np.uint32)
        looks_a[3], looks_a[5], looks_e[5] = 0x80, 0xC0, 0xC0FF7C00
        o.set_material_looks(looks_a, looks_e)
        o.mark_dirty(0, 0, W, H)  # World::init painted every cell through setTile
    for t in range(12):
        if t == 5:
            patch = np.zeros((12, 40), dtype=abi.CELL_DTYPE)
            patch["material"], patch["color"], patch["fluid_amount"] = 2, 0xABCDEF, 2.0
            o.view()[40:52, 100:140] = patch
            if render:
                o.mark_dirty(100, 40, 40, 12)
        o.tick(t)
        if device_particles:
            if t == 3:
                p = np.zeros(1, dtype=abi.LIVE_PARTICLE_DTYPE)
                p["x"], p["y"], p["vx"], p["vy"], p["ay"], p["fade_time"] = 150.5, 60.25, 1.5, -2.0, 0.1, 60
                p["tile"]["material"], p["tile"]["color"], p["tile"]["fluid_amount"] = 2, 0x112233, 2.0
                o.add_particles(p)
            o.tick_particles()
        if t % 4 == 2:
            o.tick_temperature()
        if render and t % 3 == 2:
            moving, info = o.render_dirty()
    want = o.cells()
    assert final.tobytes() == want.tobytes(), diff_report(want, final)
    if render:
        layers = np.fromfile(tmp_path / "layers.bin", dtype=np.uint8).reshape(4, H, W, 4)
        for k in range(4):
            assert layers[k].tobytes() == o.pixels[k].tobytes(), f"layer {k}: {int((layers[k] != o.pixels[k]).any(axis=2).sum())} pixels differ"
        assert "moving " + " ".join(str(int(v)) for v in moving) + " had %d %d %d" % info in r.stdout
    assert f"digest {oraclesim.hash_cells(want):016x}" in r.stdout
    if device_particles:
        got = np.fromfile(tmp_path / "particles.bin", dtype=abi.LIVE_PARTICLE_DTYPE)
        assert got.tobytes() == o.particle_list().tobytes()
    o.close()


@pytest.mark.parametrize("size,zone,mix,seed", [
    ((300, 210), (24, 19, 248, 160), "c3", 1),      # zone origin off the 16-cell grid, y odd
    ((517, 301), (40, 33, 432, 224), "liquids", 2),  # ragged world width, zone well inside
    ((1031, 130), (16, 17, 992, 96), "sand", 3),     # wide and flat: one band of 3 chunk rows
    ((136, 700), (64, 250, 56, 416), "gas", 4),      # tall and narrow, zone narrower than one warp
    ((264, 264), (120, 120, 24, 32), "fire", 5),     # the smallest legal zone class: 3 x 2 chunks
    ((776, 520), (256, 16, 496, 480), "c3", 6),      # zone starting on a 256-column storage block
])
def test_odd_geometries(size, zone, mix, seed):
    """tick zones that are not the default one: offsets, sizes and aspect ratios that exercise the block / warp / chunk
    boundaries of the permuted layout.  Cells outside the zone are live material too (no solid border)."""
    w, h = size
    cells = worlds.synthetic_world(w, h, MIXES[mix], seed=seed, defs=full_table(), border=0)
    g = gpu_world(cells, zone, particles=True)
    o = oracle_world(cells, zone, particles=True)
    for t in range(20):
        g.tick(t)
        o.tick(t)
        if t % 4 == 2:
            g.tick_temperature()
            o.tick_temperature()
        if t % 5 == 4:
            a, b = o.cells(), g.cells()
            assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    assert o.drain_particles().tobytes() == g.drain_particles().tobytes()
    g.close()
    o.close()


@pytest.mark.parametrize("origin", [(64, 48), (-40, 100), (300, -20), (450, 250)])
def test_chunk_records_merge_and_read(origin):
    """the reference's on-disk cell records (MaterialInstanceData, Chunk.hpp:27-31) merged into / read back from the device
    grid like Chunk::read + World::frame / Chunk::write do, including chunks that hang over the world's edge"""
    cells, zone = make_world((520, 336), "c3", seed=17)
    g = gpu_world(cells, zone)
    rng = np.random.default_rng(3)
    recs = np.zeros((128, 128), dtype=abi.RECORD_DTYPE)
    recs["index"] = rng.choice([M.AIR, M.STONE, M.TEST_SAND, M.WATER, M.LAVA, M.STEAM], size=recs.shape)
    recs["color"] = rng.integers(0, 1 << 32, recs.shape, dtype=np.uint64).astype(np.uint32)
    recs["temperature"] = rng.integers(-1023, 1025, recs.shape)
    x0, y0 = origin
    g.merge_chunk_records(x0, y0, recs)
    want = cells.copy()
    h, w = cells.shape
    xs0, ys0, xs1, ys1 = max(x0, 0), max(y0, 0), min(x0 + 128, w), min(y0 + 128, h)
    sub = recs[ys0 - y0:ys1 - y0, xs0 - x0:xs1 - x0]
    blk = np.zeros(sub.shape, dtype=abi.CELL_DTYPE)
    blk["material"], blk["color"], blk["temperature"], blk["fluid_amount"] = sub["index"], sub["color"], sub["temperature"], 2.0
    want[ys0:ys1, xs0:xs1] = blk
    assert g.cells().tobytes() == want.tobytes()
    back = g.read_chunk_records(x0, y0, 128, 128)
    exp = np.zeros((128, 128), dtype=abi.RECORD_DTYPE)
    exp[ys0 - y0:ys1 - y0, xs0 - x0:xs1 - x0] = sub
    for f in ("index", "color", "temperature"):
        assert (back[f] == exp[f]).all(), f
    # and the merged world ticks like the oracle (the new materials widened the kernel selection)
    o = oracle_world(want, zone)
    for t in range(4):
        g.tick(t)
        o.tick(t)
    assert g.cells().tobytes() == o.cells().tobytes()
    g.close()
    o.close()


@pytest.mark.parametrize("mix", ["c2", "liquids", "c3"])
def test_flow_side_output_matches_oracle(mix):
    """World::flowX / flowY (FSS_TRACK_FLOW): bit-identical float accumulators, and the consumer's reset"""
    cells, zone = make_world((264, 208), mix, seed=23)
    g = gpu_world(cells, zone, track_flow=True)
    o = oracle_world(cells, zone)
    for t in range(12):
        g.tick(t)
        o.tick(t)
    (gx, gy), (ox, oy) = g.flow(), o.flow()
    assert np.abs(ox).sum() > 0 and np.abs(oy).sum() > 0
    assert gx.tobytes() == ox.tobytes() and gy.tobytes() == oy.tobytes()
    assert g.cells().tobytes() == o.cells().tobytes()
    g.clear_flow()
    gx, gy = g.flow()
    assert not gx.any() and not gy.any()
    from fallingsandsurvival_b200 import world
    plain = gpu_world(cells, zone)
    with pytest.raises(world.FssError) as e:
        plain.flow()
    assert e.value.code == -5
    for w in (g, o, plain):
        w.close()


# ---- World::particles on the device: fss_tick_particles against the oracle's serial loop ----------------
def _same_particles(t, g, o):
    a, b = g.particle_list(), o.particle_list()
    assert len(a) == len(b), f"tick {t}: {len(a)} particles on the device, {len(b)} in the oracle"
    if a.tobytes() != b.tobytes():
        bad = [i for i in range(len(a)) if a[i].tobytes() != b[i].tobytes()]
        raise AssertionError(f"tick {t}: {len(bad)} particles differ, first at {bad[0]}: {a[bad[0]]} vs {b[bad[0]]}")
    ca, cb = g.cells(), o.cells()
    assert ca.tobytes() == cb.tobytes(), f"tick {t}: " + diff_report(ca, cb)


def test_device_particles_scene_matches_oracle_and_golden():
    """the hand-made scene that drives every branch of World::tickParticles; also against the hashes minted from the
    reference's own lines"""
    import hashlib
    from helpers import particle_scene
    with open(os.path.join(GOLDEN, "particle_scene.json")) as f:
        rows = json.load(f)
    cells, zone, parts = particle_scene()
    g = gpu_world(cells, zone, particles=True)
    o = oracle_world(cells, zone, particles=True)
    g.add_particles(parts)
    o.add_particles(parts)
    assert g.particles_count() == len(parts)
    assert g.particle_list().tobytes() == parts.tobytes()
    for row in rows:
        t = row["tick"]
        g.tick(t)
        o.tick(t)
        g.tick_particles()
        o.tick_particles()
        _same_particles(t, g, o)
        assert g.particles_count() == row["n"]
        assert hashlib.sha256(g.particle_list().tobytes()).hexdigest() == row["particles"]
        assert hashlib.sha256(g.cells().tobytes()).hexdigest() == row["cells"]
    assert len(g.drain_particles()) == 0  # the spawn records were adopted
    g.close()
    o.close()


@pytest.mark.parametrize("mix", ["c3", "sand", "liquids", "fire"])
def test_device_particles_with_tick_match_oracle(mix):
    """World::tick feeds the list, World::tickParticles lands it again; temperature every 4th tick"""
    cells, zone = make_world((264, 240), mix, seed=3000 + len(mix))
    g = gpu_world(cells, zone, particles=True)
    o = oracle_world(cells, zone, particles=True)
    for t in range(40):
        g.tick(t)
        o.tick(t)
        g.tick_particles()
        o.tick_particles()
        if t % 4 == 2:
            g.tick_temperature()
            o.tick_temperature()
        _same_particles(t, g, o)
    g.close()
    o.close()


def test_device_particles_every_other_tick_and_host_pushes():
    """spawn records of two ticks adopted at once; host pushes land behind them; early-out off gives the same"""
    from helpers import particle_scene
    cells, zone, parts = particle_scene(seed=9)
    worlds_ = [gpu_world(cells, zone, particles=True), gpu_world(cells, zone, particles=True, early_out=False),
               oracle_world(cells, zone, particles=True)]
    for t in range(12):
        for w in worlds_:
            w.tick(t)
            if t % 2 == 1:
                w.tick_particles()
            if t == 3:
                w.add_particles(parts[:100])
            if t == 6:
                w.add_particles(parts[100:])
        if t % 2 == 1:
            _same_particles(t, worlds_[0], worlds_[2])
            _same_particles(t, worlds_[1], worlds_[2])
    for w in worlds_:
        w.close()


def test_device_particles_dense_rain():
    """many particles landing on each other in every column: long chains of rounds; digest + list against the oracle"""
    size = (1040, 544)
    cells, zone = make_world(size, [(M.TEST_SAND, 35), (M.WATER, 15), (M.COBBLE_STONE, 3)], seed=5)
    g = gpu_world(cells, zone, particles=True, particle_capacity=1 << 21)
    o = oracle_world(cells, zone, particles=True)
    for t in range(10):
        g.tick(t)
        o.tick(t)
        g.tick_particles()
        o.tick_particles()
        a, b = g.particle_list(), o.particle_list()
        assert len(a) == len(b) and a.tobytes() == b.tobytes(), f"tick {t}"
        from oracle import oraclesim
        assert g.hash() == oraclesim.hash_cells(o.cells()), f"tick {t}"
    g.close()
    o.close()


def test_device_particles_error_paths():
    from fallingsandsurvival_b200 import world
    cells, zone = make_world((264, 336), "c2", seed=1)
    parts = world.StripWorld(cells, zone, 2, defs=full_table(), particles=True)
    with pytest.raises(world.FssError):
        parts.parts[0].tick_particles()
    parts.close()
    g = gpu_world(cells, zone, particles=True)
    bad = np.zeros(1, dtype=abi.LIVE_PARTICLE_DTYPE)
    bad["tile"]["material"] = 999
    with pytest.raises(world.FssError):
        g.add_particles(bad)
    g.tick_particles()  # empty list: nothing to do
    assert g.particles_count() == 0
    g.close()


# ---- World::dirty on the device and the dirty -> pixels job of Game::tick ------------------------------
def _looks():
    """Material::alpha / emitColor of the 41 materials (tests/golden/materials_looks.json, exported from the reference build)"""
    with open(os.path.join(GOLDEN, "materials_looks.json")) as f:
        d = json.load(f)
    return np.array(d["alpha"], dtype=np.uint8), np.array(d["emit_color"], dtype=np.uint32)


@pytest.mark.parametrize("early_out", [True, False])
@pytest.mark.parametrize("mix", ["c3", "sand", "liquids", "gas", "fire", "pockets"])
def test_device_dirty_and_pixels_match_oracle(mix, early_out):
    """every site of World::tick / World::tickParticles that sets dirty[] sets it on the device; the render job gives the
    same four RGBA layers, flags, per-material counts and consumes the flow accumulators the same way"""
    cells, zone = make_world((264, 240), mix, seed=4000 + len(mix))
    alpha, emit = _looks()
    g = gpu_world(cells, zone, particles=True, track_dirty=True, track_flow=True, early_out=early_out)
    o = oracle_world(cells, zone, particles=True)
    g.set_material_looks(alpha, emit)
    o.set_material_looks(alpha, emit)
    for t in range(24):
        g.tick(t)
        o.tick(t)
        g.tick_particles()
        o.tick_particles()
        if t == 5:  # a host edit: the game marks what it touched
            g.mark_dirty(30, 30, 50, 20)
            o.mark_dirty(30, 30, 50, 20)
        if t % 4 == 2:
            g.tick_temperature()
            o.tick_temperature()
        da, db = g.dirty(), o.dirty()
        assert da.tobytes() == db.tobytes(), f"tick {t}: dirty differs in {int((da != db).sum())} cells, first {np.argwhere(da != db)[0]}"
        if t % 3 == 1:
            mg, ig = g.render_dirty()
            mo, io = o.render_dirty()
            assert ig == io and np.array_equal(mg, mo), f"tick {t}: {ig} vs {io}"
            for layer in range(4):
                a, b = g.pixels(layer), o.pixels[layer]
                assert a.tobytes() == b.tobytes(), f"tick {t}: layer {layer} differs in {int((a != b).any(axis=2).sum())} pixels"
            assert not g.dirty().any()
            fa, fb = g.flow(), o.flow()
            assert fa[0].tobytes() == fb[0].tobytes() and fa[1].tobytes() == fb[1].tobytes()
        a, b = g.cells(), o.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    g.close()
    o.close()


@pytest.mark.parametrize("n_strips,early_out", [(2, True), (3, False), (3, True)])
def test_dirty_and_pixels_on_strips_equal_whole_world(n_strips, early_out):
    """World::dirty on a strip-sharded world: the marks a strip leaves in its ghost rows travel with the halo exchange and are
    OR-ed into the owner's plane; every strip renders the rows it owns.  Marks every tick, the four layers, the flags, the
    per-material counts (summed over the strips) and the consumed flow accumulators against the undivided world."""
    from fallingsandsurvival_b200 import world
    cells, zone = make_world((264, 400), "c3", seed=21)
    alpha, emit = _looks()
    whole = gpu_world(cells, zone, particles=True, track_dirty=True, track_flow=True, early_out=early_out)
    parts = world.StripWorld(cells, zone, n_strips, defs=full_table(), particles=True, track_dirty=True, track_flow=True, early_out=early_out)
    whole.set_material_looks(alpha, emit)
    parts.set_material_looks(alpha, emit)
    for t in range(18):
        whole.tick(t)
        parts.tick(t)
        if t == 4:  # a host edit across a strip boundary
            whole.mark_dirty(20, 100, 60, 150)
            parts.mark_dirty(20, 100, 60, 150)
        if t % 4 == 2:
            whole.tick_temperature()
            parts.tick_temperature()
        da, db = whole.dirty(), parts.dirty()
        assert da.tobytes() == db.tobytes(), f"tick {t}: dirty differs in {int((da != db).sum())} cells, first {np.argwhere(da != db)[0]}"
        if t % 3 == 1:
            ma, ia = whole.render_dirty()
            mb, ib = parts.render_dirty()
            assert ia == ib and np.array_equal(ma, mb), f"tick {t}: {ia} vs {ib}"
            for layer in range(4):
                a, b = whole.pixels(layer), parts.pixels(layer)
                assert a.tobytes() == b.tobytes(), f"tick {t}: layer {layer} differs in {int((a != b).any(axis=2).sum())} pixels"
            assert not parts.dirty().any()
            fa, fb = whole.flow(), parts.flow()
            assert fa[0].tobytes() == fb[0].tobytes() and fa[1].tobytes() == fb[1].tobytes()
        a, b = whole.cells(), parts.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(a, b)
    whole.close()
    parts.close()


def test_device_dirty_error_paths():
    from fallingsandsurvival_b200 import world
    cells, zone = make_world((264, 240), "c2", seed=1)
    g = gpu_world(cells, zone)
    with pytest.raises(world.FssError):
        g.render_dirty()
    with pytest.raises(world.FssError):
        g.mark_dirty(0, 0, 8, 8)
    g.close()
    s = world.World(cells, zone, defs=full_table(), strip=(0, 144), track_dirty=True)  # a strip renders the rows it owns
    with pytest.raises(world.FssError):
        s.tick_particles()  # device particles stay refused on strips
    s.mark_dirty(0, 100, 10, 60)  # the owned part of the rectangle
    assert int(s.dirty(0, 0, 264, 160).sum()) == 10 * 44
    s.render_dirty()
    with pytest.raises(world.FssError) as e:
        s.pixels(0, 0, 130, 10, 20)  # reaches into the neighbour's rows
    assert e.value.code == -6
    s.close()
    g = gpu_world(cells, zone, track_dirty=True)
    with pytest.raises(world.FssError):
        g.pixels(0)  # before the first render
    g.mark_dirty(-5, -5, 20, 20)  # clipped like the reference's own bounds checks
    assert int(g.dirty().sum()) == 15 * 15
    moving, info = g.render_dirty()
    assert info[0] == 1 and int(moving.sum()) == 15 * 15
    assert g.pixels(0, 3, 3, 8, 8).shape == (8, 8, 4)
    g.close()


def test_device_particles_short_list_path():
    """lists of at most 1024 particles run their rounds inside one CTA (k_particles_small): the hand-made particles that
    drive every branch, on a world that spawns none of its own, so the list stays short from the first tick on"""
    from helpers import particle_scene
    cells, zone, parts = particle_scene(seed=5)
    parts = parts[:160]  # the hand-made ones, without the random crowd
    g = gpu_world(cells, zone, particles=False)
    o = oracle_world(cells, zone, particles=False)
    g.add_particles(parts)
    o.add_particles(parts)
    for t in range(30):
        g.tick(t)
        o.tick(t)
        g.tick_particles()
        o.tick_particles()
        _same_particles(t, g, o)
        if t == 10:  # a second wave lands on what the first one built
            g.add_particles(parts[:96])
            o.add_particles(parts[:96])
    g.close()
    o.close()


@pytest.mark.parametrize("seed,n", [(1, 200), (2, 1500), (3, 6000)])
def test_device_particles_fuzz(seed, n):
    """random particles of every kind thrown into a busy world: fast ones (paths longer than the 64 cells that are listed
    one by one), target forces, `phase`, temporary ones, liquids of three materials, start cells inside solids and
    outside the zone; short-list path (n = 200), multi-kernel path, and the hand-over between them as the list shrinks"""
    from helpers import random_particles
    size = (392, 304)
    cells, zone = make_world(size, "c3", seed=50 + seed)
    parts = random_particles(seed, n, size)
    g = gpu_world(cells, zone, particles=True, track_dirty=True)
    o = oracle_world(cells, zone, particles=True)
    g.add_particles(parts)
    o.add_particles(parts)
    for t in range(14):
        g.tick(t)
        o.tick(t)
        g.tick_particles()
        o.tick_particles()
        _same_particles(t, g, o)
        assert g.dirty().tobytes() == o.dirty().tobytes(), f"tick {t}: World::dirty differs"
        if t == 6:  # a second batch while the first is still landing
            g.add_particles(parts[: n // 2])
            o.add_particles(parts[: n // 2])
    g.close()
    o.close()


def test_thin_liquids_match_oracle():
    """amounts around FLUID_MinValue (helpers.thin_liquid_world): the early exits of the SOUP rule that full cells never reach"""
    from helpers import thin_liquid_world
    cells, zone = thin_liquid_world()
    g = gpu_world(cells, zone, particles=True, track_dirty=True, track_flow=True)
    o = oracle_world(cells, zone, particles=True)
    for t in range(30):
        g.tick(t)
        o.tick(t)
        a, b = g.cells(), o.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(b, a)
        assert g.dirty().tobytes() == o.dirty().tobytes(), f"tick {t}: dirty differs"
    fa, fb = g.flow(), o.flow()
    assert fa[0].tobytes() == fb[0].tobytes() and fa[1].tobytes() == fb[1].tobytes()
    g.close()
    o.close()


def test_fat_drops_match_oracle():
    """several spawn records per cell visit (helpers.fat_drops_world): their order, the velocity draws 2j / 2j+1, landing"""
    from helpers import fat_drops_world
    cells, zone = fat_drops_world()
    g = gpu_world(cells, zone, particles=True)
    o = oracle_world(cells, zone, particles=True)
    for t in range(16):
        g.tick(t)
        o.tick(t)
        g.tick_particles()
        o.tick_particles()
        _same_particles(t, g, o)
    g.close()
    o.close()


def test_sand_reaction_below_a_temperature():
    """REACT_TEMPERATURE_BELOW on a SAND material (world.cpp:1255-1263): the shipped table only has the ABOVE kind on one"""
    defs = full_table()
    import copy
    defs = [copy.copy(d) for d in defs]
    defs[M.TEST_SAND].reactions = [(abi.REACT_TEMPERATURE_BELOW, -200, M.DIRT), (abi.REACT_TEMPERATURE_ABOVE, 300, M.GOLD_MOLTEN)]
    cells, zone = make_world((264, 208), "c3", seed=12, defs=defs)
    rng = np.random.default_rng(2)
    sand = cells["material"] == M.TEST_SAND
    cells["temperature"][sand] = rng.integers(-400, 500, cells.shape)[sand]
    from fallingsandsurvival_b200 import world
    from oracle import oraclesim
    g = world.World(cells, zone, defs=defs, particles=True)
    o = oraclesim.OracleWorld(cells, zone, defs=defs, particles=True)
    for t in range(8):
        g.tick(t)
        o.tick(t)
        if t % 4 == 2:
            g.tick_temperature()
            o.tick_temperature()
        a, b = g.cells(), o.cells()
        assert a.tobytes() == b.tobytes(), f"tick {t}: " + diff_report(b, a)
    assert (b["material"] == M.DIRT).sum() > (cells["material"] == M.DIRT).sum()
    g.close()
    o.close()


def test_full_spawn_buffer_loses_no_matter():
    """a spawn site that finds the record buffer (fss_config.particle_capacity) full must leave its cell in
    the grid instead of dropping the record of a cell that already left it.  SAND cells in the grid + SAND records handed
    out stays constant; the refusals are counted; a drain that does not fit consumes nothing."""
    import ctypes as C
    cells, zone = make_world((264, 208), "sand", seed=31)
    g = gpu_world(cells, zone, particles=True, particle_capacity=64)
    sand_ids = [d.id for d in full_table() if d.physics_type == abi.PHYS_SAND]
    before = int(np.isin(cells["material"], sand_ids).sum())
    handed_out = 0
    for t in range(6):
        g.tick(t)
        n, total = C.c_size_t(), C.c_size_t()
        small = np.zeros(4, dtype=abi.PARTICLE_DTYPE)
        rc = g.L.fss_drain_particles(g.h_, small.ctypes.data, 4, C.byref(n), C.byref(total))
        if total.value > 4:
            assert rc == -6 and n.value == 0  # nothing consumed
        recs = g.drain_particles()
        assert len(recs) <= 64 and (total.value <= 4 or len(recs) == total.value)
        handed_out += int(np.isin(recs["cell"]["material"], sand_ids).sum())
        now = int(np.isin(g.cells()["material"], sand_ids).sum())
        assert now + handed_out == before, f"tick {t}: {before - now - handed_out} SAND cells vanished"
    assert g.stats().spawns_refused > 0
    g.close()


def test_mark_dirty_list_equals_single_marks():
    cells, zone = make_world((264, 208), "c3", seed=32)
    rects = np.array([[20, 20, 5, 1], [25, 20, 9, 1], [40, 60, 3, 7], [-4, 100, 10, 10], [200, 190, 100, 100], [50, 50, 0, 4]], dtype=np.int32)
    a = gpu_world(cells, zone, track_dirty=True)
    b = gpu_world(cells, zone, track_dirty=True)
    for x, y, w, h in rects:
        a.mark_dirty(int(x), int(y), int(w), int(h))
    b.mark_dirty_list(rects)
    assert a.dirty().tobytes() == b.dirty().tobytes() and int(a.dirty().sum()) > 0
    a.close()
    b.close()


@pytest.mark.parametrize("mix", ["c2", "c3"])
def test_tick_host_pipeline_equals_upload_tick_download(mix, monkeypatch):
    """fss_tick_host (upload, the 24 sub-steps and download overlapped band by band, each band's lower edge receding by one
    chunk pair per sub-step) == fss_upload_rect + fss_tick + fss_download_rect == the oracle, with host edits in between"""
    import torch
    monkeypatch.setenv("FSS_HOST_BAND_ROWS", "832")  # the smallest band: 26 chunk pairs (read once per process: set before the first call)
    cells, zone = make_world((264, 16 + 3 * 832 + 32 + 16), mix, seed=77)
    a = gpu_world(cells, zone)
    b = gpu_world(cells, zone)
    o = oracle_world(cells, zone)
    host = torch.empty((cells.shape[0], cells.shape[1], 20), dtype=torch.uint8, pin_memory=True).numpy().view(abi.CELL_DTYPE).reshape(cells.shape)
    host[...] = cells
    plain = cells.copy()
    for t in range(4):
        if t == 2:  # a host edit between ticks: a block of sand dropped across a band boundary
            for buf in (host, plain):
                buf[820:860, 100:140] = buf[40, 40]
                buf["material"][820:860, 100:140] = M.TEST_SAND
            o.view()[820:860, 100:140] = host[820:860, 100:140]
        a.tick_host(host, t)
        b.upload(0, 0, plain)
        b.tick(t)
        plain = b.cells()
        o.tick(t)
        assert host.tobytes() == plain.tobytes(), f"tick {t}: pipelined != plain: " + diff_report(plain, host)
        assert host.tobytes() == o.cells().tobytes(), f"tick {t}: " + diff_report(o.cells(), host)
        assert a.cells().tobytes() == host.tobytes()
    for w in (a, b, o):
        w.close()


def test_worklist_launches_equal_plain_launches():
    """mostly settled worlds go through a compacted list of the tiles that still hold a not-settled chunk (k_worklist); the
    same ticks with the list forced on for every launch, with plain launches only, and with the early-out off must agree --
    on the sparse world and on a dense one, through fss_tick and through the banded fss_tick_host"""
    import subprocess, sys
    code = r"""
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch
from helpers import full_table
from fallingsandsurvival_b200 import world, worlds, abi, materials as M
size = (1056, 16 + 2 * 832 + 32 + 16)
zone = worlds.default_zone(*size)
out = []
for sparse in (True, False):
    g = world.World(None, zone, defs=full_table(), size=size, early_out=os.environ.get("EO", "1") == "1")
    kw = dict(solid_from_row=int(size[1] * 0.3), n_boxes=3, box_size=128, sparse_material=M.TEST_SAND, sparse_per_mille=2) if sparse else {}
    g.fill_synthetic(worlds.MIX_C2, seed=3, **kw)
    for t in range(40):
        g.tick(t)
    host = torch.empty((size[1], size[0], 20), dtype=torch.uint8, pin_memory=True).numpy().view(abi.CELL_DTYPE).reshape(size[1], size[0])
    host[...] = g.cells()
    for t in range(40, 44):
        g.tick_host(host, t)
    out.append("%%016x" %% g.hash())
    g.close()
print(" ".join(out))
""" % ((os.path.dirname(os.path.dirname(os.path.abspath(__file__))),) * 2)
    res = {}
    for name, env in (("list", {"FSS_WORKLIST_ALWAYS": "1"}), ("plain", {"FSS_NO_WORKLIST": "1"}), ("auto", {}), ("no early-out", {"EO": "0"})):
        e = dict(os.environ, FSS_HOST_BAND_ROWS="832", **env)
        res[name] = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, check=True).stdout.strip()
    assert len(set(res.values())) == 1, res


def test_chunk_file_from_device_planes_and_back():
    """Chunk::write / Chunk::read through the ABI: device planes -> 12-byte records (fss_read_chunk_records) -> the reference's
    file (fss_chunk_file_encode: LZ4) -> records (fss_chunk_file_decode) -> another world's planes (fss_merge_chunk_records)"""
    from fallingsandsurvival_b200 import world
    cells, zone = make_world((264, 208), "c3", seed=41)
    a = gpu_world(cells, zone)
    for t in range(3):
        a.tick(t)
    recs = a.read_chunk_records(40, 30, 128, 128)
    layer2 = np.zeros((128, 128), dtype=abi.RECORD_DTYPE)
    bg = np.arange(128 * 128, dtype=np.uint32).reshape(128, 128)
    data = world.chunk_file_encode(2, recs, layer2, bg)
    phase, r2, l2, b2 = world.chunk_file_decode(data)
    assert phase == 2 and np.array_equal(b2, bg)
    b = gpu_world(make_world((264, 208), "sand", seed=42)[0], zone)
    b.merge_chunk_records(40, 30, r2)
    ca, cb = a.cells()[30:158, 40:168], b.cells()[30:158, 40:168]
    for f in ("material", "color", "temperature"):
        assert np.array_equal(ca[f], cb[f]), f
    assert (cb["fluid_amount"] == 2.0).all() and (cb["moved"] == 0).all()  # the fields the file does not store (MaterialInstance.hpp:21)
    a.close()
    b.close()
