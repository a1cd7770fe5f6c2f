"""The per-particle functions of the product's particle kernels (fallingsandsurvival_b200/csrc/fss_particles.cuh: speculate,
cascade with its potential sets, commit, the spawn-record conversion) compiled for the host by tests/hostsim and driven like
fss_tick_particles drives its multi-kernel path; compared with the oracle's serial World::tickParticles every tick, without a
GPU, with the threads of every phase in three different orders and with hash tables small enough to collide constantly.
World::tick itself is done by a second oracle world here (the rules have their own host test, test_hostsim_rules.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import diff_report, full_table, make_world, particle_scene, random_particles
from fallingsandsurvival_b200 import abi
from fallingsandsurvival_b200 import materials as M
from oracle import oraclesim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hp(tmp_path_factory):
    so = tmp_path_factory.mktemp("hostsim_p") / "libhostsim_p.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "tests", "hostsim", "shim"), "-I", os.path.join(ROOT, "fallingsandsurvival_b200", "csrc"),
                           "-o", str(so), os.path.join(ROOT, "tests", "hostsim", "hostsim_particles.cpp")])
    L = C.CDLL(str(so))
    L.hp_create.restype = C.c_void_p
    L.hp_create.argtypes = [C.c_int] * 7
    L.hp_destroy.argtypes = [C.c_void_p]
    L.hp_set_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.hp_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.hp_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hp_add.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.hp_adopt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.hp_tick.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hp_count.restype = C.c_size_t
    L.hp_count.argtypes = [C.c_void_p]
    L.hp_list.argtypes = [C.c_void_p, C.c_void_p]
    L.hp_rounds.restype = C.c_uint64
    L.hp_rounds.argtypes = [C.c_void_p]
    return L


def _run(L, cells, zone, parts, ticks, slots, order, with_tick=True, second_wave_at=None):
    defs = full_table()
    h, w = cells.shape
    o = oraclesim.OracleWorld(cells, zone, defs=defs, particles=True)    # the serial reference: tick + tickParticles
    eng = oraclesim.OracleWorld(cells, zone, defs=defs, particles=True)  # World::tick for the host build's grid
    hw = L.hp_create(w, h, *zone, 0)
    tbl = M.table_struct(defs)
    L.hp_set_materials(hw, C.byref(tbl), len(defs))
    cur = np.ascontiguousarray(cells, dtype=abi.CELL_DTYPE)
    L.hp_upload(hw, cur.ctypes.data)
    if parts is not None and len(parts):
        parts = np.ascontiguousarray(parts, dtype=abi.LIVE_PARTICLE_DTYPE)
        o.add_particles(parts)
        L.hp_add(hw, parts.ctypes.data, len(parts))
    most_rounds = 0
    for t in range(ticks):
        if with_tick:
            o.tick(t)
            eng.view()[...] = cur
            eng.tick(t)
            recs = eng.drain_particles()
            cur = eng.cells()
            L.hp_upload(hw, cur.ctypes.data)
            if len(recs):
                recs = np.ascontiguousarray(recs)
                L.hp_adopt(hw, recs.ctypes.data, len(recs))
        o.tick_particles()
        assert L.hp_tick(hw, slots, order) == 0, f"tick {t}: a round retired nobody"
        most_rounds = max(most_rounds, L.hp_rounds(hw))
        n = L.hp_count(hw)
        got = np.zeros(max(n, 1), dtype=abi.LIVE_PARTICLE_DTYPE)
        L.hp_list(hw, got.ctypes.data)
        got = got[:n]
        want = o.particle_list()
        assert len(got) == len(want), f"tick {t}: {len(got)} particles, the serial loop has {len(want)}"
        assert got.tobytes() == want.tobytes(), f"tick {t}: particle lists differ"
        out = np.zeros((h, w), dtype=abi.CELL_DTYPE)
        L.hp_download(hw, out.ctypes.data, None)
        assert out.tobytes() == o.cells().tobytes(), f"tick {t}: " + diff_report(o.cells(), out)
        cur = out
        if second_wave_at == t:
            o.add_particles(parts[: len(parts) // 2])
            L.hp_add(hw, parts[: len(parts) // 2].ctypes.data, len(parts) // 2)
    L.hp_destroy(hw)
    o.close()
    eng.close()
    return most_rounds


@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("slots", [16, 1])
def test_particle_kernels_scene(hp, slots, order):
    """the hand-made scene that drives every branch of World::tickParticles, with World::tick feeding the list"""
    cells, zone, parts = particle_scene()
    rounds = _run(hp, cells, zone, parts, ticks=8, slots=slots, order=order)
    assert rounds > 1  # particles did depend on each other


@pytest.mark.parametrize("order", [0, 2])
@pytest.mark.parametrize("seed,n", [(1, 200), (2, 1500)])
def test_particle_kernels_fuzz(hp, seed, n, order):
    size = (392, 304)
    cells, zone = make_world(size, "c3", seed=50 + seed)
    _run(hp, cells, zone, random_particles(seed, n, size), ticks=10, slots=4, order=order, second_wave_at=5)


def test_particle_kernels_dense_rain(hp):
    """everything in the air at once: long chains of rounds"""
    cells, zone = make_world((264, 240), [(M.TEST_SAND, 35), (M.WATER, 15), (M.COBBLE_STONE, 3)], seed=5)
    rounds = _run(hp, cells, zone, None, ticks=5, slots=16, order=2)
    assert rounds > 5


def test_two_phase_commit_experiment(hp, monkeypatch):
    """a round-count experiment that only exists in the host harness (HP_TWO_PHASE, tests/hostsim/hostsim_particles.cpp): a sure
    particle does not wait for a lower particle that merely looked at its target if that one acts in the same round.  Must stay
    exact to be worth building into the kernels (profiles/r01_tuning_log.md)."""
    monkeypatch.setenv("HP_TWO_PHASE", "1")
    cells, zone, parts = particle_scene()
    _run(hp, cells, zone, parts, ticks=8, slots=16, order=2)
    size = (392, 304)
    cells, zone = make_world(size, "c3", seed=52)
    _run(hp, cells, zone, random_particles(2, 1500, size), ticks=8, slots=1, order=0, second_wave_at=4)
