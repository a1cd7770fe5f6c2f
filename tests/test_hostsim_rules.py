"""The product's movement RULES (fallingsandsurvival_b200/csrc/fss_rules.cuh: the code of kernel K1, all three specialisations,
with and without the flow and dirty variants) compiled for the host by tests/hostsim and compared with the oracle, every tick,
without a GPU.  What this does not cover is stated in tests/hostsim/hostsim.cpp (the kernel wrapper and the ABI's table
conversion, restated there); the GPU parity tests cover the real thing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import MIXES, diff_report, full_table, make_world
from fallingsandsurvival_b200 import abi
from fallingsandsurvival_b200 import materials as M
from oracle import oraclesim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GENERAL, NOGAS, LIQUID = 0, 1, 2


def _build(tmp_path_factory, name, defines=()):
    so = tmp_path_factory.mktemp(name) / "libhostsim.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", *defines,
                           "-I", os.path.join(ROOT, "tests", "hostsim", "shim"), "-I", os.path.join(ROOT, "fallingsandsurvival_b200", "csrc"),
                           "-o", str(so), os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    L = C.CDLL(str(so))
    L.hs_create.restype = C.c_void_p
    L.hs_create.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int]
    L.hs_destroy.argtypes = [C.c_void_p]
    L.hs_set_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.hs_set_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.hs_set_zone.argtypes = [C.c_void_p] + [C.c_int] * 5
    L.hs_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.hs_download.argtypes = [C.c_void_p] * 5
    L.hs_substep.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.hs_drain.restype = C.c_size_t
    L.hs_drain.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    return L


@pytest.fixture(scope="module")
def hostsim(tmp_path_factory):
    return _build(tmp_path_factory, "hostsim")


class HostSim:
    def __init__(self, L, cells, zone, defs, flags, early_out=False):
        self.L = L
        self.h, self.w = cells.shape
        self.zone = zone
        self.h_ = L.hs_create(self.w, self.h, 0xF55, flags, M.WATER)
        tbl = M.table_struct(defs)
        assert L.hs_set_materials(self.h_, C.byref(tbl), len(defs)) == 0
        for i in range(M.N_TEXTURES):
            t = np.ascontiguousarray(M.synthetic_texture(i), dtype=np.uint32)
            L.hs_set_texture(self.h_, i, t.ctypes.data, t.shape[1], t.shape[0])
        L.hs_set_zone(self.h_, *zone, 1 if early_out else 0)
        cells = np.ascontiguousarray(cells, dtype=abi.CELL_DTYPE)
        L.hs_upload(self.h_, cells.ctypes.data)

    def tick(self, t, modes):
        for it in range(6):
            for tk in range(4):
                assert self.L.hs_substep(self.h_, t, it, tk, modes[it]) == 0

    def state(self):
        cells = np.zeros((self.h, self.w), dtype=abi.CELL_DTYPE)
        fx = np.zeros((self.h, self.w), dtype=np.float32)
        fy = np.zeros((self.h, self.w), dtype=np.float32)
        dirty = np.zeros((self.h, self.w), dtype=np.uint8)
        self.L.hs_download(self.h_, cells.ctypes.data, fx.ctypes.data, fy.ctypes.data, dirty.ctypes.data)
        return cells, fx, fy, dirty

    def drain(self):
        out = np.zeros(1 << 20, dtype=abi.PARTICLE_DTYPE)
        n = self.L.hs_drain(self.h_, out.ctypes.data, len(out))
        out = out[:n]
        return out[abi.spawn_order(out, self.zone)]

    def close(self):
        self.L.hs_destroy(self.h_)


# which specialisation is exact for which iter (csrc/fss_abi.cu recompute_liq): liquid-only once no SAND / GAS / fire material can act
# any more, no-gas only if no GAS / fire material is present at all
SPECIALISED = {
    "c2": [NOGAS, NOGAS, LIQUID, LIQUID, LIQUID, LIQUID],
    "sand": [NOGAS, NOGAS, LIQUID, LIQUID, LIQUID, LIQUID],
    "liquids": [LIQUID] * 6,
    "pockets": [LIQUID] * 6,
}


@pytest.mark.parametrize("flags", [abi.PARTICLES_EMIT, abi.PARTICLES_EMIT | abi.TRACK_FLOW | abi.TRACK_DIRTY, 0])
@pytest.mark.parametrize("mix", sorted(MIXES))
def test_rule_code_matches_oracle_general_kernel(hostsim, mix, flags):
    cells, zone = make_world((200, 176), mix, seed=600 + len(mix))
    _compare(hostsim, cells, zone, [GENERAL] * 6, flags, ticks=10)


@pytest.mark.parametrize("flags", [abi.PARTICLES_EMIT, abi.PARTICLES_EMIT | abi.TRACK_FLOW | abi.TRACK_DIRTY])
@pytest.mark.parametrize("mix", sorted(SPECIALISED))
def test_rule_code_matches_oracle_specialised_kernels(hostsim, mix, flags):
    cells, zone = make_world((200, 176), mix, seed=700 + len(mix))
    _compare(hostsim, cells, zone, SPECIALISED[mix], flags, ticks=12)


@pytest.mark.parametrize("mix", sorted(SPECIALISED))
def test_register_window_liquid_path_matches_oracle(hostsim, mix):
    """csrc/fss_rules_liq.cuh (LiqWalk: the liquid-only sub-steps with the chunk rows in registers, scan 2 fused in) takes over
    from ChunkWalk<K1_LIQUID> where liq_window_applies(): iter >= 1, particles off, no flow / dirty planes"""
    cells, zone = make_world((200, 176), mix, seed=900 + len(mix))
    _compare(hostsim, cells, zone, SPECIALISED[mix], 0, ticks=16)
    _compare(hostsim, cells, zone, SPECIALISED[mix], 0, ticks=30, early_out=True)


@pytest.mark.parametrize("mix", ["c3", "gas", "fire", "randmats", "liquids"])
def test_register_window_general_path_matches_oracle(hostsim, mix):
    """WinWalk<true, *, GEN = true> (csrc/fss_rules_liq.cuh): GAS rising as an exchange with the row above, scans 2 and 3 by the
    visited bits, fire with its ignitions applied through memory (the "fire" mix sets several hundred cells alight in these
    ticks), SAND reactions on the payload temperature (hot gold ore melts on its first visit).  Taken wherever particles are off
    and no flow / dirty planes are kept; with and without the settled-chunk flags."""
    cells, zone = make_world((200, 176), mix, seed=1000 + len(mix))
    _compare(hostsim, cells, zone, [GENERAL] * 6, 0, ticks=24)
    _compare(hostsim, cells, zone, [GENERAL] * 6, 0, ticks=40, early_out=True)
    cells, zone = make_world((136, 240), mix, seed=77)
    _compare(hostsim, cells, zone, [GENERAL, GENERAL, NOGAS, LIQUID, LIQUID, LIQUID] if mix == "liquids" else [GENERAL] * 6, 0, ticks=12)


@pytest.mark.parametrize("table", ["reference", "dense sand", "sand turns to stone"])
def test_fire_prepass_and_its_conditions(hostsim, table):
    """fire_prepass_chunk + the window walk applying its verdicts (the reference's table), and the two conditions that switch the
    pre-pass off (csrc/fss_abi.cu `fire_static`, restated in tests/hostsim): a SAND material denser than fire displaces fire
    cells, a SAND reaction with a SOLID product changes which neighbours are SOLID in the middle of a sub-step"""
    import copy
    defs = copy.deepcopy(full_table())
    if table == "dense sand":
        defs[M.TEST_SAND].density = 30.0
    elif table == "sand turns to stone":
        defs[M.GOLD_ORE].reactions = [(abi.REACT_TEMPERATURE_ABOVE, 512, M.COBBLE_STONE)]
    mix = [(M.FIRE, 6), (M.SOFT_DIRT, 18), (M.COBBLE_STONE, 8), (M.TEST_SAND, 15), (M.GOLD_ORE, 5), (M.WATER, 12), (M.LAVA, 3), (M.STEAM, 3)]
    cells, zone = make_world((200, 176), mix, seed=91, defs=defs)
    cells["temperature"][cells["material"] == M.GOLD_ORE] = 600
    _compare(hostsim, cells, zone, [GENERAL] * 6, 0, ticks=16, defs=defs)
    _compare(hostsim, cells, zone, [GENERAL] * 6, 0, ticks=24, early_out=True, defs=defs)


def test_rule_code_with_early_out(hostsim):
    """the settled-chunk flags as the rules maintain them (ChunkWalk::run): skipping must not change anything"""
    cells, zone = make_world((264, 208), "pockets", seed=9)
    _compare(hostsim, cells, zone, SPECIALISED["pockets"], abi.PARTICLES_EMIT, ticks=24, early_out=True)
    cells, zone = make_world((264, 208), "settled", seed=9)
    _compare(hostsim, cells, zone, [GENERAL] * 6, abi.PARTICLES_EMIT, ticks=16, early_out=True)


def _compare(L, cells, zone, modes, flags, ticks, early_out=False, defs=None):
    defs = defs or full_table()
    hs = HostSim(L, cells, zone, defs, flags, early_out=early_out)
    o = oraclesim.OracleWorld(cells, zone, defs=defs, particles=bool(flags & abi.PARTICLES_EMIT))
    for t in range(ticks):
        hs.tick(t, modes)
        o.tick(t)
        got, fx, fy, dirty = hs.state()
        want = o.cells()
        assert got.tobytes() == want.tobytes(), f"tick {t}: " + diff_report(want, got)
        if flags & abi.TRACK_FLOW:
            ox, oy = o.flow()
            assert fx.tobytes() == ox.tobytes() and fy.tobytes() == oy.tobytes(), f"tick {t}: flow accumulators differ"
        if flags & abi.TRACK_DIRTY and not early_out:
            assert dirty.tobytes() == o.dirty().tobytes(), f"tick {t}: dirty differs"
        if flags & abi.PARTICLES_EMIT:
            assert hs.drain().tobytes() == o.drain_particles().tobytes(), f"tick {t}: spawn records differ"
    hs.close()
    o.close()
