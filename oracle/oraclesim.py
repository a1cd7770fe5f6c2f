"""ctypes wrapper around oracle/libfss_oracle.so (the plain-C restatement, fss_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; never from fallingsandsurvival_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from fallingsandsurvival_b200 import abi
from fallingsandsurvival_b200 import materials as M

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib_path():
    return os.path.join(HERE, "libfss_oracle.so")


def build():
    subprocess.check_call(["make", "-C", HERE, "libfss_oracle.so"], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(lib_path()):
        build()
    L = C.CDLL(lib_path())
    L.fsso_create.restype = C.c_void_p
    L.fsso_create.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int]
    L.fsso_destroy.argtypes = [C.c_void_p]
    L.fsso_set_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.fsso_set_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.fsso_set_tick_zone.argtypes = [C.c_void_p] + [C.c_int] * 4
    L.fsso_cells.restype = C.c_void_p
    L.fsso_cells.argtypes = [C.c_void_p]
    L.fsso_flow_x.restype = C.c_void_p
    L.fsso_flow_x.argtypes = [C.c_void_p]
    L.fsso_flow_y.restype = C.c_void_p
    L.fsso_flow_y.argtypes = [C.c_void_p]
    L.fsso_tick.argtypes = [C.c_void_p, C.c_uint32]
    L.fsso_substep_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fsso_tick_temperature.argtypes = [C.c_void_p]
    L.fsso_tick_temperature_rows.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.fsso_num_particles.restype = C.c_size_t
    L.fsso_num_particles.argtypes = [C.c_void_p]
    L.fsso_drain_particles.restype = C.c_size_t
    L.fsso_drain_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.fsso_dirty.restype = C.c_void_p
    L.fsso_dirty.argtypes = [C.c_void_p]
    L.fsso_set_material_looks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.fsso_render_dirty.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.fsso_particles_add.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.fsso_particles.restype = C.c_void_p
    L.fsso_particles.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.fsso_tick_particles.argtypes = [C.c_void_p]
    L.fsso_hash_rows.restype = C.c_uint64
    L.fsso_hash_rows.argtypes = [C.c_void_p] + [C.c_int] * 6
    L.fsso_histogram.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    _lib = L
    return L


def hash_cells(cells, x0=0, y0=0, w=None, h=None, global_y0=0):
    L = load()
    cells = np.ascontiguousarray(cells, dtype=abi.CELL_DTYPE)
    H, W = cells.shape
    w = W - x0 if w is None else w
    h = H - y0 if h is None else h
    return int(L.fsso_hash_rows(cells.ctypes.data, W, x0, y0, w, h, global_y0))


class OracleWorld:
    """Same surface as fallingsandsurvival_b200.World, computed by the C restatement on the host."""

    def __init__(self, cells, zone, defs=None, textures=None, seed=0xF55, particles=False, condense_material=M.WATER):
        self.L = load()
        h, w = cells.shape
        self.w, self.h = w, h
        self.defs = defs or M.named_materials()
        self.textures = textures if textures is not None else [M.synthetic_texture(i) for i in range(M.N_TEXTURES)]
        flags = abi.PARTICLES_EMIT if particles else 0
        self.h_ = self.L.fsso_create(w, h, seed, flags, condense_material)
        tbl = M.table_struct(self.defs)
        rc = self.L.fsso_set_materials(self.h_, C.byref(tbl), len(self.defs))
        assert rc == 0, rc
        for i, t in enumerate(self.textures):
            t = np.ascontiguousarray(t, dtype=np.uint32)
            self.L.fsso_set_texture(self.h_, i, t.ctypes.data, t.shape[1], t.shape[0])
        rc = self.L.fsso_set_tick_zone(self.h_, *zone)
        assert rc == 0, f"zone {zone} rejected: {rc}"
        self.zone = tuple(zone)
        self._view = np.ctypeslib.as_array(
            C.cast(self.L.fsso_cells(self.h_), C.POINTER(C.c_uint8)), shape=(h * w * abi.CELL_DTYPE.itemsize,)
        ).view(abi.CELL_DTYPE).reshape(h, w)
        self._view[...] = cells
        self.tick_ct = 0

    def cells(self):
        return self._view.copy()

    def flow(self):
        """World::flowX, flowY as (H, W) float32 copies"""
        out = []
        for f in (self.L.fsso_flow_x, self.L.fsso_flow_y):
            a = np.ctypeslib.as_array(C.cast(f(self.h_), C.POINTER(C.c_float)), shape=(self.h * self.w,))
            out.append(a.reshape(self.h, self.w).copy())
        return out[0], out[1]

    def view(self):
        return self._view

    def tick(self, tick_ct=None):
        if tick_ct is None:
            tick_ct = self.tick_ct
        self.L.fsso_tick(self.h_, tick_ct)
        self.tick_ct = tick_ct + 1

    def substep_rows(self, tick_ct, iter_, tk, cy_from=-(1 << 30), cy_to=1 << 30):
        self.L.fsso_substep_rows(self.h_, tick_ct, iter_, tk, cy_from, cy_to)

    def tick_temperature(self):
        self.L.fsso_tick_temperature(self.h_)

    def tick_temperature_rows(self, y_from, y_to):
        self.L.fsso_tick_temperature_rows(self.h_, y_from, y_to)

    def drain_particles(self):
        n = self.L.fsso_num_particles(self.h_)
        out = np.zeros(max(n, 1), dtype=abi.PARTICLE_DTYPE)
        m = self.L.fsso_drain_particles(self.h_, out.ctypes.data, n)
        return out[:m]

    # ---- World::dirty and the dirty -> pixels job of Game::tick (Game.cpp:2579-2650) ----
    def dirty(self):
        a = np.ctypeslib.as_array(C.cast(self.L.fsso_dirty(self.h_), C.POINTER(C.c_uint8)), shape=(self.h * self.w,))
        return a.reshape(self.h, self.w).copy()

    def mark_dirty(self, x, y, w, h):
        a = np.ctypeslib.as_array(C.cast(self.L.fsso_dirty(self.h_), C.POINTER(C.c_uint8)), shape=(self.h * self.w,)).reshape(self.h, self.w)
        a[max(y, 0):y + h, max(x, 0):x + w] = 1

    def set_material_looks(self, alpha, emit_color):
        alpha = np.ascontiguousarray(alpha, dtype=np.uint8)
        emit_color = np.ascontiguousarray(emit_color, dtype=np.uint32)
        assert self.L.fsso_set_material_looks(self.h_, alpha.ctypes.data, emit_color.ctypes.data, len(alpha)) == 0

    def render_dirty(self):
        """returns (moving_tiles, (had_dirty, had_fire, had_flow)); self.pixels[layer] is (H, W, 4) uint8 and persists"""
        if not hasattr(self, "pixels"):
            self.pixels = np.zeros((4, self.h, self.w, 4), dtype=np.uint8)
        ptrs = (C.c_void_p * 4)(*[self.pixels[k].ctypes.data for k in range(4)])
        moving = np.zeros(len(self.defs), dtype=np.uint64)
        info = np.zeros(3, dtype=np.uint32)
        self.L.fsso_render_dirty(self.h_, ptrs, moving.ctypes.data, len(self.defs), info.ctypes.data)
        return moving, tuple(int(v) for v in info)

    def tick_particles(self):
        """World::tickParticles() after adopting the spawn records of the ticks since the last call"""
        self.L.fsso_tick_particles(self.h_)

    def add_particles(self, parts):
        parts = np.ascontiguousarray(parts, dtype=abi.LIVE_PARTICLE_DTYPE)
        self.L.fsso_particles_add(self.h_, parts.ctypes.data, parts.size)

    def particle_list(self):
        n = C.c_size_t(0)
        p = self.L.fsso_particles(self.h_, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=abi.LIVE_PARTICLE_DTYPE)
        raw = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value * abi.LIVE_PARTICLE_DTYPE.itemsize,))
        return raw.view(abi.LIVE_PARTICLE_DTYPE).copy()

    def close(self):
        if self.h_:
            self._view = None
            self.L.fsso_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
