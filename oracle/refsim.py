"""ctypes wrapper around oracle/_ref/libfss_ref_<variant>.so -- the reference's OWN tick lines
compiled headless by oracle/build_ref.py.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs; never from fallingsandsurvival_b200/.
"""
import ctypes as C
import os

import numpy as np

from fallingsandsurvival_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def lib_path(variant="c4x16"):
    return os.path.join(HERE, "_ref", f"libfss_ref_{variant}.so")


def available(variant="c4x16"):
    return os.path.exists(lib_path(variant))


def load(variant="c4x16"):
    if variant in _libs:
        return _libs[variant]
    L = C.CDLL(lib_path(variant))
    L.fssref_init.argtypes = [C.c_uint]
    L.fssref_num_materials.restype = C.c_int
    L.fssref_get_material.argtypes = [C.c_int, C.POINTER(abi.Material), C.c_char_p, C.c_int]
    L.fssref_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.fssref_world_create.restype = C.c_void_p
    L.fssref_world_create.argtypes = [C.c_int] * 7
    L.fssref_world_destroy.argtypes = [C.c_void_p]
    L.fssref_set_cells.argtypes = [C.c_void_p, C.c_void_p]
    L.fssref_get_cells.argtypes = [C.c_void_p, C.c_void_p]
    L.fssref_get_flow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.fssref_tick.restype = C.c_double
    L.fssref_tick.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
    L.fssref_tick_temperature.restype = C.c_double
    L.fssref_tick_temperature.argtypes = [C.c_void_p]
    L.fssref_num_particles.restype = C.c_size_t
    L.fssref_num_particles.argtypes = [C.c_void_p]
    L.fssref_drain_particles.restype = C.c_size_t
    L.fssref_drain_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.fssref_get_dirty.argtypes = [C.c_void_p, C.c_void_p]
    L.fssref_mark_dirty.argtypes = [C.c_void_p] + [C.c_int] * 4
    L.fssref_get_material_looks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.fssref_render_dirty.restype = C.c_double
    L.fssref_render_dirty.argtypes = [C.c_void_p] * 7
    L.fssref_keep_particles.argtypes = [C.c_void_p, C.c_int]
    L.fssref_tick_particles.restype = C.c_double
    L.fssref_tick_particles.argtypes = [C.c_void_p]
    L.fssref_live_particles.restype = C.c_size_t
    L.fssref_live_particles.argtypes = [C.c_void_p]
    L.fssref_get_particles.restype = C.c_size_t
    L.fssref_get_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.fssref_add_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.fssref_texture_pixel.restype = C.c_uint32
    L.fssref_texture_pixel.argtypes = [C.c_int, C.c_int, C.c_int]
    L.fssref_init(1234)
    _libs[variant] = L
    return L


def materials(variant="c4x16"):
    """[(fss_material struct, name, interact)] exported by the reference's Materials::init() under srand(1234)."""
    L = load(variant)
    out = []
    for i in range(L.fssref_num_materials()):
        m = abi.Material()
        name = C.create_string_buffer(64)
        inter = L.fssref_get_material(i, C.byref(m), name, 64)
        out.append((m, name.value.decode(), inter))
    return out


def material_looks(variant="c4x16"):
    """(alpha, emitColor) per material id, Material.hpp:34, :38"""
    L = load(variant)
    n = L.fssref_num_materials()
    alpha = np.zeros(n, dtype=np.uint8)
    emit = np.zeros(n, dtype=np.uint32)
    L.fssref_get_material_looks(alpha.ctypes.data, emit.ctypes.data, n)
    return alpha, emit


def create(mat, x, y, seed, tick_ct, iter_, cx, cy, variant="c4x16"):
    L = load(variant)
    c = np.zeros((), dtype=abi.CELL_DTYPE)
    rc = L.fssref_create(mat, x, y, seed, tick_ct, iter_, cx, cy, c.ctypes.data)
    assert rc == 0
    return c


class RefWorld:
    """The reference's World (tiles, tickZone, tick, tickTemperature) behind the shim's C API."""

    def __init__(self, cells, zone, threads=None, variant="c4x16", seed=0xF55, particles=False):
        self.L = load(variant)
        h, w = cells.shape
        self.w, self.h = w, h
        self.zone = tuple(zone)
        self.seed = seed
        self.particles = particles
        threads = threads or os.cpu_count() or 1
        self.threads = threads
        self.h_ = self.L.fssref_world_create(w, h, *self.zone, threads)
        assert self.h_, "fssref_world_create failed"
        self.set_cells(cells)
        self.tick_ct = 0

    def set_cells(self, cells):
        cells = np.ascontiguousarray(cells, dtype=abi.CELL_DTYPE)
        assert cells.shape == (self.h, self.w)
        rc = self.L.fssref_set_cells(self.h_, cells.ctypes.data)
        assert rc == 0, "material id outside the reference's table"

    def cells(self):
        out = np.zeros((self.h, self.w), dtype=abi.CELL_DTYPE)
        self.L.fssref_get_cells(self.h_, out.ctypes.data)
        return out

    def flow(self):
        """World::flowX, flowY as (H, W) float32 arrays"""
        fx = np.zeros((self.h, self.w), dtype=np.float32)
        fy = np.zeros((self.h, self.w), dtype=np.float32)
        self.L.fssref_get_flow(self.h_, fx.ctypes.data, fy.ctypes.data)
        return fx, fy

    def tick(self, tick_ct=None):
        """one World::tick(); returns seconds inside it"""
        if tick_ct is None:
            tick_ct = self.tick_ct
        dt = self.L.fssref_tick(self.h_, self.seed, tick_ct, 1 if self.particles else 0)
        self.tick_ct = tick_ct + 1
        return dt

    def tick_temperature(self):
        return self.L.fssref_tick_temperature(self.h_)

    def drain_particles(self):
        n = self.L.fssref_num_particles(self.h_)
        xyv = np.zeros((max(n, 1), 4), dtype=np.float32)
        cells = np.zeros(max(n, 1), dtype=abi.CELL_DTYPE)
        temp = np.zeros(max(n, 1), dtype=np.int32)
        m = self.L.fssref_drain_particles(self.h_, xyv.ctypes.data, cells.ctypes.data, temp.ctypes.data, n)
        return xyv[:m], cells[:m], temp[:m]

    def dirty(self):
        out = np.zeros((self.h, self.w), dtype=np.uint8)
        self.L.fssref_get_dirty(self.h_, out.ctypes.data)
        return out

    def mark_dirty(self, x, y, w, h):
        self.L.fssref_mark_dirty(self.h_, x, y, w, h)

    def render_dirty(self):
        """the reference's own loop (Game.cpp:2584-2650); returns (moving_tiles as uint16, flags, seconds)"""
        if not hasattr(self, "pixels"):
            self.pixels = np.zeros((4, self.h, self.w, 4), dtype=np.uint8)
        n = self.L.fssref_num_materials()
        moving = np.zeros(n, dtype=np.uint16)
        info = np.zeros(3, dtype=np.uint32)
        dt = self.L.fssref_render_dirty(self.h_, self.pixels[0].ctypes.data, self.pixels[1].ctypes.data, self.pixels[2].ctypes.data,
                                        self.pixels[3].ctypes.data, moving.ctypes.data, info.ctypes.data)
        return moving, tuple(int(v) for v in info), dt

    def keep_particles(self, on=True):
        """leave World::particles in place after tick() so that tick_particles() can run on them"""
        self.L.fssref_keep_particles(self.h_, 1 if on else 0)

    def tick_particles(self):
        """one World::tickParticles(); returns seconds inside it"""
        return self.L.fssref_tick_particles(self.h_)

    def add_particles(self, parts):
        parts = np.ascontiguousarray(parts, dtype=abi.LIVE_PARTICLE_DTYPE)
        assert self.L.fssref_add_particles(self.h_, parts.ctypes.data, parts.size) == 0

    def particle_list(self):
        n = self.L.fssref_live_particles(self.h_)
        out = np.zeros(max(n, 1), dtype=abi.LIVE_PARTICLE_DTYPE)
        m = self.L.fssref_get_particles(self.h_, out.ctypes.data, n)
        return out[:m]

    def close(self):
        if self.h_:
            self.L.fssref_world_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
