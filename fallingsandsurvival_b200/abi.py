"""ctypes / numpy mirrors of include/fss.h (the C ABI).  Layouts are asserted against the
library at load time (fss_abi_version) and by tests/test_abi.py against the header."""
import ctypes as C

import numpy as np

FSS_ABI_VERSION = 2

FSS_CHUNK_W = 4
FSS_CHUNK_H = 16
FSS_ZONE_ALIGN_X = 8
FSS_ZONE_ALIGN_Y = 32
FSS_ZONE_MARGIN = 16

PHYS_AIR, PHYS_SOLID, PHYS_SAND, PHYS_SOUP, PHYS_GAS, PHYS_PASSABLE = range(6)
REACT_TEMPERATURE_BELOW = 4
REACT_TEMPERATURE_ABOVE = 5
MAX_MATERIALS = 256
MAX_REACTIONS = 2
MAX_TEXTURES = 32
MAT_IS_FIRE = 1
MAT_IS_STEAM = 2
CREATE_FLAT, CREATE_RAND, CREATE_TEXTURE = 0, 1, 2
PARTICLES_EMIT = 1
NO_EARLY_OUT = 2
TRACK_FLOW = 4
TRACK_DIRTY = 8
SITE_TILES = 10000

ERRORS = {0: "FSS_OK", -1: "FSS_ERR_INVALID", -2: "FSS_ERR_ALIGN", -3: "FSS_ERR_CUDA", -4: "FSS_ERR_NO_DEVICE",
          -5: "FSS_ERR_STATE", -6: "FSS_ERR_RANGE", -7: "FSS_ERR_COMM"}

# fss_cell: 20 bytes, no padding (MaterialInstance.hpp:12-29 minus `id`)
CELL_DTYPE = np.dtype([
    ("material", "<u2"), ("moved", "u1"), ("settle_count", "u1"), ("color", "<u4"),
    ("temperature", "<i4"), ("fluid_amount", "<f4"), ("fluid_amount_diff", "<f4"),
])
assert CELL_DTYPE.itemsize == 20


class Reaction(C.Structure):
    _fields_ = [("type", C.c_int32), ("data1", C.c_int32), ("data2", C.c_int32)]


class Material(C.Structure):
    _fields_ = [
        ("physics_type", C.c_int32), ("iterations", C.c_int32), ("slipperyness", C.c_int32),
        ("density", C.c_float), ("add_temp", C.c_uint32), ("conduction_self", C.c_float),
        ("conduction_other", C.c_float), ("color", C.c_uint32), ("flags", C.c_uint32),
        ("n_reactions", C.c_int32), ("reactions", Reaction * MAX_REACTIONS),
        ("create_kind", C.c_int32), ("create_base", C.c_uint32), ("create_rand_mod", C.c_uint32),
        ("create_rand_shift", C.c_uint32), ("create_rand_site", C.c_uint32),
        ("create_texture", C.c_int32), ("create_temperature", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("seed", C.c_uint32),
        ("device", C.c_int32), ("flags", C.c_uint32), ("particle_capacity", C.c_uint32),
        ("strip_y0", C.c_uint32), ("strip_rows", C.c_uint32), ("condense_material", C.c_int32),
    ]


class CellStruct(C.Structure):
    _fields_ = [("material", C.c_uint16), ("moved", C.c_uint8), ("settle_count", C.c_uint8),
                ("color", C.c_uint32), ("temperature", C.c_int32), ("fluid_amount", C.c_float),
                ("fluid_amount_diff", C.c_float)]


assert C.sizeof(CellStruct) == 20


class ParticleSpawn(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("site", C.c_uint32), ("cell_key", C.c_uint32),
                ("tick_iter", C.c_uint32), ("cell", CellStruct)]


PARTICLE_DTYPE = np.dtype([("x", "<u4"), ("y", "<u4"), ("site", "<u4"), ("cell_key", "<u4"),
                           ("tick_iter", "<u4"), ("cell", CELL_DTYPE)])
assert PARTICLE_DTYPE.itemsize == C.sizeof(ParticleSpawn) == 40


def spawn_order(recs, zone):
    """argsort of spawn records into the order World::tick appends them to World::particles (fss_spawn_order_key, fss.h)"""
    zx, zy = int(zone[0]), int(zone[1])
    x = recs["x"].astype(np.uint64) - np.uint64(zx)
    y = recs["y"].astype(np.uint64) - np.uint64(zy)
    ci, cj, dx, dy = x // np.uint64(4), y // np.uint64(16), x % np.uint64(4), y % np.uint64(16)
    tk = (ci & np.uint64(1)) + np.uint64(2) * (np.uint64(1) - (cj & np.uint64(1)))
    key = ((recs["tick_iter"].astype(np.uint64) & np.uint64(0x3FFFFFFF)) << np.uint64(34)) | (tk << np.uint64(32)) | \
          ((ci & np.uint64(0x3FFF)) << np.uint64(18)) | ((cj & np.uint64(0xFFF)) << np.uint64(6)) | \
          ((np.uint64(15) - dy) << np.uint64(2)) | dx
    return np.argsort(key, kind="stable")


class Particle(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("vx", C.c_float), ("vy", C.c_float), ("ax", C.c_float),
                ("ay", C.c_float), ("target_x", C.c_float), ("target_y", C.c_float), ("target_force", C.c_float),
                ("lifetime", C.c_int32), ("fade_time", C.c_int32), ("in_object_state", C.c_uint16),
                ("phase", C.c_uint8), ("temporary", C.c_uint8), ("tile", CellStruct)]


# fss_particle == Particle (Particle.hpp:12-31) without the callback
LIVE_PARTICLE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("ax", "<f4"), ("ay", "<f4"),
                                ("target_x", "<f4"), ("target_y", "<f4"), ("target_force", "<f4"),
                                ("lifetime", "<i4"), ("fade_time", "<i4"), ("in_object_state", "<u2"),
                                ("phase", "u1"), ("temporary", "u1"), ("tile", CELL_DTYPE)])
assert LIVE_PARTICLE_DTYPE.itemsize == C.sizeof(Particle) == 68


class RenderInfo(C.Structure):
    _fields_ = [("had_dirty", C.c_uint32), ("had_fire", C.c_uint32), ("had_flow", C.c_uint32)]


class FillSpec(C.Structure):
    _fields_ = [
        ("seed", C.c_uint32), ("border", C.c_uint32), ("border_material", C.c_int32), ("n_entries", C.c_int32),
        ("material", C.c_int32 * 16), ("cumulative_percent", C.c_int32 * 16),
        ("solid_from_row", C.c_uint32), ("solid_material", C.c_int32),
        ("n_boxes", C.c_int32), ("box_size", C.c_uint32), ("sparse_material", C.c_int32),
        ("sparse_per_mille", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("substeps", C.c_uint64), ("storage_bytes", C.c_uint64),
                ("storage_pitch", C.c_uint32), ("storage_rows", C.c_uint32), ("storage_y0", C.c_uint32), ("reserved", C.c_uint32),
                ("spawns_refused", C.c_uint64)]


# fss_tile_record == MaterialInstanceData (Chunk.hpp:27-31): 12 bytes, 2 bytes of padding after `index`
RECORD_DTYPE = np.dtype({"names": ["index", "color", "temperature"], "formats": ["<u2", "<u4", "<i4"], "offsets": [0, 4, 8], "itemsize": 12})
