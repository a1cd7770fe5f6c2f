"""Synthetic worlds as numpy arrays of fss_cell records.

Same definition as fss_fill_synthetic() in the CUDA library (csrc/fss_fill.cuh), so small worlds
can be built on the host for the parity tests and big ones directly on the device for benchmarks:
  h = fss_rng_cell_key(seed, x, y);  r = h % 100;  material = first entry with r < cumulative[i]
  cell = Tiles::create(material, x, y) with colour draws fss_rand(h, site, 0).
"""
from typing import List, Sequence, Tuple

import numpy as np

from . import abi
from . import materials as M

U32 = np.uint64(0xFFFFFFFF)


def _fmix32(h):
    h = h & U32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & U32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & U32
    h ^= h >> np.uint64(16)
    return h


def cell_key_grid(tkey: int, w: int, h: int, x0: int = 0, y0: int = 0):
    ys, xs = np.mgrid[y0:y0 + h, x0:x0 + w].astype(np.uint64)
    return _fmix32(np.uint64(tkey) ^ ((xs * np.uint64(0x9E3779B1)) & U32) ^ ((ys * np.uint64(0x85EBCA77)) & U32))


def _rand_grid(cellkey, site, k=0):
    return _fmix32((cellkey + np.uint64((site * 0x27D4EB2F) & 0xFFFFFFFF) + np.uint64((k * 0x165667B1) & 0xFFFFFFFF)) & U32) >> np.uint64(1)


def cells_of_material(defs: List[M.MaterialDef], textures, mat_grid, cellkey, x0=0, y0=0):
    """Vectorised Tiles::create(mat, x, y) for a grid of material ids."""
    h, w = mat_grid.shape
    out = np.zeros((h, w), dtype=abi.CELL_DTYPE)
    out["material"] = mat_grid
    out["fluid_amount"] = 2.0
    ys, xs = np.mgrid[y0:y0 + h, x0:x0 + w]
    for d in defs:
        sel = mat_grid == d.id
        if not sel.any():
            continue
        base = d.color if d.create_base is None else d.create_base
        if d.create_kind == abi.CREATE_FLAT:
            col = np.full(sel.sum(), base & 0xFFFFFFFF, dtype=np.uint64)
        elif d.create_kind == abi.CREATE_RAND:
            r = _rand_grid(cellkey[sel], d.create_rand_site, 0)
            col = np.uint64(base) + ((r % np.uint64(d.create_rand_mod)) << np.uint64(d.create_rand_shift))
        else:
            tex = textures[d.create_texture]
            th, tw = tex.shape
            col = tex[ys[sel] % th, xs[sel] % tw].astype(np.uint64)
        out["color"][sel] = (col & U32).astype(np.uint32)
        out["temperature"][sel] = d.create_temperature
    return out


def synthetic_world(width: int, height: int, mix: Sequence[Tuple[int, int]], seed: int = 0xF55,
                    border: int = 16, border_material: int = M.SOLID, defs=None, textures=None,
                    solid_from_row: int = 0, solid_material: int = M.COBBLE_STONE):
    """`mix` = [(material id, percent), ...]; the rest is AIR.  Cells closer than `border` to an edge are
    `border_material` (the reference's worlds keep >=128 untouched cells around tickZone, Game.cpp:2216)."""
    defs = defs or M.named_materials()
    textures = textures if textures is not None else [M.synthetic_texture(i) for i in range(M.N_TEXTURES)]
    key = cell_key_grid(seed, width, height)
    r = key % np.uint64(100)
    mat = np.zeros((height, width), dtype=np.uint16)
    cum = 0
    assigned = np.zeros((height, width), dtype=bool)
    for m, pct in mix:
        cum += pct
        sel = (~assigned) & (r < np.uint64(cum))
        mat[sel] = m
        assigned |= sel
    assert cum <= 100
    if solid_from_row:
        mat[solid_from_row:, :] = solid_material
    if border:
        mat[:border, :] = border_material
        mat[-border:, :] = border_material
        mat[:, :border] = border_material
        mat[:, -border:] = border_material
    return cells_of_material(defs, textures, mat, key)


def default_zone(width: int, height: int, border: int = 16):
    """tickZone = world minus `border`, shrunk to the schedule's alignment (origin multiple of 8 / any y,
    size multiple of 8 x 32)."""
    x0 = (border + abi.FSS_ZONE_ALIGN_X - 1) // abi.FSS_ZONE_ALIGN_X * abi.FSS_ZONE_ALIGN_X
    y0 = border
    zw = (width - border - x0) // abi.FSS_ZONE_ALIGN_X * abi.FSS_ZONE_ALIGN_X
    zh = (height - border - y0) // abi.FSS_ZONE_ALIGN_Y * abi.FSS_ZONE_ALIGN_Y
    return (x0, y0, zw, zh)


MIX_C2 = [(M.TEST_SAND, 25), (M.WATER, 20), (M.COBBLE_STONE, 10)]
MIX_C3 = MIX_C2 + [(M.LAVA, 3), (M.FIRE, 2), (M.GOLD_ORE, 3), (M.STEAM, 3), (M.SOFT_DIRT, 5)]


C1_SIZE = (1984, 1152)
C1_ZONE = (32, 32, 1920, 1088)


def material_test_world(defs, textures=None):
    """BASELINE.json configs[0] at full size: a 1984 x 1152 world, tick zone 1920 x 1088, laid out like
    MaterialTestGenerator::generateChunk (MaterialTestGenerator.cpp:32-48, CHUNK_W = CHUNK_H = 128): the cobblestone band
    at rows 401..450, four walled containers in chunks (1..4, 1) each filled with one random `Mat_N` that is SAND or SOUP,
    the two cobblestone ramps in chunk columns 1 and 4 -- plus painted 200 x 200 blobs of TEST_SAND, WATER and STEAM, a
    FIRE strip on a SOFT_DIRT block and a hot GOLD_ORE strip, so that every rule family runs.  Returns (cells, zone)."""
    w, h = C1_SIZE
    textures = textures if textures is not None else [M.synthetic_texture(i) for i in range(M.N_TEXTURES)]
    key = cell_key_grid(1, w, h)
    py, px = np.mgrid[0:h, 0:w]
    chx, chy, x, y = px // 128, py // 128, px % 128, py % 128
    mat = np.zeros((h, w), dtype=np.uint16)
    # the generator draws `rand() % MATERIALS.size()` per chunk until it hits a Mat_N of type SAND or SOUP: here the first four
    fillers = [d.id for d in defs if d.id >= 31 and d.physics_type in (abi.PHYS_SAND, abi.PHYS_SOUP)][:4]
    assert len(fillers) == 4, "the material table needs four Mat_N of type SAND / SOUP (helpers.full_table())"
    box = (chy == 1) & (chx >= 1) & (chx <= 4)
    wall = box & ((x < 8) | (y < 8) | (x >= 120) | ((y >= 120) & ((x < 60) | (x >= 68))))
    mat[wall] = M.COBBLE_DIRT
    for k, m in enumerate(fillers):
        mat[box & ~wall & (y > 96) & (chx == 1 + k)] = m
    ramp = (py <= 400) & (py > 300) & (((chx == 1) & (x < py - 300)) | ((chx == 4) & (128 - x < py - 300)))
    mat[ramp & ~box] = M.COBBLE_STONE
    mat[(py > 400) & (py <= 450)] = M.COBBLE_STONE
    mat[100:300, 700:900] = M.TEST_SAND
    mat[100:300, 1000:1200] = M.WATER
    mat[100:300, 1300:1500] = M.STEAM
    mat[250:300, 1600:1800] = M.SOFT_DIRT
    mat[246:250, 1610:1790] = M.FIRE
    mat[60:66, 700:900] = M.GOLD_ORE
    mat[600:800, 300:500] = M.WATER  # below the band: falls to the world floor
    mat[600:800, 900:1100] = M.TEST_SAND
    mat[:16, :] = mat[-16:, :] = M.SOLID
    mat[:, :16] = mat[:, -16:] = M.SOLID
    cells = cells_of_material(defs, textures, mat, key)
    cells["temperature"][mat == M.GOLD_ORE] = 700
    return cells, C1_ZONE
