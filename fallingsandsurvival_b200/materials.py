"""Host-side mirror of the reference's material table and tile factories.

`Materials::MATERIALS` (FallingSandSurvival/Materials.cpp:4-46 definitions, :50-199 init()) and the
colour recipes of `Tiles::create*` (FallingSandSurvival/Tiles.cpp:7-268) restated as data, in the
form fss_set_materials() (include/fss.h) takes.  tests/test_materials.py checks every row against
the table exported by the reference itself (oracle/_ref) and every recipe against Tiles::create().
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import abi

# texture indices used by the recipes below (order of the statics in Textures.hpp:11-29)
TEXTURE_INDEX = {
    "testTexture": 0, "dirt1Texture": 1, "stone1Texture": 2, "smoothStone": 3, "cobbleStone": 4,
    "flatCobbleStone": 5, "smoothDirt": 6, "cobbleDirt": 7, "flatCobbleDirt": 8, "softDirt": 9,
    "cloud": 10, "gold": 11, "goldMolten": 12, "goldSolid": 13, "iron": 14, "obsidian": 15,
}
N_TEXTURES = 16


@dataclass
class MaterialDef:
    id: int
    name: str
    physics_type: int
    slipperyness: int
    alpha: int
    density: float
    iterations: int
    color: int = 0xFFFFFFFF
    add_temp: int = 0
    conduction_self: float = 1.0
    conduction_other: float = 1.0
    flags: int = 0
    reactions: List[Tuple[int, int, int]] = field(default_factory=list)
    create_kind: int = abi.CREATE_FLAT
    create_base: Optional[int] = None  # None -> Material::color (Tiles.cpp:267 fall-through)
    create_rand_mod: int = 0
    create_rand_shift: int = 0
    create_rand_site: int = 0
    create_texture: int = -1
    create_temperature: int = 0

    def to_struct(self) -> abi.Material:
        m = abi.Material()
        m.physics_type = self.physics_type
        m.iterations = self.iterations
        m.slipperyness = self.slipperyness
        m.density = self.density
        m.add_temp = self.add_temp
        m.conduction_self = self.conduction_self
        m.conduction_other = self.conduction_other
        m.color = self.color & 0xFFFFFFFF
        m.flags = self.flags
        m.n_reactions = len(self.reactions)
        for i, (t, d1, d2) in enumerate(self.reactions[: abi.MAX_REACTIONS]):
            m.reactions[i].type, m.reactions[i].data1, m.reactions[i].data2 = t, d1, d2
        m.create_kind = self.create_kind
        m.create_base = (self.color if self.create_base is None else self.create_base) & 0xFFFFFFFF
        m.create_rand_mod = self.create_rand_mod
        m.create_rand_shift = self.create_rand_shift
        m.create_rand_site = self.create_rand_site
        m.create_texture = self.create_texture
        m.create_temperature = self.create_temperature
        return m


A, SO, SA, LQ, G, P = (abi.PHYS_AIR, abi.PHYS_SOLID, abi.PHYS_SAND, abi.PHYS_SOUP, abi.PHYS_GAS, abi.PHYS_PASSABLE)


def _tex(name):
    return dict(create_kind=abi.CREATE_TEXTURE, create_texture=TEXTURE_INDEX[name])


def _rand(base, mod, shift, tiles_line):
    return dict(create_kind=abi.CREATE_RAND, create_base=base, create_rand_mod=mod, create_rand_shift=shift,
                create_rand_site=abi.SITE_TILES + tiles_line)


def named_materials() -> List[MaterialDef]:
    """The 31 named materials, ids 0..30 (Materials.cpp:4-46); thermal constants :52-69;
    reactions :180-194; colour recipes Tiles.cpp (line cited per row)."""
    M = MaterialDef
    t = [
        M(0, "_AIR", A, 0, 255, 0.0, 0, conduction_self=0.8, conduction_other=0.8),
        M(1, "_SOLID", SO, 0, 255, 1.0, 0),
        M(2, "_SAND", SA, 20, 255, 10.0, 2),
        M(3, "_LIQUID", LQ, 0, 255, 1.5, 3),
        M(4, "_GAS", G, 0, 255, -1.0, 1),
        M(5, "_PASSABLE", P, 0, 255, 0.0, 0),
        M(6, "_OBJECT", P, 0, 255, 1000.0, 0),
        # Tiles.cpp:14-19 createTestSand: (220<<16) + ((155 + rand()%30)<<8) + 100
        M(7, "Test Sand", SA, 20, 255, 10.0, 2, **_rand((220 << 16) + (155 << 8) + 100, 30, 8, 16)),
        M(8, "Test Textured Sand", SA, 20, 255, 10.0, 2, **_tex("testTexture")),  # Tiles.cpp:21-29
        M(9, "Test Liquid", LQ, 0, 255, 1.5, 4, create_base=0x0000FF),  # Tiles.cpp:31-36
        M(10, "Stone", SO, 0, 255, 1.0, 0, **_tex("cobbleStone")),  # Tiles.cpp:38-50 (uses cobbleStone)
        # Tiles.cpp:52-57 createGrass: (40<<16) + ((120 + rand()%20)<<8) + 20
        M(11, "Grass", SA, 20, 255, 12.0, 1, **_rand((40 << 16) + (120 << 8) + 20, 20, 8, 54)),
        # Tiles.cpp:59-64 createDirt: ((60 + rand()%10)<<16) + (40<<8) + 20
        M(12, "Dirt", SA, 8, 255, 15.0, 1, **_rand((60 << 16) + (40 << 8) + 20, 10, 16, 60)),
        M(13, "Stone", SO, 0, 255, 1.0, 0, **_tex("smoothStone")),
        M(14, "Cobblestone", SO, 0, 255, 1.0, 0, conduction_self=0.01, conduction_other=0.4, **_tex("cobbleStone")),
        M(15, "Ground", SO, 0, 255, 1.0, 0, **_tex("smoothDirt")),
        M(16, "Hard Ground", SO, 0, 255, 1.0, 0, **_tex("cobbleDirt")),
        M(17, "Dirt", SO, 0, 255, 15.0, 2, **_tex("softDirt")),
        M(18, "Water", LQ, 0, 0x80, 1.5, 6, create_base=0x00B69F, create_temperature=-1023,
          reactions=[(abi.REACT_TEMPERATURE_ABOVE, 128, 26)]),  # Tiles.cpp:121-125
        M(19, "Lava", LQ, 0, 0xC0, 2.0, 1, add_temp=2, conduction_self=0.5, conduction_other=0.7,
          create_base=0xFF7C00, create_temperature=1024,
          reactions=[(abi.REACT_TEMPERATURE_BELOW, 512, 25)]),  # Tiles.cpp:127-131
        M(20, "Cloud", SO, 0, 127, 1.0, 0, **_tex("cloud")),
        M(21, "Gold Ore", SA, 20, 255, 20.0, 2, reactions=[(abi.REACT_TEMPERATURE_ABOVE, 512, 22)], **_tex("gold")),
        M(22, "Molten Gold", LQ, 0, 255, 20.0, 2, reactions=[(abi.REACT_TEMPERATURE_BELOW, 128, 23)],
          **_tex("goldMolten")),
        M(23, "Solid Gold", SO, 0, 255, 20.0, 2, **_tex("goldSolid")),
        M(24, "Iron Ore", SA, 20, 255, 20.0, 2, **_tex("iron")),
        M(25, "Obsidian", SO, 0, 255, 1.0, 0, **_tex("obsidian")),
        M(26, "Steam", G, 0, 255, -1.0, 1, flags=abi.MAT_IS_STEAM, create_base=0x666666),  # Tiles.cpp:177-179
        M(27, "Soft Dirt Sand", SA, 8, 255, 15.0, 2),
        # Tiles.cpp:181-188 createFire: (255<<16) + ((100 + rand()%50)<<8) + 50
        M(28, "Fire", P, 0, 255, 20.0, 1, flags=abi.MAT_IS_FIRE, **_rand((255 << 16) + (100 << 8) + 50, 50, 8, 184)),
        M(29, "Flat Cobblestone", SO, 0, 255, 1.0, 0, **_tex("flatCobbleStone")),
        M(30, "Flat Hard Ground", SO, 0, 255, 1.0, 0, **_tex("flatCobbleDirt")),
    ]
    return t


# ids by reference name
AIR, SOLID, SAND, LIQUID, GAS, PASSABLE, OBJECT, TEST_SAND, TEST_TEXTURED_SAND, TEST_LIQUID, STONE, GRASS, DIRT, \
    SMOOTH_STONE, COBBLE_STONE, SMOOTH_DIRT, COBBLE_DIRT, SOFT_DIRT, WATER, LAVA, CLOUD, GOLD_ORE, GOLD_MOLTEN, \
    GOLD_SOLID, IRON_ORE, OBSIDIAN, STEAM, SOFT_DIRT_SAND, FIRE, FLAT_COBBLE_STONE, FLAT_COBBLE_DIRT = range(31)


def random_materials(rng: np.random.Generator, first_id: int = 31, n: int = 10) -> List[MaterialDef]:
    """`Mat_0..9` with the distributions of Materials.cpp:104-125 (the reference draws them from
    libc rand() at every process start, so only the distribution is contractual)."""
    out = []
    for i in range(n):
        rgb = (int(rng.integers(255)) << 16) | (int(rng.integers(255)) << 8) | int(rng.integers(255))
        typ = (SA if rng.integers(2) == 0 else G) if rng.integers(2) == 0 else LQ
        dens = {SA: 5.0, LQ: 4.0, G: 3.0}[typ] + int(rng.integers(1000)) / 1000.0
        alpha = 255 if typ == SA else int(rng.integers(192)) + 63
        out.append(MaterialDef(first_id + i, f"Mat_{i}", typ, 10, alpha, float(np.float32(dens)),
                               int(rng.integers(4)) + 1, color=rgb))
    return out


def table_struct(defs: List[MaterialDef]):
    arr = (abi.Material * len(defs))()
    for i, d in enumerate(defs):
        assert d.id == i, "material ids must be dense and ordered"
        arr[i] = d.to_struct()
    return arr


def fmix32(h: int) -> int:
    """fss_fmix32 of include/fss_rng.h"""
    h &= 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def rng_tick_key(seed, tick_ct, iter_):
    return fmix32(seed ^ fmix32((tick_ct * 8 + iter_ + 0x9E3779B9) & 0xFFFFFFFF))


def rng_cell_key(tkey, x, y):
    return fmix32(tkey ^ ((x * 0x9E3779B1) & 0xFFFFFFFF) ^ ((y * 0x85EBCA77) & 0xFFFFFFFF))


def rng_rand(cellkey, site, k=0):
    return fmix32((cellkey + site * 0x27D4EB2F + k * 0x165667B1) & 0xFFFFFFFF) >> 1


def synthetic_texture(index: int, w: int = 128, h: int = 128) -> np.ndarray:
    """Deterministic stand-in for the reference's PNG assets (assets are not part of the path;
    tests and benchmarks need *some* 128x128 ARGB surface per Textures::* static)."""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.uint64)
    v = (np.uint64(index) * np.uint64(0x9E3779B1) + ys * np.uint64(128) + xs) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(16)
    v = (v * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(13)
    v = (v * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(16)
    return (np.uint64(0xFF000000) | (v & np.uint64(0x00FFFFFF))).astype(np.uint32)


def create_cell(defs: List[MaterialDef], textures, mat: int, x: int, y: int, cellkey: int = 0, k: int = 0):
    """Tiles::create(mat, x, y) (Tiles.cpp:190-268) evaluated from the recipe; returns one CELL_DTYPE record."""
    d = defs[mat]
    c = np.zeros((), dtype=abi.CELL_DTYPE)
    base = d.color if d.create_base is None else d.create_base
    if d.create_kind == abi.CREATE_FLAT:
        col = base
    elif d.create_kind == abi.CREATE_RAND:
        col = base + ((rng_rand(cellkey, d.create_rand_site, k) % d.create_rand_mod) << d.create_rand_shift)
    else:
        tex = textures[d.create_texture]
        th, tw = tex.shape
        col = int(tex[y % th, x % tw])
    c["material"] = mat
    c["color"] = col & 0xFFFFFFFF
    c["temperature"] = d.create_temperature
    c["fluid_amount"] = 2.0  # MaterialInstance.hpp:21
    return c
