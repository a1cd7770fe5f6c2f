"""Host-side mirror of the reference's `World` for the tick path, on top of the C ABI (include/fss.h).

`World` keeps what the reference's `World` keeps for this path -- the tick zone (world.hpp:132), `tickCt`
(world.cpp:2017), the material table -- and forwards `tick()` / `tick_temperature()` (world.cpp:1091, :2043)
to libfss_b200.so, where the grid lives on the GPU.  Cells cross the boundary as numpy arrays of
`abi.CELL_DTYPE` records (MaterialInstance minus the host-only `id`).

There is no CPU implementation behind this class: if the CUDA library is missing or no sm_100 device is
present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from . import materials as M

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSS_LIB") or os.path.join(HERE, "libfss_b200.so")  # FSS_LIB: kernel-tuning experiments only
_lib = None


class FssError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{abi.ERRORS.get(code, code)}: {msg}")
        self.code = code


def load():
    """dlopen libfss_b200.so and declare the prototypes of include/fss.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FssError(-4, f"{LIB_PATH} is not built (python -m fallingsandsurvival_b200.build); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
    protos = {
        "fss_create": [C.POINTER(abi.Config), C.POINTER(vp)],
        "fss_destroy": [vp],
        "fss_abi_version": [],
        "fss_set_stream": [vp, vp],
        "fss_set_materials": [vp, vp, i32],
        "fss_set_texture": [vp, i32, vp, i32, i32],
        "fss_set_tick_zone": [vp, i32, i32, i32, i32],
        "fss_upload_rect": [vp, i32, i32, i32, i32, vp, sz],
        "fss_download_rect": [vp, i32, i32, i32, i32, vp, sz],
        "fss_shift_grid": [vp, i32, i32],
        "fss_download_flow": [vp, i32, i32, i32, i32, vp, vp, sz],
        "fss_clear_flow": [vp],
        "fss_merge_chunk_records": [vp, i32, i32, i32, i32, vp],
        "fss_read_chunk_records": [vp, i32, i32, i32, i32, vp],
        "fss_chunk_file_encode": [C.c_int8, vp, vp, vp, vp, sz, C.POINTER(sz)],
        "fss_chunk_file_decode": [vp, sz, C.POINTER(C.c_int8), vp, vp, vp],
        "fss_lz4_compress": [vp, sz, vp, sz, C.POINTER(sz)],
        "fss_lz4_decompress": [vp, sz, vp, sz, C.POINTER(sz)],
        "fss_tick": [vp, u32],
        "fss_tick_host": [vp, u32, vp, sz],
        "fss_substep": [vp, u32, i32, i32],
        "fss_tick_temperature": [vp],
        "fss_drain_particles": [vp, vp, sz, C.POINTER(sz), C.POINTER(sz)],
        "fss_tick_particles": [vp],
        "fss_particles_add": [vp, vp, sz],
        "fss_particles_count": [vp, C.POINTER(sz)],
        "fss_particles_download": [vp, vp, sz, C.POINTER(sz)],
        "fss_particles_clear": [vp],
        "fss_particles_rounds": [vp, C.POINTER(C.c_uint64 * 2)],
        "fss_set_material_looks": [vp, vp, vp, i32],
        "fss_mark_dirty": [vp, i32, i32, i32, i32],
        "fss_mark_dirty_list": [vp, vp, sz],
        "fss_render_dirty": [vp, vp, i32, C.POINTER(abi.RenderInfo)],
        "fss_download_pixels": [vp, i32, i32, i32, i32, i32, vp, sz],
        "fss_download_dirty": [vp, i32, i32, i32, i32, vp, sz],
        "fss_histogram": [vp, vp, i32],
        "fss_hash": [vp, C.POINTER(C.c_uint64)],
        "fss_fluid_totals": [vp, vp, i32],
        "fss_sync": [vp],
        "fss_fill_synthetic": [vp, C.POINTER(abi.FillSpec)],
        "fss_comm_unique_id": [vp],
        "fss_comm_attach": [vp, vp, i32, i32],
        "fss_exchange_halo": [vp, i32],
        "fss_exchange_halo_temperature": [vp],
        "fss_exchange_halo_peer": [vp, vp, i32],
        "fss_shift_grid_peer": [vp, i32, i32, i32],
        "fss_exchange_halo_temperature_peer": [vp, vp],
        "fss_get_stats": [vp, C.POINTER(abi.Stats)],
    }
    for name, args in protos.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.fss_chunk_file_bound.restype = C.c_size_t
    L.fss_chunk_file_bound.argtypes = []
    L.fss_last_error.restype = C.c_char_p
    L.fss_last_error.argtypes = []
    if L.fss_abi_version() != abi.FSS_ABI_VERSION:
        raise FssError(-1, f"ABI version {L.fss_abi_version()} != {abi.FSS_ABI_VERSION}")
    _lib = L
    return L


def _check(L, rc):
    if rc != 0:
        raise FssError(rc, L.fss_last_error().decode(errors="replace"))


class World:
    """One device-resident world (or one row strip of it).

    cells: (H, W) array of abi.CELL_DTYPE to upload, or None to leave the grid empty (see fill_synthetic);
    size:  (W, H) when cells is None;  zone: tickZone (x, y, w, h);
    strip: (y0, rows) owned global rows of a strip-sharded world, None = the whole world.
    """

    def __init__(self, cells=None, zone=None, defs=None, textures=None, seed=0xF55, particles=False,
                 condense_material=M.WATER, device=0, size=None, strip=None, particle_capacity=0, early_out=True, track_flow=False,
                 track_dirty=False):
        self.L = load()
        if cells is not None:
            h, w = cells.shape
        else:
            w, h = size
        self.w, self.h = w, h
        self.defs = defs or M.named_materials()
        self.textures = textures if textures is not None else [M.synthetic_texture(i) for i in range(M.N_TEXTURES)]
        cfg = abi.Config()
        cfg.struct_size = C.sizeof(abi.Config)
        cfg.width, cfg.height, cfg.seed, cfg.device = w, h, seed, device
        cfg.flags = (abi.PARTICLES_EMIT if particles else 0) | (0 if early_out else abi.NO_EARLY_OUT) | (abi.TRACK_FLOW if track_flow else 0) | (abi.TRACK_DIRTY if track_dirty else 0)
        cfg.particle_capacity = particle_capacity
        cfg.strip_y0, cfg.strip_rows = strip if strip is not None else (0, 0)
        cfg.condense_material = condense_material
        self.strip = strip
        self.h_ = C.c_void_p()
        _check(self.L, self.L.fss_create(C.byref(cfg), C.byref(self.h_)))
        tbl = M.table_struct(self.defs)
        _check(self.L, self.L.fss_set_materials(self.h_, C.byref(tbl), len(self.defs)))
        for i, t in enumerate(self.textures):
            t = np.ascontiguousarray(t, dtype=np.uint32)
            _check(self.L, self.L.fss_set_texture(self.h_, i, t.ctypes.data, t.shape[1], t.shape[0]))
        if zone is not None:
            self.set_tick_zone(*zone)
        st = self.stats()
        self.stored = (st.storage_y0, st.storage_y0 + st.storage_rows)
        if cells is not None:
            y0, y1 = self.stored
            self.upload(0, y0, cells[y0:y1])
        self.tick_ct = 0

    # ---- World::tickZone ----
    def set_tick_zone(self, x, y, w, h):
        _check(self.L, self.L.fss_set_tick_zone(self.h_, x, y, w, h))
        self.zone = (x, y, w, h)

    # ---- host mirror <-> device ----
    def upload(self, x, y, cells):
        cells = np.ascontiguousarray(cells, dtype=abi.CELL_DTYPE)
        h, w = cells.shape
        _check(self.L, self.L.fss_upload_rect(self.h_, x, y, w, h, cells.ctypes.data, w))

    def download(self, x=0, y=None, w=None, h=None, out=None):
        y0, y1 = self.stored
        y = y0 if y is None else y
        w = self.w - x if w is None else w
        h = y1 - y if h is None else h
        if out is None:
            out = np.zeros((h, w), dtype=abi.CELL_DTYPE)
        _check(self.L, self.L.fss_download_rect(self.h_, x, y, w, h, out.ctypes.data, out.strides[0] // abi.CELL_DTYPE.itemsize))
        return out

    def cells(self):
        """the stored rows of this world (the whole grid unless strip-sharded)"""
        return self.download()

    def merge_chunk_records(self, x0, y0, records):
        """Chunk::read + the merge loop of World::frame for the tiles layer: records = (h, w) array of abi.RECORD_DTYPE"""
        records = np.ascontiguousarray(records, dtype=abi.RECORD_DTYPE)
        h, w = records.shape
        _check(self.L, self.L.fss_merge_chunk_records(self.h_, x0, y0, w, h, records.ctypes.data))

    def read_chunk_records(self, x0, y0, w, h):
        out = np.zeros((h, w), dtype=abi.RECORD_DTYPE)
        _check(self.L, self.L.fss_read_chunk_records(self.h_, x0, y0, w, h, out.ctypes.data))
        return out

    def flow(self):
        """World::flowX, flowY of the stored rows as (rows, W) float32 arrays (needs track_flow=True)"""
        y0, y1 = self.stored
        fx = np.zeros((y1 - y0, self.w), dtype=np.float32)
        fy = np.zeros((y1 - y0, self.w), dtype=np.float32)
        _check(self.L, self.L.fss_download_flow(self.h_, 0, y0, self.w, y1 - y0, fx.ctypes.data, fy.ctypes.data, self.w))
        return fx, fy

    def clear_flow(self):
        _check(self.L, self.L.fss_clear_flow(self.h_))

    def shift_grid(self, change_x, change_y):
        """World::tickChunks' camera shift (world.cpp:2617-2630)"""
        _check(self.L, self.L.fss_shift_grid(self.h_, change_x, change_y))

    # ---- the hot path ----
    def tick(self, tick_ct=None):
        if tick_ct is None:
            tick_ct = self.tick_ct
        _check(self.L, self.L.fss_tick(self.h_, tick_ct))
        self.tick_ct = tick_ct + 1

    def tick_host(self, cells, tick_ct=None):
        """World::tick() for a grid kept on the host: `cells` = the stored rows (rows, W) of abi.CELL_DTYPE, updated in place;
        upload, the 24 sub-steps and download overlap band by band (fss_tick_host); pin the array for the overlap to happen"""
        if tick_ct is None:
            tick_ct = self.tick_ct
        assert cells.dtype == abi.CELL_DTYPE and cells.flags["C_CONTIGUOUS"] and cells.shape == (self.stored[1] - self.stored[0], self.w)
        _check(self.L, self.L.fss_tick_host(self.h_, tick_ct, cells.ctypes.data, self.w))
        self.tick_ct = tick_ct + 1

    def substep(self, tick_ct, iter_, tk):
        _check(self.L, self.L.fss_substep(self.h_, tick_ct, iter_, tk))

    def tick_temperature(self):
        _check(self.L, self.L.fss_tick_temperature(self.h_))

    def sync(self):
        _check(self.L, self.L.fss_sync(self.h_))

    def set_stream(self, cuda_stream):
        _check(self.L, self.L.fss_set_stream(self.h_, C.c_void_p(cuda_stream)))

    # ---- side outputs ----
    def drain_particles(self, cap=None):
        """the spawn records since the last drain, in World::tick's push order (cap: ignored, the buffer is sized by a query)"""
        n, total = C.c_size_t(), C.c_size_t()
        _check(self.L, self.L.fss_drain_particles(self.h_, None, 0, C.byref(n), C.byref(total)))
        out = np.zeros(max(total.value, 1), dtype=abi.PARTICLE_DTYPE)
        _check(self.L, self.L.fss_drain_particles(self.h_, out.ctypes.data, total.value, C.byref(n), C.byref(total)))
        return out[: n.value]

    # ---- World::dirty and the dirty -> pixels job of Game::tick (Game.cpp:2579-2650); needs track_dirty ----
    def set_material_looks(self, alpha, emit_color):
        alpha = np.ascontiguousarray(alpha, dtype=np.uint8)
        emit_color = np.ascontiguousarray(emit_color, dtype=np.uint32)
        _check(self.L, self.L.fss_set_material_looks(self.h_, alpha.ctypes.data, emit_color.ctypes.data, len(alpha)))

    def mark_dirty(self, x, y, w, h):
        _check(self.L, self.L.fss_mark_dirty(self.h_, x, y, w, h))

    def mark_dirty_list(self, rects):
        """rects: (n, 4) int32 array of x, y, w, h -- one kernel launch for all of them"""
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
        _check(self.L, self.L.fss_mark_dirty_list(self.h_, rects.ctypes.data, len(rects)))

    def render_dirty(self):
        """(moving_tiles per material, (had_dirty, had_fire, had_flow)); the pixel layers stay on the device"""
        moving = np.zeros(len(self.defs), dtype=np.uint64)
        info = abi.RenderInfo()
        _check(self.L, self.L.fss_render_dirty(self.h_, moving.ctypes.data, len(self.defs), C.byref(info)))
        return moving, (info.had_dirty, info.had_fire, info.had_flow)

    def pixels(self, layer, x=0, y=0, w=None, h=None):
        """(h, w, 4) uint8, bytes r g b a, of layer 0 main / 1 fire / 2 emission / 3 flow"""
        w = self.w - x if w is None else w
        h = self.h - y if h is None else h
        out = np.zeros((h, w, 4), dtype=np.uint8)
        _check(self.L, self.L.fss_download_pixels(self.h_, layer, x, y, w, h, out.ctypes.data, w * 4))
        return out

    def dirty(self, x=0, y=0, w=None, h=None):
        w = self.w - x if w is None else w
        h = self.h - y if h is None else h
        out = np.zeros((h, w), dtype=np.uint8)
        _check(self.L, self.L.fss_download_dirty(self.h_, x, y, w, h, out.ctypes.data, w))
        return out

    # ---- World::particles on the device (world.cpp:2171-2311) ----
    def tick_particles(self):
        """World::tickParticles() after adopting the spawn records of the ticks since the last call"""
        _check(self.L, self.L.fss_tick_particles(self.h_))

    def add_particles(self, parts):
        parts = np.ascontiguousarray(parts, dtype=abi.LIVE_PARTICLE_DTYPE)
        _check(self.L, self.L.fss_particles_add(self.h_, parts.ctypes.data, parts.size))

    def particles_count(self):
        n = C.c_size_t()
        _check(self.L, self.L.fss_particles_count(self.h_, C.byref(n)))
        return n.value

    def particle_list(self):
        n = self.particles_count()
        out = np.zeros(max(n, 1), dtype=abi.LIVE_PARTICLE_DTYPE)
        got = C.c_size_t()
        _check(self.L, self.L.fss_particles_download(self.h_, out.ctypes.data, n, C.byref(got)))
        return out[: got.value]

    def particle_rounds(self):
        r = (C.c_uint64 * 2)()
        _check(self.L, self.L.fss_particles_rounds(self.h_, r))
        return int(r[0]), int(r[1])

    def clear_particles(self):
        _check(self.L, self.L.fss_particles_clear(self.h_))

    def histogram(self):
        counts = np.zeros(len(self.defs), dtype=np.uint64)
        _check(self.L, self.L.fss_histogram(self.h_, counts.ctypes.data, len(self.defs)))
        return counts

    def hash(self):
        d = C.c_uint64()
        _check(self.L, self.L.fss_hash(self.h_, C.byref(d)))
        return d.value

    def fluid_totals(self):
        t = np.zeros(len(self.defs), dtype=np.float64)
        _check(self.L, self.L.fss_fluid_totals(self.h_, t.ctypes.data, len(self.defs)))
        return t

    def stats(self):
        s = abi.Stats()
        _check(self.L, self.L.fss_get_stats(self.h_, C.byref(s)))
        return s

    def fill_synthetic(self, mix, seed=0xF55, border=16, border_material=M.SOLID, solid_from_row=0,
                       solid_material=M.COBBLE_STONE, n_boxes=0, box_size=0, sparse_material=0, sparse_per_mille=0):
        """device-side twin of worlds.synthetic_world(); mix = [(material, percent), ...]"""
        s = abi.FillSpec()
        s.seed, s.border, s.border_material = seed, border, border_material
        s.n_entries = len(mix)
        cum = 0
        for i, (m, pct) in enumerate(mix):
            cum += pct
            s.material[i], s.cumulative_percent[i] = m, cum
        s.solid_from_row, s.solid_material = solid_from_row, solid_material
        s.n_boxes, s.box_size, s.sparse_material, s.sparse_per_mille = n_boxes, box_size, sparse_material, sparse_per_mille
        _check(self.L, self.L.fss_fill_synthetic(self.h_, C.byref(s)))

    # ---- strips over NCCL ----
    def comm_attach(self, unique_id: bytes, rank: int, n_ranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        _check(self.L, self.L.fss_comm_attach(self.h_, buf, rank, n_ranks))

    def close(self):
        if getattr(self, "h_", None):
            self.L.fss_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    L = load()
    buf = C.create_string_buffer(128)
    _check(L, L.fss_comm_unique_id(buf))
    return buf.raw


class StripWorld:
    """A world cut into row strips inside ONE process (one strip per entry of `devices`), halo rows moved by
    peer copies (fss_exchange_halo_peer).  Same surface as World; used to check the strip schedule on a single
    GPU and as the single-process multi-GPU mode.  The multi-process mode (one rank per GPU, NCCL) is
    strips.RankWorld."""

    def __init__(self, cells, zone, n_strips, devices=None, **kw):
        from . import strips
        h, w = cells.shape
        self.w, self.h, self.zone = w, h, tuple(zone)
        self.plan = strips.plan_strips(h, zone, n_strips)
        devices = devices or [0] * n_strips
        self.parts = [World(cells, zone, strip=s, device=d, **kw) for s, d in zip(self.plan, devices)]
        self.L = self.parts[0].L
        self.tick_ct = 0

    def tick(self, tick_ct=None):
        from . import strips
        if tick_ct is None:
            tick_ct = self.tick_ct
        for op, it, tk in strips.tick_schedule():
            if op == "substep":
                for p in self.parts:
                    p.substep(tick_ct, it, tk)
            else:
                for a, b in zip(self.parts[:-1], self.parts[1:]):
                    _check(self.L, self.L.fss_exchange_halo_peer(a.h_, b.h_, tk))
        self.tick_ct = tick_ct + 1

    def tick_temperature(self):
        for p in self.parts:
            p.tick_temperature()
        for a, b in zip(self.parts[:-1], self.parts[1:]):
            _check(self.L, self.L.fss_exchange_halo_temperature_peer(a.h_, b.h_))

    # ---- World::dirty and the dirty -> pixels job: every strip keeps, renders and hands out the rows it owns ----
    def dirty(self):
        out = np.zeros((self.h, self.w), dtype=np.uint8)
        for p, (y0, rows) in zip(self.parts, self.plan):
            out[y0:y0 + rows] = p.dirty(0, y0, self.w, rows)
        return out

    def mark_dirty(self, x, y, w, h):
        for p in self.parts:
            p.mark_dirty(x, y, w, h)

    def render_dirty(self):
        moving, info = None, (0, 0, 0)
        for p in self.parts:
            m, i = p.render_dirty()
            moving = m if moving is None else moving + m
            info = tuple(a | b for a, b in zip(info, i))
        return moving, info

    def pixels(self, layer):
        out = np.zeros((self.h, self.w, 4), dtype=np.uint8)
        for p, (y0, rows) in zip(self.parts, self.plan):
            out[y0:y0 + rows] = p.pixels(layer, 0, y0, self.w, rows)
        return out

    def flow(self):
        fx = np.zeros((self.h, self.w), dtype=np.float32)
        fy = np.zeros((self.h, self.w), dtype=np.float32)
        for p, (y0, rows) in zip(self.parts, self.plan):
            a, b = p.flow()  # the stored rows
            lo = y0 - p.stored[0]
            fx[y0:y0 + rows], fy[y0:y0 + rows] = a[lo:lo + rows], b[lo:lo + rows]
        return fx, fy

    def set_material_looks(self, alpha, emit):
        for p in self.parts:
            p.set_material_looks(alpha, emit)

    def shift_grid(self, change_x, change_y):
        """World::tickChunks' camera shift (world.cpp:2617-2630) across the strips: rows that cross a boundary come from the neighbour"""
        arr = (C.c_void_p * len(self.parts))(*[p.h_ for p in self.parts])
        _check(self.L, self.L.fss_shift_grid_peer(arr, len(self.parts), change_x, change_y))

    def cells(self):
        out = np.zeros((self.h, self.w), dtype=abi.CELL_DTYPE)
        for p, (y0, rows) in zip(self.parts, self.plan):
            p.download(0, y0, self.w, rows, out=out[y0:y0 + rows])
        return out

    def hash(self):
        return sum(p.hash() for p in self.parts) & 0xFFFFFFFFFFFFFFFF

    def drain_particles(self):
        recs = np.concatenate([p.drain_particles() for p in self.parts])
        order = abi.spawn_order(recs, self.zone)
        return recs[order]

    def close(self):
        for p in self.parts:
            p.close()


# ---- the reference's chunk files (Chunk::read / Chunk::write, Chunk.cpp:47-281): host code, no world, no GPU ----
FILE_CELLS = 128 * 128


def chunk_file_encode(generation_phase, tiles, layer2, background) -> bytes:
    """tiles, layer2: (128, 128) arrays of abi.RECORD_DTYPE; background: (128, 128) uint32"""
    L = load()
    tiles = np.ascontiguousarray(tiles, dtype=abi.RECORD_DTYPE)
    layer2 = np.ascontiguousarray(layer2, dtype=abi.RECORD_DTYPE)
    background = np.ascontiguousarray(background, dtype=np.uint32)
    assert tiles.size == layer2.size == background.size == FILE_CELLS
    cap = L.fss_chunk_file_bound()
    out = np.zeros(cap, dtype=np.uint8)
    n = C.c_size_t()
    _check(L, L.fss_chunk_file_encode(generation_phase, tiles.ctypes.data, layer2.ctypes.data, background.ctypes.data, out.ctypes.data, cap, C.byref(n)))
    return out[: n.value].tobytes()


def chunk_file_decode(data: bytes):
    """-> (generation_phase, tiles, layer2, background)"""
    L = load()
    buf = np.frombuffer(data, dtype=np.uint8)
    tiles = np.zeros((128, 128), dtype=abi.RECORD_DTYPE)
    layer2 = np.zeros((128, 128), dtype=abi.RECORD_DTYPE)
    background = np.zeros((128, 128), dtype=np.uint32)
    phase = C.c_int8()
    _check(L, L.fss_chunk_file_decode(buf.ctypes.data, len(buf), C.byref(phase), tiles.ctypes.data, layer2.ctypes.data, background.ctypes.data))
    return phase.value, tiles, layer2, background


def lz4_compress(data: bytes) -> bytes:
    L = load()
    src = np.frombuffer(data, dtype=np.uint8)
    cap = len(src) + len(src) // 255 + 16
    out = np.zeros(cap, dtype=np.uint8)
    n = C.c_size_t()
    _check(L, L.fss_lz4_compress(src.ctypes.data if len(src) else None, len(src), out.ctypes.data, cap, C.byref(n)))
    return out[: n.value].tobytes()


def lz4_decompress(data: bytes, size: int) -> bytes:
    L = load()
    src = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(max(size, 1), dtype=np.uint8)
    n = C.c_size_t()
    _check(L, L.fss_lz4_decompress(src.ctypes.data, len(src), out.ctypes.data, size, C.byref(n)))
    return out[: n.value].tobytes()
