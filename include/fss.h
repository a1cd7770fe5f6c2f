/*
 * fss.h -- C ABI of the B200-native falling-sand core.
 *
 * The reference (PieKing1215/FallingSandSurvival) has no plugin / FFI layer: `World` is a
 * C++ class with public raw arrays (FallingSandSurvival/world.hpp:98-146) and the two hot
 * functions `World::tick()` (world.cpp:1091-2041) and `World::tickTemperature()`
 * (world.cpp:2043-2143).  This header is the boundary a maintainer binds instead of the
 * bodies of those two functions: the grid lives on the GPU, the host keeps its
 * `MaterialInstance tiles[]` mirror and syncs rectangles through fss_upload_rect /
 * fss_download_rect.  INTEGRATION.md shows the reference-side shim.
 *
 * Conventions: every function returns 0 on success or a negative FSS_ERR_* code;
 * fss_last_error() returns a human-readable description of the last failure on the calling
 * thread.  No C++ types, exceptions or torch types cross this boundary.  A world is
 * externally synchronised (one caller thread at a time, like the reference's main-thread
 * `Game::tick`, Game.cpp:2428-2430, :2759-2761) and internally asynchronous on one CUDA
 * stream; fss_download_* / fss_sync are the synchronisation points.
 */
#ifndef FSS_H
#define FSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSS_ABI_VERSION 2

/* error codes */
#define FSS_OK 0
#define FSS_ERR_INVALID (-1)   /* bad argument */
#define FSS_ERR_ALIGN (-2)     /* tick zone / strip not aligned to the micro-chunk grid */
#define FSS_ERR_CUDA (-3)      /* CUDA runtime failure, see fss_last_error() */
#define FSS_ERR_NO_DEVICE (-4) /* no usable sm_100 device: there is NO CPU fallback */
#define FSS_ERR_STATE (-5)     /* call out of order (e.g. tick before materials) */
#define FSS_ERR_RANGE (-6)     /* rectangle outside this world's storage */
#define FSS_ERR_COMM (-7)      /* NCCL failure */

/* micro-chunk geometry of the checkerboard schedule (the reference's CHUNK_W/CHUNK_H,
 * Chunk.hpp:18-19, shrunk; see DESIGN.md "schedule") */
#define FSS_CHUNK_W 4
#define FSS_CHUNK_H 16
#define FSS_ZONE_ALIGN_X 8  /* = 2*CHUNK_W: tick zone x, w are multiples of this */
#define FSS_ZONE_ALIGN_Y 32 /* = 2*CHUNK_H: tick zone h and strip heights are multiples */
#define FSS_ZONE_MARGIN 16  /* tick zone must stay this far from every world edge */

/* PhysicsType.hpp:6-12 */
enum {
    FSS_PHYS_AIR = 0,
    FSS_PHYS_SOLID = 1,
    FSS_PHYS_SAND = 2,
    FSS_PHYS_SOUP = 3,
    FSS_PHYS_GAS = 4,
    FSS_PHYS_PASSABLE = 5, /* == OBJECT */
};

/* Material.hpp:18-19 */
#define FSS_REACT_TEMPERATURE_BELOW 4
#define FSS_REACT_TEMPERATURE_ABOVE 5

#define FSS_MAX_MATERIALS 256
#define FSS_MAX_REACTIONS 2
#define FSS_MAX_TEXTURES 32

/* fss_material.flags */
#define FSS_MAT_IS_FIRE 1u  /* `mat->id == Materials::FIRE.id`  (world.cpp:1171) */
#define FSS_MAT_IS_STEAM 2u /* `mat->id == Materials::STEAM.id` (world.cpp:1965) */

/* fss_material.create_kind: how Tiles::create(mat, x, y) (Tiles.cpp:190-268) colours a new cell */
#define FSS_CREATE_FLAT 0    /* color = create_base                                   */
#define FSS_CREATE_RAND 1    /* color = create_base + ((rand() % mod) << shift)        */
#define FSS_CREATE_TEXTURE 2 /* color = texture[create_texture] at (x mod w, y mod h)   */

/* One cell in the host exchange format (row-major arrays of these).
 * Mirrors MaterialInstance (MaterialInstance.hpp:12-29) minus `id`, the host-only unique
 * counter (MaterialInstance.cpp:7), with `Material* mat` replaced by its id. */
typedef struct fss_cell {
    uint16_t material;       /* Material::id */
    uint8_t moved;           /* MaterialInstance::moved (sand: friction released; liquid: settled) */
    uint8_t settle_count;    /* MaterialInstance::settleCount */
    uint32_t color;          /* MaterialInstance::color */
    int32_t temperature;     /* MaterialInstance::temperature */
    float fluid_amount;      /* MaterialInstance::fluidAmount (constructor default 2.0f) */
    float fluid_amount_diff; /* MaterialInstance::fluidAmountDiff */
} fss_cell;

typedef struct fss_reaction {
    int32_t type;  /* FSS_REACT_TEMPERATURE_BELOW / _ABOVE (MaterialInteraction::type) */
    int32_t data1; /* temperature threshold */
    int32_t data2; /* target material id */
} fss_reaction;

/* One row of Materials::MATERIALS (Material.hpp:29-59, Materials.cpp:4-199) plus the
 * colour recipe of the matching Tiles::create* factory. */
typedef struct fss_material {
    int32_t physics_type;
    int32_t iterations;
    int32_t slipperyness;
    float density;
    uint32_t add_temp; /* Uint32 in the reference (Material.hpp:40) */
    float conduction_self;
    float conduction_other;
    uint32_t color; /* Material::color */
    uint32_t flags; /* FSS_MAT_IS_* */
    int32_t n_reactions; /* evaluated only if `react` (we store 0 when react is false) */
    fss_reaction reactions[FSS_MAX_REACTIONS];
    int32_t create_kind;
    uint32_t create_base;
    uint32_t create_rand_mod;
    uint32_t create_rand_shift;
    uint32_t create_rand_site; /* RNG site of the colour draw (FSS_SITE_TILES + Tiles.cpp line) */
    int32_t create_texture;    /* index given to fss_set_texture */
    int32_t create_temperature; /* Tiles::createWater -1023, createLava 1024, else 0 */
} fss_material;

/* fss_config.flags */
#define FSS_PARTICLES_EMIT 1u /* reference behaviour: free-falling sand/liquid and fire sparks leave the
                                 grid as Particle spawn records (world.cpp:1180, :1288-1294, :1352-1374).
                                 Without it those three sites are disabled and cells keep falling in-grid. */

#define FSS_NO_EARLY_OUT 2u   /* visit every micro-chunk in every sub-step like the reference does, instead of skipping
                                 chunks verified settled (results are identical either way; for measurements) */

#define FSS_TRACK_FLOW 4u     /* keep World::flowX / flowY (world.hpp:99-100): the SOUP rule adds every flow it moves to the source
                                 cell's entry (world.cpp:1405, :1447, :1477, :1509); the renderer reads and zeroes them
                                 (Game.cpp:2631-2646).  Two more 4-byte planes; off by default. */

#define FSS_TRACK_DIRTY 8u    /* keep World::dirty (world.hpp:139) on the device: every site of World::tick / tickParticles that sets
                                 it does (world.cpp:1061, :1194 ... :1968, :2269-2303), fss_mark_dirty for host edits, consumed by
                                 fss_render_dirty.  One more byte per cell; a compile-time kernel variant, off by default.  Works with
                                 the settled-chunk early-out: the reference re-marks every liquid cell in every sub-step
                                 (world.cpp:1814, :1823), so for a chunk whose visit is skipped the kernel notes the iter and the
                                 marks are added before World::dirty is next looked at.  Strip-sharded worlds: the marks a strip
                                 leaves in the ghost rows next to its boundary travel with the halo exchange to the strip that
                                 owns the rows; every strip marks, renders and hands out the rows it owns. */

typedef struct fss_config {
    uint32_t struct_size; /* sizeof(fss_config), for ABI evolution */
    uint32_t width;       /* world width in cells (reference: uint16_t, world.hpp:106) */
    uint32_t height;      /* GLOBAL world height */
    uint32_t seed;        /* RNG seed, see fss_rng.h */
    int32_t device;       /* CUDA device ordinal */
    uint32_t flags;
    uint32_t particle_capacity; /* max spawn records kept per drain interval (0 = default 1<<20) */
    /* row-strip sharding (one world object per GPU).  strip_rows == 0 means the whole world. */
    uint32_t strip_y0;   /* first owned global row, multiple of FSS_ZONE_ALIGN_Y from the zone origin */
    uint32_t strip_rows; /* owned rows */
    /* steam condensation product (world.cpp:1967 Tiles::createWater()) */
    int32_t condense_material; /* material id created when blocked steam condenses; -1 disables */
} fss_config;

/* particle spawn record: the host builds its Particle objects from these with the reference's
 * velocity formulas (world.cpp:1180, :1291, :1366) and fss_rand(cellkey, site, k). */
typedef struct fss_particle_spawn {
    uint32_t x, y;     /* cell that spawned it (the reference spawns at y-1 / y+1, site tells which) */
    uint32_t site;     /* 1180 fire spark, 1291 falling sand, 1366 falling liquid */
    uint32_t cell_key; /* fss_rng_cell_key of the visit, for the velocity draws */
    uint32_t tick_iter; /* tickCt * 8 + iter */
    fss_cell cell;     /* the cell that left the grid (fire sparks: the fire cell itself) */
} fss_particle_spawn;

typedef struct fss_world fss_world;

/* ---- lifecycle -------------------------------------------------------------------------------
 * replaces the array part of World::init / ~World (world.cpp:86-88, :145-184, :3617-3711) */
int fss_create(const fss_config* cfg, fss_world** out);
int fss_destroy(fss_world* w);
const char* fss_last_error(void);
int fss_abi_version(void);

/* run all work of this world on an existing CUDA stream (a cudaStream_t passed as void*);
 * NULL restores the world's own stream.  Lets a host time the path with its own events. */
int fss_set_stream(fss_world* w, void* cuda_stream);

/* ---- material table: call after Materials::init() (Game.cpp:521) and whenever it changes ---- */
int fss_set_materials(fss_world* w, const fss_material* table, int n);
int fss_set_texture(fss_world* w, int texture_index, const uint32_t* argb, int tex_w, int tex_h);

/* ---- World::tickZone (world.hpp:132, set by Game.cpp:2216) ---------------------------------- */
int fss_set_tick_zone(fss_world* w, int x, int y, int width, int height);

/* ---- host mirror <-> device, rectangles in GLOBAL cell coordinates ---------------------------
 * `pitch_cells` = cells per row of the host array.  Replaces the direct `tiles[]` indexing of
 * Game.cpp (brush :1147, rigid-body stamping :2355-2372, chunk merge world.cpp:2530-2543, ...). */
int fss_upload_rect(fss_world* w, int x, int y, int width, int height, const fss_cell* host, size_t pitch_cells);
int fss_download_rect(fss_world* w, int x, int y, int width, int height, fss_cell* host, size_t pitch_cells);

/* ---- chunk records: the reference's on-disk cell format (after LZ4, which stays on the host) ----------------
 * MaterialInstanceData, Chunk.hpp:27-31: 12 bytes per cell.  fss_merge_chunk_records does what Chunk::read
 * (Chunk.cpp:139-156) + the merge loop of World::frame (world.cpp:2526-2546) do for the `tiles` layer: the cell at
 * (x0 + i, y0 + j) becomes MaterialInstance(MATERIALS[index], color, temperature) with the constructor's defaults for
 * the fields the file does not store (moved = false, fluidAmount = 2.0, fluidAmountDiff = 0, settleCount = 0); cells
 * outside the world are skipped (`if(tx < 0 || tx >= width ...) continue`).  fss_read_chunk_records is the way back
 * for Chunk::write (Chunk.cpp:189-281); cells outside the world read as material 0. */
typedef struct fss_tile_record {
    uint16_t index;      /* Material::id */
    uint32_t color;
    int32_t temperature;
} fss_tile_record;
int fss_merge_chunk_records(fss_world* w, int x0, int y0, int chunk_w, int chunk_h, const fss_tile_record* records);
int fss_read_chunk_records(fss_world* w, int x0, int y0, int chunk_w, int chunk_h, fss_tile_record* records);

/* The chunk FILE around those records: Chunk::write / Chunk::read (Chunk.cpp:189-281, :47-187).  Layout: int8 generationPhase;
 * int32 src_size, compressed_size, src_size2, compressed_size2; an LZ4 block holding the 128 x 128 records of the `tiles`
 * layer followed by those of `layer2` (MaterialInstanceData, Chunk.hpp:27-31); an LZ4 block holding the 128 x 128 Uint32 of
 * `background`.  Host code only (no GPU, no world): together with fss_read_chunk_records / fss_merge_chunk_records a save
 * round-trips between the device planes and the reference's files without the 40-byte host mirror.  The LZ4 block format is
 * implemented in the library (the reference links liblz4, which this build does not); blocks written here are read by any
 * LZ4 decoder and the other way round.  layer2 and background are not part of the tick path: the caller keeps them. */
#define FSS_CHUNK_FILE_W 128 /* CHUNK_W, CHUNK_H of the reference (Chunk.hpp:18-19): the unit of its files, not the micro-chunk */
#define FSS_CHUNK_FILE_H 128
size_t fss_chunk_file_bound(void); /* buffer size fss_chunk_file_encode needs */
int fss_chunk_file_encode(int8_t generation_phase, const fss_tile_record* tiles, const fss_tile_record* layer2, const uint32_t* background,
                          uint8_t* out, size_t cap, size_t* size);
int fss_chunk_file_decode(const uint8_t* file, size_t size, int8_t* generation_phase, fss_tile_record* tiles, fss_tile_record* layer2,
                          uint32_t* background);
/* the two halves of the codec on their own (LZ4_compress_fast / LZ4_decompress_safe of Chunk.cpp:226, :127); cap of
 * fss_lz4_compress must be at least n + n/255 + 16 */
int fss_lz4_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_size);
int fss_lz4_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_size);

/* World::tickChunks' grid shift when the camera moves (world.cpp:2617-2630): every cell moves by (change_x, change_y);
 * cells whose source would lie outside the grid keep their content, exactly like the reference's in-place loop.  The
 * caller then uploads the freshly loaded chunks into the vacated band.  A device-resident particle list moves with the grid
 * (world.cpp:2729-2731).
 * Strip-sharded worlds: rows cross strip boundaries, so the call is COLLECTIVE -- every rank of the communicator makes it with
 * the same arguments (each strip fetches the |change_y| rows of its neighbour that become its first / last stored rows, NCCL
 * send / recv, then shifts); fss_shift_grid_peer does the same for the strips of one process (`strips` ordered top to bottom).
 * |change_y| must not exceed any strip's rows less the 16 ghost rows (FSS_ERR_RANGE: nothing was moved). */
int fss_shift_grid(fss_world* w, int change_x, int change_y);
int fss_shift_grid_peer(fss_world** strips, int n, int change_x, int change_y);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* one World::tick(): 6 iter x 4 tk checkerboard sub-steps x 3 scans (world.cpp:1108-2011);
 * `tick_ct` is World::tickCt (world.cpp:2017), the RNG key.  Asynchronous (strip-sharded worlds with a communicator: the strips
 * first tell each other which materials they hold -- a 256-byte all-reduce and a host read -- so that a material uploaded into one
 * strip switches the kernels of all of them before it can cross a boundary).  For a strip-sharded
 * world with a communicator attached this also performs the per-sub-step halo exchange. */
int fss_tick(fss_world* w, uint32_t tick_ct);
/* One World::tick() for a caller that keeps `tiles[]` on the host (the shim's worst case: the whole grid crosses PCIe both
 * ways every tick).  Same result as
 *     fss_upload_rect(w, 0, y0, width, rows, cells, pitch); fss_tick(w, tick_ct); fss_download_rect(w, 0, y0, width, rows, cells, pitch);
 * for the stored rows [y0, y0 + rows) of the world (all rows unless strip-sharded; `cells` points at row y0), but pipelined in
 * row bands on three streams: while band b+1 is still being uploaded, the 24 sub-steps already run on the rows that can no
 * longer see it (a sub-step reaches 26 rows down, so the band's lower edge recedes by one 32-row chunk pair per sub-step), and
 * the rows that are final go back to the host meanwhile -- both PCIe directions and the kernels overlap.  Synchronous: the
 * caller owns `cells` again on return.  `cells` should be page-locked (cudaHostAlloc / cudaHostRegister) for the overlap to
 * happen.  Strip-sharded worlds run the plain sequence above. */
int fss_tick_host(fss_world* w, uint32_t tick_ct, fss_cell* cells, size_t pitch_cells);
/* one sub-step (iter, tk) of the above without halo exchange */
int fss_substep(fss_world* w, uint32_t tick_ct, int iter, int tk);
/* one World::tickTemperature() (world.cpp:2043-2143).  Asynchronous. */
int fss_tick_temperature(fss_world* w);

/* World::flowX / flowY of a rectangle (needs FSS_TRACK_FLOW), row-major floats, `pitch` floats per host row; and the
 * reset the consumer does after reading them (Game.cpp:2644-2645) */
int fss_download_flow(fss_world* w, int x, int y, int width, int height, float* flow_x, float* flow_y, size_t pitch);
int fss_clear_flow(fss_world* w);

/* ---- dirty -> pixels: the `dirty` job of Game::tick, Game.cpp:2579-2650, and its memset (:2751) -------------
 * Needs FSS_TRACK_DIRTY (and FSS_TRACK_FLOW for a meaningful flow layer).  For every dirty cell: count it in
 * moving_tiles[material] (Game.cpp:2589; the reference's counters are uint16_t and wrap, these do not), then
 *   AIR:   main, fire and emission pixels := 0, flow accumulators := 0;
 *   else:  main := (r, g, b of MaterialInstance::color, Material::alpha); emission := Material::emitColor; fire layer := main
 *          if the material is fire; liquids: prevFlow += (flow - prevFlow) * 0.25 (negative y halved), flow pixel =
 *          clamp(prevFlow * (3.0 / iterations + 0.5) / 4.0 + 0.5, 0, 1) * 255 in double like the reference; flow := 0.
 * The four RGBA layers (byte order r, g, b, a as the reference writes them) live on the device, row-major, and only change where
 * cells were dirty; fss_download_pixels copies a rectangle of one layer to the host.  dirty is cleared afterwards.
 * Strip-sharded worlds: each strip renders and counts the rows it owns (fss_download_pixels: rectangles inside those rows;
 * moving_tiles and the three flags of the whole world = the sums / ORs over the strips); fss_mark_dirty marks the owned part of
 * a rectangle, so host edits are marked on every strip they touch. */
typedef struct fss_render_info {
    uint32_t had_dirty, had_fire, had_flow; /* Game.cpp:2588, :2626, :2639 */
} fss_render_info;
#define FSS_LAYER_MAIN 0
#define FSS_LAYER_FIRE 1
#define FSS_LAYER_EMISSION 2
#define FSS_LAYER_FLOW 3
/* Material::alpha (Material.hpp:34) and Material::emitColor (:38) per material id; defaults 255 and 0 */
int fss_set_material_looks(fss_world* w, const uint8_t* alpha, const uint32_t* emit_color, int n);
/* host code that writes tiles[] sets dirty itself (World::setTile world.cpp:1061, Game.cpp:1148, :2357 ...): after uploading the
 * cells it marks the same rectangle here */
int fss_mark_dirty(fss_world* w, int x, int y, int width, int height);
/* the same for many rectangles with one kernel launch (a brush stroke or a rigid-body stamp is hundreds of small edits) */
typedef struct fss_rect {
    int32_t x, y, w, h;
} fss_rect;
int fss_mark_dirty_list(fss_world* w, const fss_rect* rects, size_t n);
int fss_render_dirty(fss_world* w, uint64_t* moving_tiles, int n_materials, fss_render_info* info);
int fss_download_pixels(fss_world* w, int layer, int x, int y, int width, int height, uint8_t* rgba, size_t pitch_bytes);
/* World::dirty of a rectangle, one byte per cell (tests) */
int fss_download_dirty(fss_world* w, int x, int y, int width, int height, uint8_t* out, size_t pitch_bytes);

/* ---- side outputs ---------------------------------------------------------------------------- */
/* spawn records since the last drain in the order World::tick appends them to World::particles
 * (fss_spawn_order_key below).  *n_total = records waiting.  If they do not fit into `cap`, NOTHING is drained and the
 * call returns FSS_ERR_RANGE (cap == 0: FSS_OK, a plain query): come back with room for *n_total.  The device buffer holds
 * fss_config.particle_capacity records; a spawn site that finds it full leaves its cell in the grid (it keeps falling
 * there) and is counted in fss_stats.spawns_refused. */
int fss_drain_particles(fss_world* w, fss_particle_spawn* out, size_t cap, size_t* n_out, size_t* n_total);

/* The order of World::particles after a tick: per sub-step (iter, tk) the chunk futures are collected in launch order
 * (cx outer, cy inner: world.cpp:1140-1141, :1989-1994) and every chunk pushes in visit order (rows bottom-up, columns
 * left to right: world.cpp:1154-1157).  Ascending key = that order; equal keys are the n records of one falling
 * liquid cell (world.cpp:1358-1371), whose j-th record is the j-th push.  Ticks enter modulo 2^27. */
#if defined(__CUDACC__)
#define FSS_INLINE __host__ __device__ static inline
#else
#define FSS_INLINE static inline
#endif
FSS_INLINE uint64_t fss_spawn_order_key(uint32_t tick_iter, uint32_t x, uint32_t y, int zone_x, int zone_y) {
    const uint32_t ci = (x - (uint32_t)zone_x) / FSS_CHUNK_W, cj = (y - (uint32_t)zone_y) / FSS_CHUNK_H;
    const uint32_t dx = (x - (uint32_t)zone_x) % FSS_CHUNK_W, dy = (y - (uint32_t)zone_y) % FSS_CHUNK_H;
    const uint32_t tk = (ci & 1u) + 2u * (1u - (cj & 1u)); /* chOfsX = tk % 2, chOfsY = 1 - tk / 2 (world.cpp:1118-1119) */
    return ((uint64_t)(tick_iter & 0x3FFFFFFFu) << 34) | ((uint64_t)tk << 32) | ((uint64_t)(ci & 0x3FFFu) << 18) |
           ((uint64_t)(cj & 0xFFFu) << 6) | ((uint64_t)(FSS_CHUNK_H - 1u - dy) << 2) | dx;
}

/* ---- World::particles on the device ---------------------------------------------------------
 * One Particle (Particle.hpp:12-31) without its killCallback: particles that need a host callback stay on the host. */
typedef struct fss_particle {
    float x, y, vx, vy, ax, ay;
    float target_x, target_y, target_force;
    int32_t lifetime;
    int32_t fade_time;
    uint16_t in_object_state;
    uint8_t phase;
    uint8_t temporary;
    fss_cell tile;
} fss_particle;

/* How a spawn record becomes a Particle (world.cpp:1180-1183, :1291, :1366; `r(k)` = fss_rand(cell_key, site, k); g++ evaluates
 * the constructor's arguments right to left, so the vy draw is the first call at the site and the vx draw the second;
 * j = index of the record among those of the same cell visit):
 *   site 1180  Particle(tile, x, y - 1, (r(1) % 10 - 5) / 20.0f, -((r(0) % 10) / 10.0f) / 3.0f + -0.5f, 0, 0.01f), temporary, lifetime 30, fadeTime 10
 *   site 1291  Particle(tile, x, y + 1, (r(1) % 10 - 5) / 20.0f, -((r(0) % 2) + 3) / 10.0f + 1.5f, 0, 0.1f)
 *   site 1366  Particle(nt,   x, y + 1, (r(2j+1) % 10 - 5) / 30.0f, -((r(2j) % 2) + 3) / 10.0f + 1.0f, 0, 0.1f)
 *
 * fss_tick_particles = one World::tickParticles() (world.cpp:2171-2311) on the device list, after moving the spawn records
 * of the ticks since the last call to its end (so fss_drain_particles then returns nothing).  Particles act in list
 * order exactly like the reference's remove_if loop; the kernels find the particles whose outcome cannot depend on an
 * earlier, still undecided one and retire those together, round after round.  The 32 x 32 landing search
 * (world.cpp:2251-2283) indexes tiles[] without a bounds test in the reference; here positions outside the world are
 * skipped.  Not available on strip-sharded worlds. */
int fss_tick_particles(fss_world* w);
/* particles.push_back(...) by host code (tools, explosions: Game.cpp) */
int fss_particles_add(fss_world* w, const fss_particle* particles, size_t n);
int fss_particles_count(fss_world* w, size_t* n);
/* the list in order, for World::renderParticles (world.cpp:2145-2169) */
int fss_particles_download(fss_world* w, fss_particle* out, size_t cap, size_t* n_out);
int fss_particles_clear(fss_world* w);
/* rounds[0] = retire rounds of the last fss_tick_particles (1 when no particle depended on another), rounds[1] = cascade
 * passes over all its rounds */
int fss_particles_rounds(fss_world* w, uint64_t rounds[2]);
/* per-material cell counts over the owned rows (movingTiles replacement, Game.cpp:2579-2589) */
int fss_histogram(fss_world* w, uint64_t* counts, int n_materials);
/* order-independent digest (sum of fss_cell_digest over owned rows); k-GPU digests add up */
int fss_hash(fss_world* w, uint64_t* digest);
/* sum over owned rows of fluid_amount + fluid_amount_diff of SOUP cells, per material (double) */
int fss_fluid_totals(fss_world* w, double* totals, int n_materials);
int fss_sync(fss_world* w);

/* ---- synthetic worlds for benchmarks ----------------------------------
 * Fills every cell of this world's storage: cells within `border` of a world edge become
 * material `border_material`; every other cell draws key = fss_rng_cell_key(seed, x, y), r = key % 100
 * and takes the first entry i with r < cumulative_percent[i]; remaining cells are material 0.
 * Cells are initialised like Tiles::create(mat,x,y) with the table's create_* recipe. */
typedef struct fss_fill_spec {
    uint32_t seed;
    uint32_t border;
    int32_t border_material;
    int32_t n_entries;
    int32_t material[16];
    int32_t cumulative_percent[16];
    /* optional: rows >= solid_from_row are all `solid_material` (config 5, mostly settled), except inside the boxes below */
    uint32_t solid_from_row; /* 0 = off */
    int32_t solid_material;
    /* optional: entries apply only inside boxes; outside, sparse_material with probability
     * sparse_per_mille/1000 (config 5).  n_boxes == 0 = entries apply everywhere. */
    int32_t n_boxes;
    uint32_t box_size;
    int32_t sparse_material;
    uint32_t sparse_per_mille;
} fss_fill_spec;
int fss_fill_synthetic(fss_world* w, const fss_fill_spec* spec);

/* ---- row-strip sharding over NCCL (one process per GPU) --------------------------------------
 * rank r owns strip r; neighbours are r-1 (above, smaller y) and r+1 (below).  The unique id is
 * 128 bytes produced on rank 0 and distributed by the host (torch.distributed, MPI, a file...). */
int fss_comm_unique_id(void* id128);
int fss_comm_attach(fss_world* w, const void* id128, int rank, int n_ranks);
/* explicit halo exchange after sub-step tk (what fss_tick does internally: rows move after tk 1 and
 * tk 3 only, other tk are no-ops), and of the temperature plane after fss_tick_temperature */
int fss_exchange_halo(fss_world* w, int tk);
int fss_exchange_halo_temperature(fss_world* w);
/* the same two exchanges between two adjacent strip worlds of ONE process (same or different GPU),
 * by peer copies instead of NCCL: `upper` owns the rows just above `lower`.  The caller drives the
 * schedule with fss_substep(): for each iter, tk 0, tk 1, exchange(1), tk 2, tk 3, exchange(3). */
int fss_exchange_halo_peer(fss_world* upper, fss_world* lower, int tk);
int fss_exchange_halo_temperature_peer(fss_world* upper, fss_world* lower);

/* ---- introspection used by benchmarks -------------------------------------------------------- */
typedef struct fss_stats {
    uint64_t kernel_launches;   /* kernels of this library launched since creation */
    uint64_t substeps;          /* movement sub-steps executed */
    uint64_t storage_bytes;     /* device bytes held by the planes */
    uint32_t storage_pitch;     /* cells per stored row */
    uint32_t storage_rows;
    uint32_t storage_y0;        /* global row of the first stored row */
    uint32_t reserved;
    uint64_t spawns_refused;    /* spawn sites that found the record buffer (fss_config.particle_capacity) full: their cells stayed
                                 * in the grid (no matter is lost) instead of leaving as particles; counted when the host next
                                 * looks at the buffer (fss_drain_particles, fss_tick_particles) */
} fss_stats;
int fss_get_stats(fss_world* w, fss_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* FSS_H */
