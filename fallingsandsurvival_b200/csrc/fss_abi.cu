// fss_abi.cu -- the C ABI of include/fss.h on top of the kernels in fss_kernels.cuh.
//
// Host-side counterpart of the array part of `World` (FallingSandSurvival/world.hpp:98-146): owns the
// device planes, the material table, the tick zone, the particle spawn buffer and (for strip-sharded
// worlds) the NCCL communicator.  There is no CPU implementation of the path in this library: without
// a CUDA device fss_create() fails with FSS_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <cmath>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "fss_chunkio.h"
#include "fss_kernels.cuh"
#include "fss_particles.cuh"
#include "fss_render.cuh"

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if(e_ != cudaSuccess) return fail(FSS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while(0)

// ---- NCCL, bound lazily so that single-GPU users need no libnccl ------------------------------------------
struct NcclId {
    char internal[128];
};
typedef int (*nccl_get_unique_id_t)(NcclId*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclId, int);
typedef int (*nccl_comm_destroy_t)(void*);
typedef int (*nccl_send_t)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_recv_t)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_group_t)(void);
typedef const char* (*nccl_err_t)(int);
static struct {
    void* lib;
    nccl_get_unique_id_t get_unique_id;
    nccl_comm_init_rank_t comm_init_rank;
    nccl_comm_destroy_t comm_destroy;
    nccl_send_t send;
    nccl_recv_t recv;
    nccl_all_reduce_t all_reduce;
    nccl_group_t group_start, group_end;
    nccl_err_t err_string;
} g_nccl;
#define NCCL_UINT32 3 /* ncclUint32, nccl.h */
#define NCCL_UINT8 1  /* ncclUint8 */
#define NCCL_MAX 2    /* ncclMax */

static int nccl_bind() {
    if(g_nccl.lib) return FSS_OK;
    // RTLD_NOLOAD first: inside a torch process this picks up the libnccl torch already loaded
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if(!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if(!h) h = dlopen("libnccl.so", RTLD_NOW);
    if(!h) return fail(FSS_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.get_unique_id = (nccl_get_unique_id_t)dlsym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (nccl_comm_init_rank_t)dlsym(h, "ncclCommInitRank");
    g_nccl.comm_destroy = (nccl_comm_destroy_t)dlsym(h, "ncclCommDestroy");
    g_nccl.send = (nccl_send_t)dlsym(h, "ncclSend");
    g_nccl.recv = (nccl_recv_t)dlsym(h, "ncclRecv");
    g_nccl.all_reduce = (nccl_all_reduce_t)dlsym(h, "ncclAllReduce");
    g_nccl.group_start = (nccl_group_t)dlsym(h, "ncclGroupStart");
    g_nccl.group_end = (nccl_group_t)dlsym(h, "ncclGroupEnd");
    g_nccl.err_string = (nccl_err_t)dlsym(h, "ncclGetErrorString");
    if(!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.send || !g_nccl.recv || !g_nccl.all_reduce || !g_nccl.group_start || !g_nccl.group_end)
        return fail(FSS_ERR_COMM, "libnccl.so.2 lacks a required symbol");
    g_nccl.lib = h;
    return FSS_OK;
}
#define NC(call)                                                                                                   \
    do {                                                                                                           \
        int e_ = (call);                                                                                           \
        if(e_ != 0) return fail(FSS_ERR_COMM, "%s: %s", #call, g_nccl.err_string ? g_nccl.err_string(e_) : "nccl error"); \
    } while(0)

// ---- the world ------------------------------------------------------------------------------------------
struct fss_world {
    fss_config cfg;
    int device;
    cudaStream_t own_stream, stream;
    Grid g;
    int32_t* temp2;
    int own_lo, own_hi; // owned global rows
    bool sharded;
    int zx, zy, zw, zh;
    bool zone_set, mats_set;
    int n_mat, fire_id;
    uint32_t nothing_mf;
    DMat h_mats[FSS_MAX_MATERIALS];
    DMat* d_mats;
    DTex h_tex[FSS_MAX_TEXTURES];
    DTex* d_tex;
    fss_particle_spawn* d_parts;
    unsigned long long* d_part_count;
    uint32_t part_cap;
    void* d_stage;
    size_t stage_bytes;
    // fss_tick_host: its own staging buffer for the way back, the two copy streams, one event pair per band
    void* d_stage_down;
    cudaStream_t up_stream, down_stream;
    cudaEvent_t ev_up[32], ev_wave[32], ev_stage_free;
    unsigned long long* d_red; // FSS_MAX_MATERIALS u64 + FSS_MAX_MATERIALS f64
    fss_stats stats;
    // settled-chunk flags (SubstepParams::q): (q_h x q_w) bytes, chunk (ci, cj) of the zone at [(cj - q_cj0 + 1) * q_w + ci + 1]
    uint8_t* d_q;
    uint8_t* d_qskip; // FSS_TRACK_DIRTY: SubstepParams::qskip
    int q_w, q_h, q_cj0;
    // which materials can be in the grid (only ever grows): decides from which iter on the liquid-only kernel is exact
    bool present[FSS_MAX_MATERIALS];
    bool present_all;
    uint8_t* d_present;
    int liq_from_iter;    // from this iter on only liquids can act
    int nogas_from_iter;  // from this iter on no GAS / fire can act
    bool sand_reacts;     // a SAND material with reactions can be present (world.cpp:1251-1274): the general register-window kernel
    int gas_iters;        // largest `iterations` of a GAS material that can be present (0: none): GAS acts in scan 1 while iter < this
    int fire_iters;       // largest `iterations` of a fire material that can be present (0: none): fire acts while iter < this
    bool fire_static;     // fire can be split off into k_fire_prepass: no SAND material denser than a fire material, no SAND reaction with a SOLID product
    uint8_t* d_fire_verdict; // one byte per stored cell (allocated when the first such sub-step runs)
    bool one_liquid;      // at most one SOUP material can be present: two different liquids never meet (world.cpp:1406-1412, :1510-1516)
    // World::particles (fss_particles.cuh): the list, its scratch copy and the per-particle round state, all `pl_cap` long
    fss_particle *pl_list, *pl_next;
    uint8_t *pl_decided, *pl_keep, *pl_sure, *pl_big;
    uint32_t *pl_ran, *pl_listed;
    int4 *pl_act, *pl_seen, *pl_reach;
    uint32_t pl_n, pl_cap;
    uint32_t* pl_res_cell; // four hashed cell tables (W_act, R_act, W_pot, R_pot) of pl_res_mask + 1 entries each
    uint32_t pl_res_mask;
    uint32_t* pl_small;    // the tables of the cooperative-launch path for short lists
    int pl_coop_ctas;      // how many of its CTAs are resident at once
    uint32_t* pl_res_tile; // one entry per 16 x 16 tile of the world
    uint32_t* pl_ctl;      // [0] left undecided, [1] cutoff, [2] list length after compaction
    uint64_t *pl_key[2];   // spawn order keys / indices, `part_cap` long, for the sort
    uint32_t *pl_idx[2];
    void* pl_tmp; // cub scratch
    size_t pl_tmp_bytes;
    bool pl_used; // the caller keeps World::particles on the device (fss_tick_particles / fss_particles_add were called)
    uint64_t pl_rounds, pl_passes; // rounds / cascade passes of the last fss_tick_particles
    // dirty -> pixels (fss_render.cuh): the four RGBA layers, the flow map's smoothing state, Material::alpha / emitColor
    uint32_t* px[4];
    float *prev_fx, *prev_fy;
    uint8_t h_alpha[FSS_MAX_MATERIALS];
    uint32_t h_emit[FSS_MAX_MATERIALS];
    int32_t h_iterations[FSS_MAX_MATERIALS];
    DLook* d_looks;
    bool looks_stale;
    unsigned long long* d_rinfo; // FSS_MAX_MATERIALS counters, then hadDirty / hadFire / hadFlow as uint32
    // compacted worklist of not-settled tiles (mostly settled worlds): SubstepParams::work
    uint32_t* d_work;
    uint32_t* d_work_n;
    volatile uint32_t* h_work_n; // pinned: the latest count the device has reported (read without synchronising: a hint only)
    uint32_t* d_work_n_host;     // its device address
    int work_flip;               // which of the two counters d_work_n[0 / 1] the next list is counted in
    bool work_zeroed;            // ... and whether the last tick kernel has already zeroed it
    uint8_t *h_present_mask, *d_present_mask; // strip-sharded with a communicator: which materials ANY strip can hold (share_present)
    uint8_t* d_dirty_halo;       // strip-sharded + FSS_TRACK_DIRTY: the marks received with a halo exchange (2 rows of bytes)
    uint8_t* h_seen;             // fss_tick_host: pinned + mapped, 32 bands x FSS_MAX_MATERIALS bytes (which materials each band held)
    uint8_t* d_seen_host;
    size_t work_cap;
    uint32_t work_tiles;         // tiles of the launch that count belongs to
    bool work_mode;              // launches go through the worklist
    bool edge_launch;            // fss_tick is launching a strip's edge band on the exchange stream
    // strips
    cudaStream_t comm_stream; // the halo exchange of fss_tick runs here, beside the interior bands
    cudaEvent_t ev_edge, ev_halo;
    void* comm;
    int rank, n_ranks;
};

#define STAGE_BYTES ((size_t)256 << 20)

static int use_device(fss_world* w) {
    CU(cudaSetDevice(w->device));
    return FSS_OK;
}

static size_t plane_cells(const fss_world* w) { return (size_t)w->g.rows * w->g.pitch; }

// rows of 32-row bands of the zone owned by this strip (boundaries are multiples of 2*CHUNK_H from zone.y)
static void owned_bands(const fss_world* w, int& j_lo, int& j_hi) {
    j_lo = w->own_lo <= w->zy ? 0 : (w->own_lo - w->zy) / FSS_ZONE_ALIGN_Y;
    j_hi = w->own_hi >= w->zy + w->zh ? w->zh / FSS_ZONE_ALIGN_Y : (w->own_hi - w->zy) / FSS_ZONE_ALIGN_Y;
}

// From which iter on can only liquids act?  A SAND / GAS / fire material m acts while iter < iterations(m); once iter has
// passed the largest such count among the materials that can be present (initial content, uploads, and everything they
// can turn into: reaction targets, condensed steam, fire), the liquid-only kernel k_substep<true> gives identical results.
static void recompute_liq(fss_world* w) {
    bool pres[FSS_MAX_MATERIALS];
    // (a strip without a communicator cannot know what the other strips hold: everything; with one, share_present() tells it)
    const bool unknown = w->present_all || (w->sharded && !w->comm);
    for(int m = 0; m < FSS_MAX_MATERIALS; m++) pres[m] = m < w->n_mat && (unknown || w->present[m]);
    for(bool grew = true; grew;) {
        grew = false;
        for(int m = 0; m < w->n_mat; m++) {
            if(!pres[m]) continue;
            const DMat& d = w->h_mats[m];
            if(mf_type(d.bits) == FSS_PHYS_SAND) // only SAND evaluates its reactions (world.cpp:1251-1274)
                for(int r = 0; r < d.n_react; r++)
                    if(!pres[d.r_target[r]]) pres[d.r_target[r]] = grew = true;
            if((d.flags & FSS_MAT_IS_STEAM) && w->cfg.condense_material >= 0 && !pres[w->cfg.condense_material]) pres[w->cfg.condense_material] = grew = true;
        }
    }
    int from = 0;
    bool gas_or_fire = false;
    w->sand_reacts = false;
    w->fire_iters = 0;
    w->gas_iters = 0;
    int liquids = 0;
    for(int m = 0; m < w->n_mat; m++) {
        if(!pres[m]) continue;
        if(mf_type(w->h_mats[m].bits) == FSS_PHYS_SOUP) liquids++;
        if(w->h_mats[m].bits & MF_FIRE) w->fire_iters = std::max(w->fire_iters, std::min(mf_iters(w->h_mats[m].bits), 6));
        if(mf_type(w->h_mats[m].bits) == FSS_PHYS_GAS) w->gas_iters = std::max(w->gas_iters, std::min(mf_iters(w->h_mats[m].bits), 6));
        if(mf_type(w->h_mats[m].bits) == FSS_PHYS_SAND && w->h_mats[m].n_react > 0) w->sand_reacts = true;
        const int type = mf_type(w->h_mats[m].bits), iters = std::min(mf_iters(w->h_mats[m].bits), 6);
        const bool gf = type == FSS_PHYS_GAS || (w->h_mats[m].bits & MF_FIRE);
        gas_or_fire = gas_or_fire || gf;
        if(type == FSS_PHYS_SAND || gf) from = std::max(from, iters);
    }
    // While SAND moves, a GAS cell it displaces lands on an unvisited position and is then run by scans 2 and 3 whatever its
    // own `iterations` (they have no such gate): the no-gas kernel is only exact when no GAS / fire material is present at all.
    // The liquid-only kernel has no such case: nothing moves a cell there (SOUP-SOUP swaps at iter 0 aside).
    w->one_liquid = liquids <= 1;
    // fire_prepass_chunk's two conditions (fss_rules_liq.cuh), over the materials that can be present
    {
        uint32_t sand_rank = 0, fire_rank = 0xFFFFFFFFu;
        bool solid_product = false;
        for(int m = 0; m < w->n_mat; m++) {
            if(!pres[m]) continue;
            const DMat& d = w->h_mats[m];
            if(mf_type(d.bits) == FSS_PHYS_SAND) {
                sand_rank = std::max(sand_rank, d.bits & MF_RANK_MASK);
                for(int r = 0; r < d.n_react; r++)
                    if(mf_type(w->h_mats[d.r_target[r]].bits) == FSS_PHYS_SOLID) solid_product = true;
            }
            if(d.bits & MF_FIRE) fire_rank = std::min(fire_rank, d.bits & MF_RANK_MASK);
        }
        w->fire_static = sand_rank <= fire_rank && !solid_product;
    }
    w->nogas_from_iter = gas_or_fire ? 99 : 0;
    w->liq_from_iter = from;
    if(getenv("FSS_DEBUG")) {
        fprintf(stderr, "[fss] materials that can be present:");
        for(int m = 0; m < w->n_mat; m++)
            if(pres[m]) fprintf(stderr, " %d", m);
        fprintf(stderr, " -> liquid-only kernel from iter %d, no-gas kernel %s\n", w->liq_from_iter, gas_or_fire ? "never" : "for the other iters");
    }
}

// Forget the "settled" verdict of every chunk whose read window ([-2,+5] columns, [-2,+25] rows around the chunk)
// can see a cell of the rectangle (global cell coordinates; rw, rh < 0 = everything).
static int dirty_chunks(fss_world* w, int rx, int ry, int rw, int rh) {
    if(!w->d_q) return FSS_OK;
    int c0 = 0, c1 = w->q_w, r0 = 0, r1 = w->q_h;
    if(rw >= 0) {
        auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
        c0 = std::max(0, fdiv(rx - 5 - w->zx, FSS_CHUNK_W) + 1);
        c1 = std::min(w->q_w, fdiv(rx + rw - 1 + 2 - w->zx, FSS_CHUNK_W) + 2);
        r0 = std::max(0, fdiv(ry - 25 - w->zy, FSS_CHUNK_H) - w->q_cj0 + 1);
        r1 = std::min(w->q_h, fdiv(ry + rh - 1 + 2 - w->zy, FSS_CHUNK_H) - w->q_cj0 + 2);
        if(c0 >= c1 || r0 >= r1) return FSS_OK;
    }
    CU(cudaMemset2DAsync(w->d_q + (size_t)r0 * w->q_w + c0, (size_t)w->q_w, 0xFF, (size_t)(c1 - c0), (size_t)(r1 - r0), w->stream));
    return FSS_OK;
}

static void particles_free(fss_world* w) {
    void* ptrs[] = {w->pl_list, w->pl_next, w->pl_decided, w->pl_keep, w->pl_sure, w->pl_big, w->pl_ran, w->pl_listed, w->pl_act, w->pl_seen, w->pl_reach,
                    w->pl_res_cell, w->pl_res_tile, w->pl_small, w->pl_ctl, w->pl_key[0], w->pl_key[1], w->pl_idx[0], w->pl_idx[1], w->pl_tmp};
    for(void* p : ptrs) cudaFree(p);
}

extern "C" {

int fss_abi_version(void) { return FSS_ABI_VERSION; }
const char* fss_last_error(void) { return g_err.c_str(); }

int fss_create(const fss_config* cfg, fss_world** out) {
    if(!cfg || !out) return fail(FSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if(cfg->struct_size != sizeof(fss_config)) return fail(FSS_ERR_INVALID, "fss_config.struct_size %u != %zu", cfg->struct_size, sizeof(fss_config));
    if(cfg->width < 2 * FSS_ZONE_MARGIN + FSS_ZONE_ALIGN_X || cfg->height < 2 * FSS_ZONE_MARGIN + FSS_ZONE_ALIGN_Y)
        return fail(FSS_ERR_INVALID, "world %ux%u too small", cfg->width, cfg->height);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if(e != cudaSuccess || n_dev == 0) return fail(FSS_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if(cfg->device < 0 || cfg->device >= n_dev) return fail(FSS_ERR_NO_DEVICE, "device %d of %d", cfg->device, n_dev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if(prop.major < 10) return fail(FSS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

    fss_world* w = new fss_world();
    memset(w, 0, sizeof *w);
    w->cfg = *cfg;
    w->device = cfg->device;
    w->fire_id = -1;
    w->sharded = cfg->strip_rows != 0;
    if(w->sharded) {
        if(cfg->strip_y0 + cfg->strip_rows > cfg->height || cfg->strip_rows < 2 * FSS_HALO_ROWS) {
            delete w;
            return fail(FSS_ERR_INVALID, "strip [%u, +%u): outside height %u or shorter than %d rows", cfg->strip_y0, cfg->strip_rows, cfg->height, 2 * FSS_HALO_ROWS);
        }
        w->own_lo = (int)cfg->strip_y0;
        w->own_hi = (int)(cfg->strip_y0 + cfg->strip_rows);
    } else {
        w->own_lo = 0;
        w->own_hi = (int)cfg->height;
    }
    const int s_lo = std::max(0, w->own_lo - FSS_HALO_ROWS), s_hi = std::min((int)cfg->height, w->own_hi + FSS_HALO_ROWS);
    w->g.pitch = (cfg->width + FSS_BLOCK_COLS - 1) / FSS_BLOCK_COLS * FSS_BLOCK_COLS;
    w->g.sy0 = s_lo;
    w->g.rows = s_hi - s_lo;
    w->part_cap = cfg->particle_capacity ? cfg->particle_capacity : (1u << 20);
    if(sizeof(cidx_t) == 4 && (unsigned long long)w->g.rows * w->g.pitch >= (1ull << 32)) {
        delete w;
        return fail(FSS_ERR_INVALID, "%d stored rows x pitch %u is 2^32 cells or more: shard the world into more strips", s_hi - s_lo, (cfg->width + FSS_BLOCK_COLS - 1) / FSS_BLOCK_COLS * FSS_BLOCK_COLS);
    }

#define CUX(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if(e_ != cudaSuccess) {                                                               \
            fail(FSS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                      \
            fss_destroy(w);                                                                   \
            return FSS_ERR_CUDA;                                                              \
        }                                                                                     \
    } while(0)
    CUX(cudaSetDevice(w->device));
    CUX(cudaStreamCreateWithFlags(&w->own_stream, cudaStreamNonBlocking));
    w->stream = w->own_stream;
    const size_t n = plane_cells(w);
    CUX(cudaMalloc(&w->g.mf, n * 4));
    CUX(cudaMalloc(&w->g.col, n * 4));
    CUX(cudaMalloc(&w->g.temp, n * 4));
    CUX(cudaMalloc(&w->temp2, n * 4));
    CUX(cudaMalloc(&w->g.amt, n * 4));
    CUX(cudaMalloc(&w->g.dif, n * 4));
    CUX(cudaMemsetAsync(w->g.mf, 0, n * 4, w->stream));
    CUX(cudaMemsetAsync(w->g.col, 0, n * 4, w->stream));
    CUX(cudaMemsetAsync(w->g.temp, 0, n * 4, w->stream));
    CUX(cudaMemsetAsync(w->temp2, 0, n * 4, w->stream));
    CUX(cudaMemsetAsync(w->g.amt, 0, n * 4, w->stream));
    CUX(cudaMemsetAsync(w->g.dif, 0, n * 4, w->stream));
    if(cfg->flags & FSS_TRACK_FLOW) {
        CUX(cudaMalloc(&w->g.fx, n * 4));
        CUX(cudaMalloc(&w->g.fy, n * 4));
        CUX(cudaMemsetAsync(w->g.fx, 0, n * 4, w->stream));
        CUX(cudaMemsetAsync(w->g.fy, 0, n * 4, w->stream));
    }
    if(cfg->flags & FSS_TRACK_DIRTY) {
        // (strip-sharded worlds: the marks a strip leaves in its ghost rows travel with the halo exchange, dirty_halo_*)
        CUX(cudaMalloc(&w->g.dirty, n));
        CUX(cudaMemsetAsync(w->g.dirty, 0, n, w->stream));
    }
    for(int m = 0; m < FSS_MAX_MATERIALS; m++) w->h_alpha[m] = 255;
    CUX(cudaMalloc(&w->d_mats, sizeof(DMat) * FSS_MAX_MATERIALS));
    CUX(cudaMalloc(&w->d_tex, sizeof(DTex) * FSS_MAX_TEXTURES));
    CUX(cudaMemsetAsync(w->d_tex, 0, sizeof(DTex) * FSS_MAX_TEXTURES, w->stream));
    CUX(cudaMalloc(&w->d_parts, sizeof(fss_particle_spawn) * (size_t)w->part_cap));
    CUX(cudaMalloc(&w->d_part_count, 16)); // [0] records written, [1] spawns refused because the buffer was full
    CUX(cudaMemsetAsync(w->d_part_count, 0, 16, w->stream));
    CUX(cudaMalloc(&w->d_red, FSS_MAX_MATERIALS * 16));
    CUX(cudaMalloc(&w->d_present, FSS_MAX_MATERIALS));
    CUX(cudaMemsetAsync(w->d_present, 0, FSS_MAX_MATERIALS, w->stream));
    w->present[0] = true; // Tiles::NOTHING
    w->stage_bytes = std::min(STAGE_BYTES, std::max((size_t)1 << 20, (size_t)cfg->width * sizeof(fss_cell) * (size_t)w->g.rows));
    CUX(cudaMalloc(&w->d_stage, w->stage_bytes));
    CUX(cudaStreamSynchronize(w->stream));
#undef CUX
    w->stats.storage_bytes = n * 4 * 6;
    w->stats.storage_pitch = w->g.pitch;
    w->stats.storage_rows = (uint32_t)w->g.rows;
    w->stats.storage_y0 = (uint32_t)w->g.sy0;
    *out = w;
    return FSS_OK;
}

int fss_destroy(fss_world* w) {
    if(!w) return FSS_OK;
    cudaSetDevice(w->device);
    if(w->own_stream) cudaStreamSynchronize(w->own_stream);
    if(w->comm_stream) cudaStreamSynchronize(w->comm_stream);
    if(w->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(w->comm);
    if(w->comm_stream) cudaStreamDestroy(w->comm_stream);
    if(w->ev_edge) cudaEventDestroy(w->ev_edge);
    if(w->ev_halo) cudaEventDestroy(w->ev_halo);
    cudaFree(w->g.mf);
    cudaFree(w->g.col);
    cudaFree(w->g.temp);
    cudaFree(w->temp2);
    cudaFree(w->g.amt);
    cudaFree(w->g.dif);
    cudaFree(w->g.fx);
    cudaFree(w->g.dirty);
    for(int k = 0; k < 4; k++) cudaFree(w->px[k]);
    cudaFree(w->prev_fx);
    cudaFree(w->prev_fy);
    cudaFree(w->d_looks);
    cudaFree(w->d_rinfo);
    cudaFree(w->g.fy);
    cudaFree(w->d_mats);
    for(int i = 0; i < FSS_MAX_TEXTURES; i++) cudaFree((void*)w->h_tex[i].px);
    cudaFree(w->d_tex);
    cudaFree(w->d_parts);
    cudaFree(w->d_part_count);
    cudaFree(w->d_red);
    cudaFree(w->d_stage);
    cudaFree(w->d_stage_down);
    if(w->up_stream) cudaStreamDestroy(w->up_stream);
    if(w->down_stream) cudaStreamDestroy(w->down_stream);
    for(int k = 0; k < 32; k++) {
        if(w->ev_up[k]) cudaEventDestroy(w->ev_up[k]);
        if(w->ev_wave[k]) cudaEventDestroy(w->ev_wave[k]);
    }
    if(w->ev_stage_free) cudaEventDestroy(w->ev_stage_free);
    cudaFree(w->d_q);
    cudaFree(w->d_qskip);
    cudaFree(w->d_present);
    cudaFree(w->d_work);
    cudaFree(w->d_work_n);
    if(w->h_work_n) cudaFreeHost((void*)w->h_work_n);
    if(w->h_seen) cudaFreeHost(w->h_seen);
    cudaFree(w->d_dirty_halo);
    cudaFree(w->d_fire_verdict);
    cudaFree(w->d_present_mask);
    if(w->h_present_mask) cudaFreeHost(w->h_present_mask);
    particles_free(w);
    if(w->own_stream) cudaStreamDestroy(w->own_stream);
    delete w;
    return FSS_OK;
}

int fss_set_stream(fss_world* w, void* cuda_stream) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(use_device(w)) return FSS_ERR_CUDA;
    CU(cudaStreamSynchronize(w->stream));
    w->stream = cuda_stream ? (cudaStream_t)cuda_stream : w->own_stream;
    return FSS_OK;
}

int fss_sync(fss_world* w) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(use_device(w)) return FSS_ERR_CUDA;
    CU(cudaStreamSynchronize(w->stream));
    CU(cudaGetLastError());
    return FSS_OK;
}

// ---- material table -----------------------------------------------------------------------------------
int fss_set_materials(fss_world* w, const fss_material* t, int n) {
    if(!w || !t) return fail(FSS_ERR_INVALID, "null argument");
    if(n < 1 || n > FSS_MAX_MATERIALS) return fail(FSS_ERR_INVALID, "%d materials (1..%d)", n, FSS_MAX_MATERIALS);
    if(use_device(w)) return FSS_ERR_CUDA;
    int fire_id = -1;
    DMat tmp[FSS_MAX_MATERIALS];
    memset(tmp, 0, sizeof tmp);
    for(int i = 0; i < n; i++) {
        const fss_material& m = t[i];
        DMat& d = tmp[i];
        if(m.physics_type < 0 || m.physics_type > FSS_PHYS_PASSABLE) return fail(FSS_ERR_INVALID, "material %d: physics type %d", i, m.physics_type);
        if(m.physics_type == FSS_PHYS_SAND && (m.slipperyness < 1 || m.slipperyness > 0xFFFF))
            return fail(FSS_ERR_INVALID, "material %d: SAND needs 1 <= slipperyness <= 65535 (rand() %% (2*slip), world.cpp:1736)", i);
        if((m.flags & FSS_MAT_IS_FIRE) && m.physics_type != FSS_PHYS_PASSABLE)
            return fail(FSS_ERR_INVALID, "material %d: the fire material must be PASSABLE like Materials::FIRE (Materials.cpp:42)", i);
        if(m.n_reactions < 0 || m.n_reactions > FSS_MAX_REACTIONS) return fail(FSS_ERR_INVALID, "material %d: %d reactions", i, m.n_reactions);
        if(m.create_kind == FSS_CREATE_RAND && m.create_rand_mod == 0) return fail(FSS_ERR_INVALID, "material %d: create_rand_mod == 0", i);
        if(!std::isfinite(m.conduction_other) || !std::isfinite(m.conduction_self))
            return fail(FSS_ERR_INVALID, "material %d: conductionSelf / conductionOther must be finite (world.cpp:2072-2132)", i);
        if(m.create_kind == FSS_CREATE_TEXTURE && m.create_texture >= FSS_MAX_TEXTURES) return fail(FSS_ERR_INVALID, "material %d: texture %d", i, m.create_texture);
        d.density = m.density;
        const int iters = m.iterations < 0 ? 0 : (m.iterations > 7 ? 7 : m.iterations); // iter < 6 always: clamping is exact
        d.bits = ((uint32_t)m.physics_type << MF_TYPE_SHIFT) | ((uint32_t)iters << MF_ITERS_SHIFT) | ((m.flags & FSS_MAT_IS_FIRE) ? MF_FIRE : 0u);
        d.flags = m.flags;
        d.slip = m.slipperyness;
        // `int maxStability = 8 / sqrt(slipperyness) + 1;` world.cpp:1710: double arithmetic, truncated
        d.max_stability = m.slipperyness > 0 ? (int)(8 / sqrt((double)m.slipperyness) + 1) : 255;
        d.n_react = m.n_reactions;
        for(int r = 0; r < m.n_reactions; r++) {
            if(m.reactions[r].data2 < 0 || m.reactions[r].data2 >= n) return fail(FSS_ERR_INVALID, "material %d: reaction target %d", i, m.reactions[r].data2);
            d.r_type[r] = m.reactions[r].type;
            d.r_thr[r] = m.reactions[r].data1;
            d.r_target[r] = m.reactions[r].data2;
        }
        d.create_kind = m.create_kind;
        d.create_base = m.create_base;
        d.create_mod = m.create_rand_mod;
        d.create_shift = m.create_rand_shift;
        d.create_site = m.create_rand_site;
        d.create_tex = m.create_texture;
        d.create_temp = m.create_temperature;
        d.add_temp = m.add_temp;
        d.cond_self = m.conduction_self;
        d.cond_other = m.conduction_other;
        if((m.flags & FSS_MAT_IS_FIRE) && fire_id < 0) fire_id = i;
    }
    if(w->cfg.condense_material >= n) return fail(FSS_ERR_INVALID, "condense_material %d >= %d materials", w->cfg.condense_material, n);
    { // density ranks for `belowTile.mat->density < tile.mat->density` (world.cpp:1276-1285): equal densities, equal rank
        std::vector<float> ds;
        for(int i = 0; i < n; i++) ds.push_back(t[i].density);
        std::sort(ds.begin(), ds.end());
        ds.erase(std::unique(ds.begin(), ds.end()), ds.end());
        for(int i = 0; i < n; i++) {
            if(t[i].density != t[i].density) return fail(FSS_ERR_INVALID, "material %d: density is NaN", i);
            tmp[i].bits |= (uint32_t)(std::lower_bound(ds.begin(), ds.end(), t[i].density) - ds.begin()) << MF_RANK_SHIFT;
        }
    }
    memcpy(w->h_mats, tmp, sizeof tmp);
    for(int i = 0; i < n; i++) w->h_iterations[i] = t[i].iterations; // the flow map divides by the real count (Game.cpp:2636)
    w->looks_stale = true;
    w->n_mat = n;
    w->fire_id = fire_id;
    w->nothing_mf = 0u | tmp[0].bits;
    CU(cudaMemcpyAsync(w->d_mats, w->h_mats, sizeof(DMat) * FSS_MAX_MATERIALS, cudaMemcpyHostToDevice, w->stream));
    k_rederive<<<148 * 8, 256, 0, w->stream>>>(w->g, w->d_mats, n, plane_cells(w));
    w->stats.kernel_launches++;
    CU(cudaStreamSynchronize(w->stream)); // h_mats may be rewritten by the next call
    w->mats_set = true;
    recompute_liq(w);
    if(dirty_chunks(w, 0, 0, -1, -1)) return FSS_ERR_CUDA;
    return FSS_OK;
}

int fss_set_texture(fss_world* w, int index, const uint32_t* argb, int tw, int th) {
    if(!w || !argb) return fail(FSS_ERR_INVALID, "null argument");
    if(index < 0 || index >= FSS_MAX_TEXTURES || tw < 1 || th < 1) return fail(FSS_ERR_INVALID, "texture %d (%dx%d)", index, tw, th);
    if(use_device(w)) return FSS_ERR_CUDA;
    CU(cudaStreamSynchronize(w->stream));
    // the replacement is complete on the device before the table entry changes; the old pixels are freed last
    uint32_t* d = nullptr;
    CU(cudaMalloc(&d, (size_t)tw * th * 4));
    cudaError_t e = cudaMemcpy(d, argb, (size_t)tw * th * 4, cudaMemcpyHostToDevice);
    if(e != cudaSuccess) {
        cudaFree(d);
        return fail(FSS_ERR_CUDA, "texture upload: %s", cudaGetErrorString(e));
    }
    const DTex old = w->h_tex[index];
    w->h_tex[index].px = d;
    w->h_tex[index].w = tw;
    w->h_tex[index].h = th;
    e = cudaMemcpy(w->d_tex, w->h_tex, sizeof(DTex) * FSS_MAX_TEXTURES, cudaMemcpyHostToDevice);
    if(e != cudaSuccess) { // the device table still names the old pixels: keep them
        w->h_tex[index] = old;
        cudaFree(d);
        return fail(FSS_ERR_CUDA, "texture table upload: %s", cudaGetErrorString(e));
    }
    cudaFree((void*)old.px);
    return FSS_OK;
}

static int skipped_marks(fss_world* w);
static int exchange_halo_on(fss_world* w, int tk, cudaStream_t st);
static int halo_received(fss_world* w, int tk);

int fss_set_tick_zone(fss_world* w, int x, int y, int zw, int zh) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(x % FSS_ZONE_ALIGN_X || zw % FSS_ZONE_ALIGN_X || zh % FSS_ZONE_ALIGN_Y || zw <= 0 || zh <= 0)
        return fail(FSS_ERR_ALIGN, "tick zone (%d,%d,%d,%d): x and width must be multiples of %d, height of %d", x, y, zw, zh, FSS_ZONE_ALIGN_X, FSS_ZONE_ALIGN_Y);
    if(x < FSS_ZONE_MARGIN || y < FSS_ZONE_MARGIN || x + zw > (int)w->cfg.width - FSS_ZONE_MARGIN || y + zh > (int)w->cfg.height - FSS_ZONE_MARGIN)
        return fail(FSS_ERR_ALIGN, "tick zone (%d,%d,%d,%d) must stay %d cells inside the %ux%u world", x, y, zw, zh, FSS_ZONE_MARGIN, w->cfg.width, w->cfg.height);
    if(w->sharded) {
        // interior strip boundaries must fall on the phase pattern (multiples of 2*CHUNK_H from the zone origin)
        const bool first = w->own_lo == 0, last = w->own_hi == (int)w->cfg.height;
        if(!first && ((w->own_lo - y) % FSS_ZONE_ALIGN_Y || w->own_lo < y || w->own_lo > y + zh))
            return fail(FSS_ERR_ALIGN, "strip start %d is not zone.y + k*%d inside the zone", w->own_lo, FSS_ZONE_ALIGN_Y);
        if(!last && ((w->own_hi - y) % FSS_ZONE_ALIGN_Y || w->own_hi < y || w->own_hi > y + zh))
            return fail(FSS_ERR_ALIGN, "strip end %d is not zone.y + k*%d inside the zone", w->own_hi, FSS_ZONE_ALIGN_Y);
    }
    if(use_device(w)) return FSS_ERR_CUDA;
    if(int rc = skipped_marks(w)) return rc; // notes taken under the old zone
    w->zx = x;
    w->zy = y;
    w->zw = zw;
    w->zh = zh;
    w->zone_set = true;
    CU(cudaStreamSynchronize(w->stream));
    cudaFree(w->d_q);
    w->d_q = nullptr;
    cudaFree(w->d_qskip);
    w->d_qskip = nullptr;
    if(!(w->cfg.flags & FSS_NO_EARLY_OUT)) {
        int j_lo, j_hi;
        owned_bands(w, j_lo, j_hi);
        w->q_w = zw / FSS_CHUNK_W + 2;
        w->q_h = std::max(0, j_hi - j_lo) * 2 + 2;
        w->q_cj0 = j_lo * 2;
        CU(cudaMalloc(&w->d_q, (size_t)w->q_w * w->q_h));
        if(w->g.dirty) {
            CU(cudaMalloc(&w->d_qskip, (size_t)w->q_w * w->q_h));
            CU(cudaMemsetAsync(w->d_qskip, 0xFF, (size_t)w->q_w * w->q_h, w->stream));
        }
        if(dirty_chunks(w, 0, 0, -1, -1)) return FSS_ERR_CUDA;
    }
    return FSS_OK;
}

// ---- host mirror <-> device ------------------------------------------------------------------------------
static int rect_clip(fss_world* w, int& x, int& y, int& rw, int& rh, size_t& skip_rows) {
    if(rw <= 0 || rh <= 0 || x < 0 || y < 0 || (uint32_t)(x + rw) > w->cfg.width || (uint32_t)(y + rh) > w->cfg.height)
        return fail(FSS_ERR_RANGE, "rectangle (%d,%d,%d,%d) outside the %ux%u world", x, y, rw, rh, w->cfg.width, w->cfg.height);
    const int lo = std::max(y, w->g.sy0), hi = std::min(y + rh, w->g.sy0 + w->g.rows);
    if(lo >= hi) return fail(FSS_ERR_RANGE, "rows [%d,%d) are not stored by this strip [%d,%d)", y, y + rh, w->g.sy0, w->g.sy0 + w->g.rows);
    skip_rows = (size_t)(lo - y);
    y = lo;
    rh = hi - lo;
    return FSS_OK;
}

static int rect_copy(fss_world* w, int x, int y, int rw, int rh, fss_cell* host, size_t pitch_cells, bool upload) {
    if(!w || !host) return fail(FSS_ERR_INVALID, "null argument");
    if(upload && !w->mats_set) return fail(FSS_ERR_STATE, "fss_upload_rect before fss_set_materials");
    if(pitch_cells < (size_t)rw) return fail(FSS_ERR_INVALID, "pitch %zu < width %d", pitch_cells, rw);
    if(use_device(w)) return FSS_ERR_CUDA;
    size_t skip = 0;
    int rc = rect_clip(w, x, y, rw, rh, skip);
    if(rc) return rc;
    host += skip * pitch_cells;
    const int rows_per_pass = (int)std::min<size_t>(65535, std::max<size_t>(1, w->stage_bytes / ((size_t)rw * sizeof(fss_cell)))); // gridDim.y
    const int nblk = ((x + rw - 1) >> 8) - (x >> 8) + 1;
    for(int r0 = 0; r0 < rh; r0 += rows_per_pass) {
        const int nr = std::min(rows_per_pass, rh - r0);
        RectParams p;
        p.g = w->g;
        p.mats = w->d_mats;
        p.n_mat = w->n_mat;
        p.aos = (uint32_t*)w->d_stage;
        p.present = w->d_present;
        p.rx = x;
        p.ry = y + r0;
        p.rw = rw;
        p.rh = nr;
        fss_cell* h = host + (size_t)r0 * pitch_cells;
        if(upload) {
            CU(cudaMemcpy2DAsync(w->d_stage, (size_t)rw * sizeof(fss_cell), h, pitch_cells * sizeof(fss_cell), (size_t)rw * sizeof(fss_cell), nr,
                                 cudaMemcpyHostToDevice, w->stream));
            k_upload<<<dim3(nblk, nr), 256, 0, w->stream>>>(p);
        } else {
            k_download<<<dim3(nblk, nr), 256, 0, w->stream>>>(p);
            CU(cudaMemcpy2DAsync(h, pitch_cells * sizeof(fss_cell), w->d_stage, (size_t)rw * sizeof(fss_cell), (size_t)rw * sizeof(fss_cell), nr,
                                 cudaMemcpyDeviceToHost, w->stream));
        }
        w->stats.kernel_launches++;
        CU(cudaGetLastError());
    }
    if(upload && dirty_chunks(w, x, y, rw, rh)) return FSS_ERR_CUDA;
    uint8_t seen[FSS_MAX_MATERIALS];
    if(upload) CU(cudaMemcpyAsync(seen, w->d_present, FSS_MAX_MATERIALS, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream)); // the caller owns `host` again when we return
    if(upload) {
        for(int m = 0; m < FSS_MAX_MATERIALS; m++) w->present[m] = w->present[m] || seen[m];
        recompute_liq(w);
    }
    return FSS_OK;
}

int fss_upload_rect(fss_world* w, int x, int y, int width, int height, const fss_cell* host, size_t pitch_cells) {
    return rect_copy(w, x, y, width, height, (fss_cell*)host, pitch_cells, true);
}
int fss_download_rect(fss_world* w, int x, int y, int width, int height, fss_cell* host, size_t pitch_cells) {
    return rect_copy(w, x, y, width, height, host, pitch_cells, false);
}

// ---- the hot path -----------------------------------------------------------------------------------------
static int ready(fss_world* w) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(!w->mats_set) return fail(FSS_ERR_STATE, "call fss_set_materials first");
    if(!w->zone_set) return fail(FSS_ERR_STATE, "call fss_set_tick_zone first");
    return use_device(w);
}

// jr_lo / jr_hi >= 0: only the 32-row bands [jr_lo, jr_hi) of the zone (fss_tick_host's waves), else all the owned ones
static int launch_substep(fss_world* w, uint32_t tick_ct, int iter, int tk, int jr_lo = -1, int jr_hi = -1) {
    // world.cpp:1115-1141: phase tk activates chunks at zone + (tk%2, 1 - tk/2) * chunk + 2*chunk*(i, j)
    const int ofs_x = tk % 2, ofs_y = 1 - tk / 2;
    // rows of active chunks owned by this strip (boundaries are multiples of 2*CHUNK_H from zone.y)
    int j_lo, j_hi;
    owned_bands(w, j_lo, j_hi);
    const int own_j_lo = j_lo;
    if(jr_lo >= 0) {
        j_lo = std::max(j_lo, jr_lo);
        j_hi = std::min(j_hi, jr_hi);
    }
    SubstepParams p;
    p.g = w->g;
    p.mats = w->d_mats;
    p.tex = w->d_tex;
    p.n_mat = w->n_mat;
    p.x_begin = w->zx + ofs_x * FSS_CHUNK_W;
    p.y_begin = w->zy + ofs_y * FSS_CHUNK_H + j_lo * FSS_ZONE_ALIGN_Y;
    p.nx = w->zw / FSS_ZONE_ALIGN_X;
    p.ny = j_hi - j_lo;
    p.tkey = fss_rng_tick_key(w->cfg.seed, tick_ct, (uint32_t)iter);
    p.iter = iter;
    p.tick_iter = tick_ct * 8u + (uint32_t)iter;
    p.flags = w->cfg.flags;
    p.condense = w->cfg.condense_material;
    p.fire_id = w->fire_id;
    p.nothing_mf = w->nothing_mf;
    p.parts = w->d_parts;
    p.part_count = w->d_part_count;
    p.part_cap = w->part_cap;
    // (the kernels index the flags by blockIdx.y: start them at this launch's first band)
    const size_t q_ofs = (size_t)std::max(0, 2 * (j_lo - own_j_lo)) * (size_t)w->q_w;
    p.q = w->d_q ? w->d_q + q_ofs : nullptr;
    p.qskip = w->d_qskip ? w->d_qskip + q_ofs : nullptr;
    p.q_w = w->q_w;
    p.ofs_x = ofs_x;
    p.ofs_y = ofs_y;
    if(jr_lo < 0) w->stats.substeps++;
    if(p.ny <= 0 || p.nx <= 0) return FSS_OK;
    {
        dim3 grid((p.nx + K1_THREADS - 1) / K1_THREADS, p.ny);
        uint32_t work_tiles_now = 0;
        p.work = nullptr;
        p.work_n = nullptr;
        p.fire_verdict = w->d_fire_verdict;
        // Mostly settled worlds: a small kernel lists the tiles that still hold a not-settled chunk and a fixed grid strides over
        // that list, so settled regions cost neither CTAs nor flag loads.  Whether a launch goes through the list is decided from
        // the last count the device reported (a hint read without synchronising: both ways give the same cells); while launches are
        // plain, the count is refreshed once per tick.  (FSS_TRACK_DIRTY needs a note per skipped chunk: plain launches only.)
        static const bool no_worklist = getenv("FSS_NO_WORKLIST") != nullptr;
        static const bool always_worklist = getenv("FSS_WORKLIST_ALWAYS") != nullptr; // tests: every launch through the list
        const uint32_t tiles_x = (uint32_t)(p.nx + 31) / 32; // warp tiles per chunk row
        if(w->d_q && !w->d_qskip && !no_worklist && !w->edge_launch && tiles_x <= 0x3FFFu && grid.y <= 0x3FFFFu) {
            const uint32_t tiles = tiles_x * grid.y;
            if(tiles > w->work_cap) {
                CU(cudaStreamSynchronize(w->stream));
                cudaFree(w->d_work);
                w->d_work = nullptr;
                w->work_cap = 0;
                CU(cudaMalloc(&w->d_work, (size_t)tiles * sizeof(uint32_t)));
                w->work_cap = tiles;
                if(!w->d_work_n) {
                    CU(cudaMalloc(&w->d_work_n, 2 * sizeof(uint32_t)));
                    CU(cudaMemsetAsync(w->d_work_n, 0, 2 * sizeof(uint32_t), w->stream));
                    // mapped: the count is written by a one-thread kernel, not by a copy (a cudaMemcpyAsync would queue behind the
                    // 256 MB passes of fss_tick_host's download on the device -> host copy engine and stall this stream)
                    CU(cudaHostAlloc((void**)&w->h_work_n, sizeof(uint32_t), cudaHostAllocMapped));
                    CU(cudaHostGetDevicePointer((void**)&w->d_work_n_host, (void*)w->h_work_n, 0));
                    *w->h_work_n = 0xFFFFFFFFu;
                }
            }
            if(w->work_tiles && *w->h_work_n != 0xFFFFFFFFu) w->work_mode = (double)*w->h_work_n < 0.5 * (double)w->work_tiles;
            if(always_worklist) w->work_mode = true;
            if(w->work_mode || (iter == 0 && tk == 0)) {
                // two counters take turns: the tick kernel that strides over list i zeroes the counter of list i+1 and reports its
                // own count to the host, so a launch through the list is two launches (list, tick kernel), not four
                uint32_t* cnt = w->d_work_n + w->work_flip;
                if(!w->work_zeroed) CU(cudaMemsetAsync(cnt, 0, sizeof(uint32_t), w->stream));
                k_worklist<<<(tiles + 127) / 128, 128, 0, w->stream>>>(p.q, p.q_w, ofs_x, ofs_y, iter, p.nx, (int)tiles_x, p.ny, w->d_work, cnt);
                w->work_tiles = tiles;
                w->stats.kernel_launches++;
                if(w->work_mode) {
                    p.work = w->d_work;
                    p.work_n = cnt;
                    p.work_n_host = w->d_work_n_host;
                    p.work_n_next = w->d_work_n + (w->work_flip ^ 1);
                    w->work_flip ^= 1;
                    w->work_zeroed = true;
                    work_tiles_now = tiles; // the grid is sized below, once the kernel (and how many of its CTAs fit an SM) is known
                } else { // a plain launch follows: the count only refreshes the hint (once per tick)
                    k_copy_words<<<1, 32, 0, w->stream>>>(cnt, w->d_work_n_host, 1);
                    w->work_zeroed = false;
                }
            }
        }
        static const bool no_spec = getenv("FSS_NO_SPECIALISE") != nullptr; // measurements: always the general kernel
        const int mode = no_spec ? K1_GENERAL : (iter >= w->liq_from_iter ? K1_LIQUID : (iter >= w->nogas_from_iter ? K1_NOGAS : K1_GENERAL));
        const bool flow = w->g.fx != nullptr;
        const bool dirty = w->g.dirty != nullptr;
        // with a worklist: exactly as many CTAs as are resident at once (a second wave would start when the first is nearly done)
        auto persistent = [&](const void* kernel) -> dim3 {
            if(!work_tiles_now) return grid;
            static std::vector<std::pair<const void*, int>> per_sm; // per kernel: CTAs per SM
            int n = 0;
            for(auto& e : per_sm)
                if(e.first == kernel) n = e.second;
            if(!n) {
                if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, K1_THREADS, 0) != cudaSuccess || n < 1) n = 4;
                per_sm.push_back({kernel, n});
            }
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
            return dim3(std::min<uint32_t>((work_tiles_now + K1_WARPS - 1) / K1_WARPS, (uint32_t)(sms * n)), 1);
        };
#define FSS_LAUNCH_K1(M, F, D) k_substep<M, F, D><<<persistent((const void*)k_substep<M, F, D>), K1_THREADS, 0, w->stream>>>(p)
#define FSS_LAUNCH_K1_MODE(M)                                                                      \
    do {                                                                                           \
        if(dirty) {                                                                                \
            if(flow) FSS_LAUNCH_K1(M, true, true); else FSS_LAUNCH_K1(M, false, true);             \
        } else {                                                                                   \
            if(flow) FSS_LAUNCH_K1(M, true, false); else FSS_LAUNCH_K1(M, false, false);           \
        }                                                                                          \
    } while(0)
        // (FSS_NO_WINDOW, read at every launch: the lock-step kernels for everything -- the cross-check of the two kernel families at
        // sizes the oracle does not reach, tests/test_gpu_parity.py::test_full_size_kernel_families_agree)
        if(liq_window_applies(mode, iter, w->cfg.flags, flow, dirty, w->sand_reacts, w->fire_iters, w->fire_static) && !getenv("FSS_NO_WINDOW")) {
#define FSS_LAUNCH_WIN(...) k_substep_win<__VA_ARGS__><<<persistent((const void*)k_substep_win<__VA_ARGS__>), K1_THREADS, 0, w->stream>>>(p)
            // the kernels with the iter-0 code (two different liquids changing places) are a superset and exact for every iter, but
            // 7 KB longer, and K1's speed follows its code size (same sub-step, 56 KB against 63 KB of code: 7.78 against 8.59 ms):
            // they run at iter 0 only, and not at all while at most one liquid material can be present
            static const bool force_iter0 = getenv("FSS_FORCE_ITER0_KERNEL") != nullptr; // measurements
            const bool swaps = (iter == 0 && !w->one_liquid) || force_iter0;
            const bool fire_acts = mode == K1_GENERAL && iter < w->fire_iters; // (only here with fire_static: liq_window_applies)
            if(fire_acts) { // the fire cells' visits are decided first (fire_prepass_chunk), the window kernel applies the verdicts
                if(!w->d_fire_verdict) {
                    CU(cudaMalloc(&w->d_fire_verdict, plane_cells(w)));
                    CU(cudaMemsetAsync(w->d_fire_verdict, 0, plane_cells(w), w->stream));
                }
                p.fire_verdict = w->d_fire_verdict;
                k_fire_prepass<<<dim3((p.nx + 127) / 128, p.ny), 128, 0, w->stream>>>(p);
                w->stats.kernel_launches++;
            }
            if(mode == K1_LIQUID) {
                if(swaps) FSS_LAUNCH_WIN(false, true);
                else FSS_LAUNCH_WIN(false, false);
            } else if(liq_window_general(mode, w->sand_reacts)) {
                // GAS / SAND reactions can be present (and fire that cannot act any more): the general window walk
                if(swaps || fire_acts) FSS_LAUNCH_WIN(true, true, 2);
                else if(iter < w->gas_iters) FSS_LAUNCH_WIN(true, false, 2);
                else FSS_LAUNCH_WIN(true, false, 1); // no GAS material acts in scan 1 any more: the shorter kernel
            } else {
                if(swaps) FSS_LAUNCH_WIN(true, true);
                else FSS_LAUNCH_WIN(true, false);
            }
#undef FSS_LAUNCH_WIN
        } else if(mode == K1_LIQUID) FSS_LAUNCH_K1_MODE(K1_LIQUID);
        else if(mode == K1_NOGAS) FSS_LAUNCH_K1_MODE(K1_NOGAS);
        else FSS_LAUNCH_K1_MODE(K1_GENERAL);
#undef FSS_LAUNCH_K1_MODE
#undef FSS_LAUNCH_K1
    }
    w->stats.kernel_launches++;
    return FSS_OK;
}

static int share_present(fss_world* w);

int fss_substep(fss_world* w, uint32_t tick_ct, int iter, int tk) {
    int rc = ready(w);
    if(rc) return rc;
    if(iter < 0 || iter > 5 || tk < 0 || tk > 3) return fail(FSS_ERR_INVALID, "sub-step (iter %d, tk %d)", iter, tk);
    if(w->comm && iter == 0 && tk == 0 && (rc = share_present(w))) return rc; // (a caller that drives the schedule itself: once per tick)
    rc = launch_substep(w, tick_ct, iter, tk);
    if(rc) return rc;
    CU(cudaGetLastError());
    return FSS_OK;
}

int fss_tick(fss_world* w, uint32_t tick_ct) {
    int rc = ready(w);
    if(rc) return rc;
    if(w->sharded && !w->comm) return fail(FSS_ERR_STATE, "strip-sharded world without communicator: attach one, or drive fss_substep + fss_exchange_halo_peer");
    static const bool serial_halo = getenv("FSS_SERIAL_HALO") != nullptr; // measurements: exchange on the compute stream
    if(w->comm && (rc = share_present(w))) return rc;
    if(w->comm && !w->comm_stream) {
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU(cudaStreamCreateWithPriority(&w->comm_stream, cudaStreamNonBlocking, prio_hi));
        CU(cudaEventCreateWithFlags(&w->ev_edge, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w->ev_halo, cudaEventDisableTiming));
    }
    int j_lo, j_hi;
    owned_bands(w, j_lo, j_hi);
    for(int iter = 0; iter < 6; iter++) {
        for(int tk = 0; tk < 4; tk++) {
            const bool exchange = w->comm && (tk == 1 || tk == 3);
            if(!exchange || serial_halo || j_hi - j_lo < 3) {
                rc = launch_substep(w, tick_ct, iter, tk);
                if(rc) return rc;
                if(exchange && (rc = fss_exchange_halo(w, tk))) return rc;
                continue;
            }
            // The rows that cross a strip boundary after this sub-step are written by ONE 32-row band: the strip's last band
            // after tk 1 (its lower chunk row ran in tk 0 / 1: rows [B-2, B+2) go down), its first band after tk 3 (rows
            // [B-2, B+10) go up).  That band runs first; while the other bands run, the ncclSend / ncclRecv group is on its own
            // stream.  What is received lands in rows this sub-step does not touch (tk 1: the ghost rows and the first two
            // rows of the strip, which only the upper chunk rows of tk 2 / 3 look at; tk 3: the ghost rows below, ten of which
            // the last band's tk 0 / 1 chunks read): the next sub-step waits for it.
            // (a launch lasts as long as one chunk visit however few chunks it has, so the edge band does not go in front of
            // the others on the world's stream: it runs beside them on the high-priority exchange stream, followed there by
            // the NCCL group)
            const int edge = tk == 1 ? j_hi - 1 : j_lo;
            cudaStream_t main_stream = w->stream;
            CU(cudaEventRecord(w->ev_edge, main_stream)); // the sub-steps so far
            CU(cudaStreamWaitEvent(w->comm_stream, w->ev_edge, 0));
            w->stream = w->comm_stream;
            w->edge_launch = true; // (one band, beside the launch that may be using the worklist buffers: a plain launch)
            rc = launch_substep(w, tick_ct, iter, tk, edge, edge + 1);
            w->edge_launch = false;
            w->stream = main_stream;
            if(rc) return rc;
            if((rc = exchange_halo_on(w, tk, w->comm_stream))) return rc;
            CU(cudaEventRecord(w->ev_halo, w->comm_stream));
            if((rc = tk == 1 ? launch_substep(w, tick_ct, iter, tk, j_lo, edge) : launch_substep(w, tick_ct, iter, tk, edge + 1, j_hi))) return rc;
            CU(cudaStreamWaitEvent(main_stream, w->ev_halo, 0));
            if((rc = halo_received(w, tk))) return rc;
            w->stats.substeps++;
        }
    }
    CU(cudaGetLastError());
    return FSS_OK;
}

// ---- fss_tick_host: upload | 24 sub-steps | download, pipelined in row bands ----------------------------------------------
// Band b = stored rows [R_b, R_b+1), interior boundaries at zone.y + 32 * K_b.  After band b has arrived, "wave b" runs the 24
// sub-steps s = 0 .. 23 on the 32-row bands [F_b-1(s), F_b(s)) of the zone, F_b(s) = K_b - (s + 1) (last band: all the rest):
// a chunk visit reads at most 10 rows below its chunk and writes at most 2, so everything sub-step s reads inside
// [.., F_b(s)) was brought to the state after sub-step s-1 by this or an earlier wave and lies above row zone.y + 32 * K_b,
// i.e. in the bands that have arrived; and the chunks a later wave adds to sub-step s are, as in every sub-step, independent
// of the ones that already ran it (the checkerboard phase is race-free in any order).  Rows above F_b(23) - 1 are final
// after wave b and go back to the host while the next band is still on its way in.
#ifndef FSS_HOST_BAND_ROWS
#define FSS_HOST_BAND_ROWS 1024
#endif
static int host_copy_rows(fss_world* w, cudaStream_t st, void* stage, int y, int nrows, fss_cell* cells, size_t pitch_cells, bool upload) {
    const int width = (int)w->cfg.width;
    const int rows_per_pass = (int)std::min<size_t>(65535, std::max<size_t>(1, w->stage_bytes / ((size_t)width * sizeof(fss_cell))));
    const int nblk = ((width - 1) >> 8) + 1;
    for(int r0 = 0; r0 < nrows; r0 += rows_per_pass) {
        const int nr = std::min(rows_per_pass, nrows - r0);
        RectParams p;
        p.g = w->g;
        p.mats = w->d_mats;
        p.n_mat = w->n_mat;
        p.aos = (uint32_t*)stage;
        p.present = w->d_present;
        p.rx = 0;
        p.ry = y + r0;
        p.rw = width;
        p.rh = nr;
        fss_cell* h = cells + (size_t)(y + r0 - w->g.sy0) * pitch_cells;
        const size_t row_bytes = (size_t)width * sizeof(fss_cell);
        if(upload) {
            if(pitch_cells == (size_t)width) CU(cudaMemcpyAsync(stage, h, row_bytes * nr, cudaMemcpyHostToDevice, st));
            else CU(cudaMemcpy2DAsync(stage, row_bytes, h, pitch_cells * sizeof(fss_cell), row_bytes, nr, cudaMemcpyHostToDevice, st));
            k_upload<<<dim3(nblk, nr), 256, 0, st>>>(p);
        } else {
            k_download<<<dim3(nblk, nr), 256, 0, st>>>(p);
            if(pitch_cells == (size_t)width) CU(cudaMemcpyAsync(h, stage, row_bytes * nr, cudaMemcpyDeviceToHost, st));
            else CU(cudaMemcpy2DAsync(h, pitch_cells * sizeof(fss_cell), stage, row_bytes, row_bytes, nr, cudaMemcpyDeviceToHost, st));
        }
        w->stats.kernel_launches++;
        CU(cudaGetLastError());
    }
    return FSS_OK;
}

int fss_tick_host(fss_world* w, uint32_t tick_ct, fss_cell* cells, size_t pitch_cells) {
    int rc = ready(w);
    if(rc) return rc;
    if(!cells) return fail(FSS_ERR_INVALID, "null argument");
    if(pitch_cells < (size_t)w->cfg.width) return fail(FSS_ERR_INVALID, "pitch %zu < width %u", pitch_cells, w->cfg.width);
    const int y0 = w->g.sy0, rows = w->g.rows, width = (int)w->cfg.width;
    static const int band_rows = getenv("FSS_HOST_BAND_ROWS") ? std::max(32 * 26, atoi(getenv("FSS_HOST_BAND_ROWS")) / 32 * 32) : FSS_HOST_BAND_ROWS;
    int n_bands = std::min(32, w->zh / band_rows);
    if(w->sharded || n_bands < 2) { // nothing to overlap (or halo exchanges in between): the plain sequence
        if((rc = fss_upload_rect(w, 0, y0, width, rows, cells, pitch_cells))) return rc;
        if((rc = fss_tick(w, tick_ct))) return rc;
        return fss_download_rect(w, 0, y0, width, rows, cells, pitch_cells);
    }
    if(!w->up_stream) {
        CU(cudaStreamCreateWithFlags(&w->up_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&w->down_stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&w->d_stage_down, w->stage_bytes));
        for(int k = 0; k < 32; k++) {
            CU(cudaEventCreateWithFlags(&w->ev_up[k], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&w->ev_wave[k], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&w->ev_stage_free, cudaEventDisableTiming));
        CU(cudaHostAlloc((void**)&w->h_seen, 32 * FSS_MAX_MATERIALS, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void**)&w->d_seen_host, (void*)w->h_seen, 0));
    }
    // whatever the world's stream still has queued comes first
    CU(cudaEventRecord(w->ev_stage_free, w->stream));
    CU(cudaStreamWaitEvent(w->up_stream, w->ev_stage_free, 0));
    CU(cudaStreamWaitEvent(w->down_stream, w->ev_stage_free, 0));
    const int zbands = w->zh / FSS_ZONE_ALIGN_Y; // 32-row bands of the zone
    auto K = [&](int b) { return b >= n_bands ? zbands : (int)((long long)zbands * b / n_bands); };   // band b ends before 32-row band K(b+1)
    auto row_end = [&](int b) { return b + 1 >= n_bands ? y0 + rows : w->zy + K(b + 1) * FSS_ZONE_ALIGN_Y; }; // first stored row after band b
    auto front = [&](int b, int s) { return b < 0 ? 0 : (b + 1 >= n_bands ? zbands : std::max(0, K(b + 1) - (s + 1))); };
    // (written by a small kernel into mapped host memory: a device -> host copy on the upload stream would wait in the copy
    // engine's queue behind the download stream's 256 MB passes, and the next band's upload behind it)
    const uint8_t* seen = w->h_seen;
    auto upload_band = [&](int b) -> int {
        const int lo = b == 0 ? y0 : row_end(b - 1), hi = row_end(b);
        if(int r = host_copy_rows(w, w->up_stream, w->d_stage, lo, hi - lo, cells, pitch_cells, true)) return r;
        k_copy_words<<<1, 64, 0, w->up_stream>>>((const uint32_t*)w->d_present, (uint32_t*)(w->d_seen_host + (size_t)b * FSS_MAX_MATERIALS), FSS_MAX_MATERIALS / 4);
        CU(cudaEventRecord(w->ev_up[b], w->up_stream));
        return FSS_OK;
    };
    if((rc = upload_band(0))) return rc;
    int down_lo = y0; // next row to go back to the host
    for(int b = 0; b < n_bands; b++) {
        if(b + 1 < n_bands && (rc = upload_band(b + 1))) return rc; // keep the bus busy while the host waits below
        // which materials the bands so far hold decides which kernel specialisation is exact for this wave: everything wave b
        // computes is a function of the bands 0 .. b only (see above)
        CU(cudaEventSynchronize(w->ev_up[b]));
        for(int m = 0; m < FSS_MAX_MATERIALS; m++) w->present[m] = w->present[m] || seen[(size_t)b * FSS_MAX_MATERIALS + m];
        recompute_liq(w);
        const int lo = b == 0 ? y0 : row_end(b - 1), hi = row_end(b);
        CU(cudaStreamWaitEvent(w->stream, w->ev_up[b], 0));
        if(dirty_chunks(w, 0, lo, width, hi - lo)) return FSS_ERR_CUDA;
        for(int s = 0; s < 24; s++) {
            const int j0 = front(b - 1, s), j1 = front(b, s);
            if(j1 > j0 && (rc = launch_substep(w, tick_ct, s / 4, s % 4, j0, j1))) return rc;
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(w->ev_wave[b], w->stream));
        // rows that no later wave touches: above the last sub-step's front, less one chunk pair (a visit writes 2 rows up)
        const int final_hi = b + 1 >= n_bands ? y0 + rows : std::max(down_lo, w->zy + (front(b, 23) - 1) * FSS_ZONE_ALIGN_Y);
        if(final_hi > down_lo) {
            CU(cudaStreamWaitEvent(w->down_stream, w->ev_wave[b], 0));
            if((rc = host_copy_rows(w, w->down_stream, w->d_stage_down, down_lo, final_hi - down_lo, cells, pitch_cells, false))) return rc;
            down_lo = final_hi;
        }
    }
    w->stats.substeps += 24;
    CU(cudaStreamSynchronize(w->down_stream));
    CU(cudaStreamSynchronize(w->up_stream));
    CU(cudaStreamSynchronize(w->stream));
    return FSS_OK;
}

int fss_tick_temperature(fss_world* w) {
    int rc = ready(w);
    if(rc) return rc;
    TempParams p;
    p.g = w->g;
    p.out = w->temp2;
    p.mats = w->d_mats;
    p.n_mat = w->n_mat;
    p.zx = w->zx;
    p.zy = w->zy;
    p.zw = w->zw;
    p.zh = w->zh;
    p.own_lo = w->own_lo;
    p.own_hi = w->own_hi;
    const uint32_t slots = w->g.pitch / 8 * (8 / K2_COLS); // one thread per K2_COLS columns (pitch is a multiple of 256)
    const dim3 grid((slots + TEMP_THREADS - 1) / TEMP_THREADS, (w->g.rows + TEMP_ROWS - 1) / TEMP_ROWS);
    k_temperature<<<grid, TEMP_THREADS, 0, w->stream>>>(p);
    w->stats.kernel_launches++;
    CU(cudaGetLastError());
    std::swap(w->g.temp, w->temp2);
    if(w->comm) return fss_exchange_halo_temperature(w);
    return FSS_OK;
}

// ---- side outputs -------------------------------------------------------------------------------------------
// the two counters of the spawn buffer; refusals (SubstepParams::part_count[1]) are moved into the statistics
static int spawn_counters(fss_world* w, unsigned long long& total) {
    unsigned long long c[2] = {0, 0};
    CU(cudaMemcpyAsync(c, w->d_part_count, 16, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    total = c[0];
    if(c[1]) {
        w->stats.spawns_refused += c[1];
        CU(cudaMemsetAsync(w->d_part_count + 1, 0, 8, w->stream));
    }
    return FSS_OK;
}

int fss_drain_particles(fss_world* w, fss_particle_spawn* out, size_t cap, size_t* n_out, size_t* n_total) {
    if(!w || (!out && cap)) return fail(FSS_ERR_INVALID, "null argument");
    if(use_device(w)) return FSS_ERR_CUDA;
    unsigned long long total = 0;
    if(int rc = spawn_counters(w, total)) return rc;
    if(n_out) *n_out = 0;
    if(n_total) *n_total = (size_t)total;
    if(total > cap) { // nothing is consumed: the caller comes back with room for *n_total records (cap == 0: a plain query)
        if(cap == 0) return FSS_OK;
        return fail(FSS_ERR_RANGE, "%llu spawn records waiting, room for %zu: nothing was drained", total, cap);
    }
    const size_t n = (size_t)total; // (the kernels never write more than part_cap records: reserve_spawns, fss_rules.cuh)
    std::vector<fss_particle_spawn> v(n);
    if(n) CU(cudaMemcpy(v.data(), w->d_parts, n * sizeof(fss_particle_spawn), cudaMemcpyDeviceToHost));
    CU(cudaMemsetAsync(w->d_part_count, 0, 8, w->stream));
    const int zx = w->zx, zy = w->zy; // the order World::tick pushes them (fss_spawn_order_key); equal keys are identical records
    std::stable_sort(v.begin(), v.end(), [zx, zy](const fss_particle_spawn& a, const fss_particle_spawn& b) {
        return fss_spawn_order_key(a.tick_iter, a.x, a.y, zx, zy) < fss_spawn_order_key(b.tick_iter, b.x, b.y, zx, zy);
    });
    if(n) memcpy(out, v.data(), n * sizeof(fss_particle_spawn));
    if(n_out) *n_out = n;
    return FSS_OK;
}


// ---- World::particles on the device (fss_particles.cuh) -----------------------------------------------------
static int particles_spiral_init();
static int particles_reserve(fss_world* w, size_t need) {
    if(need > 0x7FFFFFF0u) return fail(FSS_ERR_RANGE, "%zu particles", need);
    if(!w->pl_ctl) {
        if(int rc = particles_spiral_init()) return rc;
        CU(cudaMalloc(&w->pl_ctl, 8 * sizeof(uint32_t)));
        const size_t tiles = (size_t)((w->cfg.width + FSS_PTILE - 1) / FSS_PTILE) * ((w->cfg.height + FSS_PTILE - 1) / FSS_PTILE);
        CU(cudaMalloc(&w->pl_res_tile, 3 * tiles * sizeof(uint32_t))); // res_tile, w_tile, w_when
    }
    if(need <= w->pl_cap) return FSS_OK;
    size_t cap = std::max<size_t>(w->pl_cap ? (size_t)w->pl_cap * 2 : 4096, need);
    fss_particle *list = nullptr, *next = nullptr;
    CU(cudaMalloc(&list, cap * sizeof(fss_particle)));
    CU(cudaMalloc(&next, cap * sizeof(fss_particle)));
    if(w->pl_n) CU(cudaMemcpyAsync(list, w->pl_list, (size_t)w->pl_n * sizeof(fss_particle), cudaMemcpyDeviceToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    void* old[] = {w->pl_list, w->pl_next, w->pl_decided, w->pl_keep, w->pl_sure, w->pl_big, w->pl_ran, w->pl_listed, w->pl_act, w->pl_seen, w->pl_reach, w->pl_res_cell};
    for(void* p : old) cudaFree(p);
    w->pl_list = list;
    w->pl_next = next;
    CU(cudaMalloc(&w->pl_decided, cap));
    CU(cudaMalloc(&w->pl_keep, cap));
    CU(cudaMalloc(&w->pl_sure, cap));
    CU(cudaMalloc(&w->pl_big, cap));
    CU(cudaMalloc(&w->pl_ran, cap * sizeof(uint32_t)));
    CU(cudaMalloc(&w->pl_listed, cap * sizeof(uint32_t)));
    CU(cudaMalloc(&w->pl_act, cap * sizeof(int4)));
    CU(cudaMalloc(&w->pl_seen, cap * sizeof(int4)));
    CU(cudaMalloc(&w->pl_reach, cap * sizeof(int4)));
    size_t slots = 4096; // >= 16 slots per particle keeps chance collisions (which only delay a particle by a round) rare
    while(slots < cap * 16) slots *= 2;
    CU(cudaMalloc(&w->pl_res_cell, 4 * slots * sizeof(uint32_t)));
    w->pl_res_mask = (uint32_t)(slots - 1);
    w->pl_cap = (uint32_t)cap;
    return FSS_OK;
}

// the visiting order of the landing search, world.cpp:2256-2290 (`X = Y = 32`, x/y/dx/dy as there)
static int particles_spiral_init() {
    char2 h[FSS_PSPIRAL_N];
    int X = 32, Y = 32, x = 0, y = 0, dx = 0, dy = -1, n = 0;
    for(int j = 0; j < 32 * 32; j++) {
        if((-X / 2 <= x) && (x <= X / 2) && (-Y / 2 <= y) && (y <= Y / 2)) {
            h[n].x = (signed char)x;
            h[n].y = (signed char)y;
            n++;
        }
        if((x == y) || ((x < 0) && (x == -y)) || ((x > 0) && (x == 1 - y))) {
            int t = dx;
            dx = -dy;
            dy = t;
        }
        x += dx;
        y += dy;
    }
    if(n != FSS_PSPIRAL_N) return fail(FSS_ERR_STATE, "landing search visits %d cells, expected %d", n, FSS_PSPIRAL_N);
    CU(cudaMemcpyToSymbol(c_spiral, h, sizeof h));
    return FSS_OK;
}

static int cub_scratch(fss_world* w, size_t bytes) {
    if(bytes <= w->pl_tmp_bytes) return FSS_OK;
    cudaFree(w->pl_tmp);
    w->pl_tmp = nullptr;
    w->pl_tmp_bytes = 0;
    CU(cudaMalloc(&w->pl_tmp, bytes));
    w->pl_tmp_bytes = bytes;
    return FSS_OK;
}

// what World::tick's `particles.insert(particles.end(), pts.begin(), pts.end())` (world.cpp:1989-1994) did for the ticks
// since the last call: the spawn records, in push order, become Particles at the end of the list
static int particles_adopt(fss_world* w) {
    unsigned long long total = 0;
    if(int rc = spawn_counters(w, total)) return rc;
    if(total == 0) return FSS_OK;
    const uint32_t n = (uint32_t)std::min<unsigned long long>(total, w->part_cap); // (never more than part_cap: reserve_spawns)
    if(int rc = particles_reserve(w, (size_t)w->pl_n + n)) return rc;
    if(!w->pl_key[0]) {
        for(int k = 0; k < 2; k++) {
            CU(cudaMalloc(&w->pl_key[k], (size_t)w->part_cap * sizeof(uint64_t)));
            CU(cudaMalloc(&w->pl_idx[k], (size_t)w->part_cap * sizeof(uint32_t)));
        }
    }
    AdoptParams A;
    A.spawn = w->d_parts;
    A.key = w->pl_key[0];
    A.idx = w->pl_idx[0];
    A.key_sorted = w->pl_key[1];
    A.idx_sorted = w->pl_idx[1];
    A.dst = w->pl_list + w->pl_n;
    A.n = n;
    A.zx = w->zx;
    A.zy = w->zy;
    k_spawn_keys<<<(n + 255) / 256, 256, 0, w->stream>>>(A);
    size_t bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, w->pl_key[0], w->pl_key[1], w->pl_idx[0], w->pl_idx[1], (int)n, 0, 64, w->stream));
    if(int rc = cub_scratch(w, bytes)) return rc;
    CU(cub::DeviceRadixSort::SortPairs(w->pl_tmp, bytes, w->pl_key[0], w->pl_key[1], w->pl_idx[0], w->pl_idx[1], (int)n, 0, 64, w->stream));
    k_spawn_to_particles<<<(n + 255) / 256, 256, 0, w->stream>>>(A);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(w->d_part_count, 0, 8, w->stream));
    w->stats.kernel_launches += 2;
    w->pl_n += n;
    return FSS_OK;
}

int fss_particles_add(fss_world* w, const fss_particle* particles, size_t n) {
    if(!w || (!particles && n)) return fail(FSS_ERR_INVALID, "null argument");
    if(use_device(w)) return FSS_ERR_CUDA;
    if(w->sharded) return fail(FSS_ERR_STATE, "device particles are not available on strip-sharded worlds");
    if(!w->mats_set) return fail(FSS_ERR_STATE, "fss_set_materials first");
    for(size_t i = 0; i < n; i++)
        if(particles[i].tile.material >= w->n_mat) return fail(FSS_ERR_INVALID, "particle %zu: material %u of %d", i, particles[i].tile.material, w->n_mat);
    if(n == 0) return FSS_OK;
    w->pl_used = true;
    if(int rc = particles_adopt(w)) return rc; // host pushes come after what the ticks so far pushed
    if(int rc = particles_reserve(w, (size_t)w->pl_n + n)) return rc;
    CU(cudaMemcpyAsync(w->pl_list + w->pl_n, particles, n * sizeof(fss_particle), cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->pl_n += (uint32_t)n;
    bool grew = false;
    for(size_t i = 0; i < n; i++) {
        grew = grew || !w->present[particles[i].tile.material];
        w->present[particles[i].tile.material] = true;
    }
    if(grew) recompute_liq(w);
    return FSS_OK;
}

int fss_particles_count(fss_world* w, size_t* n) {
    if(!w || !n) return fail(FSS_ERR_INVALID, "null argument");
    *n = w->pl_n;
    return FSS_OK;
}

int fss_particles_download(fss_world* w, fss_particle* out, size_t cap, size_t* n_out) {
    if(!w || (!out && cap)) return fail(FSS_ERR_INVALID, "null argument");
    if(use_device(w)) return FSS_ERR_CUDA;
    const size_t n = std::min<size_t>(cap, w->pl_n);
    if(n) CU(cudaMemcpyAsync(out, w->pl_list, n * sizeof(fss_particle), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    if(n_out) *n_out = n;
    return n < w->pl_n ? fail(FSS_ERR_RANGE, "%u particles, room for %zu", w->pl_n, cap) : FSS_OK;
}

int fss_particles_clear(fss_world* w) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    w->pl_n = 0;
    return FSS_OK;
}

int fss_particles_rounds(fss_world* w, uint64_t* rounds) {
    if(!w || !rounds) return fail(FSS_ERR_INVALID, "null argument");
    rounds[0] = w->pl_rounds;
    rounds[1] = w->pl_passes;
    return FSS_OK;
}

int fss_tick_particles(fss_world* w) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(use_device(w)) return FSS_ERR_CUDA;
    if(w->sharded) return fail(FSS_ERR_STATE, "device particles are not available on strip-sharded worlds");
    if(!w->zone_set || !w->mats_set) return fail(FSS_ERR_STATE, "fss_set_materials and fss_set_tick_zone first");
    w->pl_used = true;
    if(int rc = particles_adopt(w)) return rc;
    w->pl_rounds = w->pl_passes = 0;
    const uint32_t n = w->pl_n;
    if(n == 0) return FSS_OK;
    if(int rc = particles_reserve(w, n)) return rc;
    ParticleParams P;
    P.g = w->g;
    P.mats = w->d_mats;
    P.n_mat = w->n_mat;
    P.width = (int)w->cfg.width;
    P.height = (int)w->cfg.height;
    P.zx = w->zx;
    P.zy = w->zy;
    P.zw = w->zw;
    P.zh = w->zh;
    P.in = w->pl_list;
    P.out = w->pl_next;
    P.decided = w->pl_decided;
    P.keep = w->pl_keep;
    P.sure = w->pl_sure;
    P.act = w->pl_act;
    P.seen = w->pl_seen;
    P.reach = w->pl_reach;
    const size_t slots = (size_t)w->pl_res_mask + 1;
    P.w_act = w->pl_res_cell;
    P.r_act = w->pl_res_cell + slots;
    P.w_pot = w->pl_res_cell + 2 * slots;
    P.r_pot = w->pl_res_cell + 3 * slots;
    P.res_mask = w->pl_res_mask;
    P.res_tile = w->pl_res_tile;
    P.big = w->pl_big;
    P.tiles_w = (int)((w->cfg.width + FSS_PTILE - 1) / FSS_PTILE);
    P.tiles_h = (int)((w->cfg.height + FSS_PTILE - 1) / FSS_PTILE);
    const size_t n_tiles = (size_t)P.tiles_w * P.tiles_h;
    P.w_tile = w->pl_res_tile + n_tiles;
    P.w_when = w->pl_res_tile + 2 * n_tiles;
    P.ran = w->pl_ran;
    P.listed = w->pl_listed;
    P.ctl = w->pl_ctl;
    P.n = n;
    P.pass = 0;
    P.q = (w->cfg.flags & FSS_NO_EARLY_OUT) ? nullptr : w->d_q;
    P.q_w = w->q_w;
    P.q_h = w->q_h;
    P.q_cj0 = w->q_cj0;
    P.tile_mask = 0;
    static const bool no_small = getenv("FSS_PARTICLES_NO_SMALL") != nullptr; // measurements: always the multi-kernel path
    if(n <= FSS_PCOOP_MAX && !no_small) {
        // a short list: rounds, passes and compaction inside one cooperative launch (k_particles_coop), on its own small tables
        // table sizes follow the list (they are cleared every round): >= 64 cell slots and >= 8 tile slots per particle
        uint32_t cslots = 4096, tslots = 1024;
        while(cslots < 64 * n) cslots *= 2;
        while(tslots < 8 * n) tslots *= 2;
        if(!w->pl_small) {
            CU(cudaMalloc(&w->pl_small, (4 * (64u * FSS_PCOOP_MAX) + 3 * (8u * FSS_PCOOP_MAX)) * sizeof(uint32_t)));
            int per_sm = 0, sms = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_particles_coop, FSS_PCOOP_THREADS, 0));
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device));
            w->pl_coop_ctas = std::max(1, per_sm * sms);
        }
        SmallTables T;
        T.words = w->pl_small;
        T.n_ff = 4 * cslots + 2 * tslots;
        T.n_zero = tslots;
        P.w_act = w->pl_small;
        P.r_act = w->pl_small + cslots;
        P.w_pot = w->pl_small + 2 * cslots;
        P.r_pot = w->pl_small + 3 * cslots;
        P.res_mask = cslots - 1;
        P.res_tile = w->pl_small + 4 * cslots;
        P.w_tile = P.res_tile + tslots;
        P.w_when = P.w_tile + tslots;
        P.tile_mask = tslots - 1;
        const unsigned ctas = (unsigned)std::min<size_t>((size_t)w->pl_coop_ctas, ((size_t)n * FSS_PCOOP_SPREAD + FSS_PCOOP_THREADS - 1) / FSS_PCOOP_THREADS);
        void* args[] = {&P, &T};
        CU(cudaLaunchCooperativeKernel((const void*)k_particles_coop, dim3(ctas), dim3(FSS_PCOOP_THREADS), args, 0, w->stream));
        w->stats.kernel_launches++;
        w->pl_rounds = w->pl_passes = 0; // not counted on this path
        uint32_t kept = 0;
        CU(cudaMemcpyAsync(&kept, w->pl_ctl + 2, sizeof kept, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        w->pl_n = kept;
        return FSS_OK;
    }
    const unsigned blocks = (n + 127) / 128;
    CU(cudaMemsetAsync(w->pl_decided, 0, n, w->stream));
    for(uint32_t left = n; left;) {
        const uint32_t ctl0[4] = {0u, 0xFFFFFFFFu, 0u, 0u};
        CU(cudaMemcpyAsync(w->pl_ctl, ctl0, sizeof ctl0, cudaMemcpyHostToDevice, w->stream));
        CU(cudaMemsetAsync(w->pl_res_cell, 0xFF, 4 * slots * sizeof(uint32_t), w->stream));
        CU(cudaMemsetAsync(w->pl_res_tile, 0xFF, 2 * n_tiles * sizeof(uint32_t), w->stream));
        CU(cudaMemsetAsync(P.w_when, 0, n_tiles * sizeof(uint32_t), w->stream));
        k_particles_speculate<<<blocks, 128, 0, w->stream>>>(P);
        w->stats.kernel_launches++;
        // "who is unsure" to its fixed point: a pass that changes nobody ends it
        for(uint32_t last_change = 0; P.pass < FSS_PCASCADE_MAX;) {
            for(int k = 0; k < FSS_PCASCADE_BATCH; k++) {
                P.pass++;
                k_particles_cascade<false><<<blocks, 128, 0, w->stream>>>(P);
            }
            w->stats.kernel_launches += FSS_PCASCADE_BATCH;
            CU(cudaMemcpyAsync(&last_change, w->pl_ctl + 3, sizeof last_change, cudaMemcpyDeviceToHost, w->stream));
            CU(cudaStreamSynchronize(w->stream));
            w->pl_passes += FSS_PCASCADE_BATCH;
            if(last_change < P.pass) break;
        }
        P.pass = 0;
        k_particles_cascade<true><<<blocks, 128, 0, w->stream>>>(P);
        k_particles_commit<<<blocks, 128, 0, w->stream>>>(P);
        CU(cudaGetLastError());
        w->stats.kernel_launches += 2;
        w->pl_rounds++;
        uint32_t now = 0;
        CU(cudaMemcpyAsync(&now, w->pl_ctl, sizeof now, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        if(now >= left) return fail(FSS_ERR_CUDA, "particle round %llu retired nobody (%u left)", (unsigned long long)w->pl_rounds, now);
        left = now;
    }
    // particles.erase(std::remove_if(...)): the survivors, in order
    size_t bytes = 0;
    CU(cub::DeviceSelect::Flagged(nullptr, bytes, w->pl_next, w->pl_keep, w->pl_list, w->pl_ctl + 2, (int)n, w->stream));
    if(int rc = cub_scratch(w, bytes)) return rc;
    CU(cub::DeviceSelect::Flagged(w->pl_tmp, bytes, w->pl_next, w->pl_keep, w->pl_list, w->pl_ctl + 2, (int)n, w->stream));
    uint32_t kept = 0;
    CU(cudaMemcpyAsync(&kept, w->pl_ctl + 2, sizeof kept, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->pl_n = kept;
    return FSS_OK;
}


// ---- dirty -> pixels (fss_render.cuh; Game.cpp:2579-2650, :2751) ----------------------------------------------
int fss_set_material_looks(fss_world* w, const uint8_t* alpha, const uint32_t* emit_color, int n) {
    if(!w || !alpha || !emit_color) return fail(FSS_ERR_INVALID, "null argument");
    if(n < 0 || n > FSS_MAX_MATERIALS) return fail(FSS_ERR_INVALID, "%d materials", n);
    memcpy(w->h_alpha, alpha, (size_t)n);
    memcpy(w->h_emit, emit_color, (size_t)n * 4);
    w->looks_stale = true;
    return FSS_OK;
}

// the marks that skipped visits of settled chunks stand for (k_skipped_marks), before anybody looks at World::dirty
static int skipped_marks(fss_world* w) {
    if(!w->d_qskip || !w->d_q || !w->g.dirty) return FSS_OK;
    SkippedMarks p;
    p.g = w->g;
    p.qskip = w->d_qskip;
    p.q_w = w->q_w;
    p.q_cj0 = w->q_cj0;
    p.zx = w->zx;
    p.zy = w->zy;
    p.zw = w->zw;
    p.zh = w->zh;
    p.cells = plane_cells(w);
    k_skipped_marks<<<(unsigned)((p.cells + 255) / 256), 256, 0, w->stream>>>(p);
    CU(cudaGetLastError());
    w->stats.kernel_launches++;
    CU(cudaMemsetAsync(w->d_qskip, 0xFF, (size_t)w->q_w * w->q_h, w->stream));
    return FSS_OK;
}

static int dirty_ready(fss_world* w) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(!w->g.dirty) return fail(FSS_ERR_STATE, "the world was created without FSS_TRACK_DIRTY");
    return use_device(w);
}

int fss_mark_dirty(fss_world* w, int x, int y, int width, int height) {
    if(int rc = dirty_ready(w)) return rc;
    // like the reference's own writes, cells outside the world are skipped
    // (a strip marks the rows it owns: the same call on the neighbouring strip marks the rest)
    const int x0 = std::max(x, 0), y0 = std::max(y, w->own_lo), x1 = std::min<long long>((long long)x + width, w->cfg.width), y1 = std::min<long long>((long long)y + height, w->own_hi);
    if(x1 <= x0 || y1 <= y0) return FSS_OK;
    const long long cells = (long long)(x1 - x0) * (y1 - y0);
    if(cells > 0x7FFFFFFFll) return fail(FSS_ERR_RANGE, "rectangle of %lld cells", cells);
    k_mark_dirty<<<(unsigned)((cells + 255) / 256), 256, 0, w->stream>>>(w->g, x0, y0, x1 - x0, y1 - y0);
    CU(cudaGetLastError());
    w->stats.kernel_launches++;
    return FSS_OK;
}

int fss_mark_dirty_list(fss_world* w, const fss_rect* rects, size_t n) {
    if(int rc = dirty_ready(w)) return rc;
    if(!rects && n) return fail(FSS_ERR_INVALID, "null argument");
    // clipped like fss_mark_dirty; one launch per 65535 rectangles (gridDim.y), the list goes through the staging buffer
    std::vector<int4> v;
    long long biggest = 0;
    for(size_t i = 0; i < n; i++) {
        const int x0 = std::max(rects[i].x, 0), y0 = std::max(rects[i].y, w->own_lo);
        const int x1 = (int)std::min<long long>((long long)rects[i].x + rects[i].w, w->cfg.width), y1 = (int)std::min<long long>((long long)rects[i].y + rects[i].h, w->own_hi);
        if(x1 <= x0 || y1 <= y0) continue;
        v.push_back(make_int4(x0, y0, x1 - x0, y1 - y0));
        biggest = std::max(biggest, (long long)(x1 - x0) * (y1 - y0));
    }
    const size_t per_pass = std::min<size_t>(65535, w->stage_bytes / sizeof(int4));
    for(size_t k = 0; k < v.size(); k += per_pass) {
        const size_t m = std::min(per_pass, v.size() - k);
        CU(cudaMemcpyAsync(w->d_stage, v.data() + k, m * sizeof(int4), cudaMemcpyHostToDevice, w->stream));
        const unsigned bx = (unsigned)std::min<long long>(1024, (biggest + 255) / 256);
        k_mark_dirty_list<<<dim3(bx, (unsigned)m), 256, 0, w->stream>>>(w->g, (const int4*)w->d_stage);
        CU(cudaGetLastError());
        w->stats.kernel_launches++;
        CU(cudaStreamSynchronize(w->stream)); // `v` and the staging buffer are reused
    }
    return FSS_OK;
}

int fss_render_dirty(fss_world* w, uint64_t* moving_tiles, int n_materials, fss_render_info* info) {
    if(int rc = dirty_ready(w)) return rc;
    if(!w->mats_set) return fail(FSS_ERR_STATE, "fss_set_materials first");
    if(n_materials < 0 || n_materials > FSS_MAX_MATERIALS) return fail(FSS_ERR_INVALID, "%d materials", n_materials);
    if(int rc = skipped_marks(w)) return rc;
    const size_t cells = plane_cells(w), pixels = (size_t)w->cfg.width * w->g.rows; // the layers hold the rows this world stores
    if(!w->px[0]) {
        for(int k = 0; k < 4; k++) {
            CU(cudaMalloc(&w->px[k], pixels * 4));
            CU(cudaMemsetAsync(w->px[k], 0, pixels * 4, w->stream));
        }
        CU(cudaMalloc(&w->prev_fx, cells * 4));
        CU(cudaMalloc(&w->prev_fy, cells * 4));
        CU(cudaMemsetAsync(w->prev_fx, 0, cells * 4, w->stream)); // world.cpp:179-180
        CU(cudaMemsetAsync(w->prev_fy, 0, cells * 4, w->stream));
        CU(cudaMalloc(&w->d_looks, sizeof(DLook) * FSS_MAX_MATERIALS));
        CU(cudaMalloc(&w->d_rinfo, 8 * FSS_MAX_MATERIALS + 16));
        w->looks_stale = true;
    }
    if(w->looks_stale) {
        DLook h[FSS_MAX_MATERIALS];
        for(int m = 0; m < FSS_MAX_MATERIALS; m++) {
            h[m].alpha = w->h_alpha[m];
            h[m].emit_color = w->h_emit[m];
            h[m].iterations = w->h_iterations[m];
        }
        CU(cudaMemcpyAsync(w->d_looks, h, sizeof h, cudaMemcpyHostToDevice, w->stream));
        CU(cudaStreamSynchronize(w->stream)); // h is on the stack
        w->looks_stale = false;
    }
    CU(cudaMemsetAsync(w->d_rinfo, 0, 8 * FSS_MAX_MATERIALS + 16, w->stream));
    RenderParams p;
    p.g = w->g;
    p.looks = w->d_looks;
    p.n_mat = w->n_mat;
    p.width = (int)w->cfg.width;
    for(int k = 0; k < 4; k++) p.px[k] = w->px[k];
    p.prev_fx = w->prev_fx;
    p.prev_fy = w->prev_fy;
    p.moving = w->d_rinfo;
    p.info = (uint32_t*)(w->d_rinfo + FSS_MAX_MATERIALS);
    p.cells = cells;
    p.own_lo = w->own_lo;
    p.own_hi = w->own_hi;
    k_render_dirty<<<(unsigned)((cells + 255) / 256), 256, 0, w->stream>>>(p);
    CU(cudaGetLastError());
    w->stats.kernel_launches++;
    unsigned long long h[FSS_MAX_MATERIALS + 2];
    CU(cudaMemcpyAsync(h, w->d_rinfo, sizeof h, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    if(moving_tiles)
        for(int m = 0; m < n_materials; m++) moving_tiles[m] = h[m];
    if(info) {
        const uint32_t* f = (const uint32_t*)(h + FSS_MAX_MATERIALS);
        info->had_dirty = f[0];
        info->had_fire = f[1];
        info->had_flow = f[2];
    }
    return FSS_OK;
}

int fss_download_pixels(fss_world* w, int layer, int x, int y, int width, int height, uint8_t* rgba, size_t pitch_bytes) {
    if(int rc = dirty_ready(w)) return rc;
    if(!rgba || layer < 0 || layer > 3) return fail(FSS_ERR_INVALID, "bad argument");
    if(!w->px[0]) return fail(FSS_ERR_STATE, "fss_render_dirty first");
    if(width <= 0 || height <= 0 || x < 0 || y < w->own_lo || (uint32_t)(x + width) > w->cfg.width || y + height > w->own_hi)
        return fail(FSS_ERR_RANGE, "rectangle (%d,%d,%d,%d) outside the rows [%d, %d) this world renders (width %u)", x, y, width, height, w->own_lo, w->own_hi, w->cfg.width);
    if(pitch_bytes < (size_t)width * 4) return fail(FSS_ERR_INVALID, "pitch %zu < %d pixels", pitch_bytes, width);
    CU(cudaMemcpy2DAsync(rgba, pitch_bytes, w->px[layer] + (size_t)(y - w->g.sy0) * w->cfg.width + x, (size_t)w->cfg.width * 4, (size_t)width * 4, (size_t)height,
                         cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return FSS_OK;
}

int fss_download_dirty(fss_world* w, int x, int y, int width, int height, uint8_t* out, size_t pitch_bytes) {
    if(int rc = dirty_ready(w)) return rc;
    if(!out) return fail(FSS_ERR_INVALID, "null argument");
    if(pitch_bytes < (size_t)width) return fail(FSS_ERR_INVALID, "pitch %zu < width %d", pitch_bytes, width);
    if(int rc = skipped_marks(w)) return rc;
    size_t skip = 0;
    if(int rc = rect_clip(w, x, y, width, height, skip)) return rc;
    const int rows_per_pass = (int)std::max<size_t>(1, w->stage_bytes / (size_t)width);
    for(int r0 = 0; r0 < height; r0 += rows_per_pass) {
        const int nr = std::min(rows_per_pass, height - r0);
        k_download_bytes<<<(unsigned)(((size_t)width * nr + 255) / 256), 256, 0, w->stream>>>(w->g.dirty, (uint8_t*)w->d_stage, w->g.pitch, w->g.sy0, x, y + r0, width, nr);
        w->stats.kernel_launches++;
        CU(cudaMemcpy2DAsync(out + (skip + r0) * pitch_bytes, pitch_bytes, w->d_stage, (size_t)width, (size_t)width, nr, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
    }
    return FSS_OK;
}

static int reduce_params(fss_world* w, ReduceParams& p, int n_materials) {
    if(use_device(w)) return FSS_ERR_CUDA;
    p.g = w->g;
    p.width = w->cfg.width;
    p.own_lo = w->own_lo;
    p.own_hi = w->own_hi;
    p.n_mat = n_materials;
    p.out_u64 = w->d_red;
    p.out_f64 = nullptr;
    p.mats = w->d_mats;
    CU(cudaMemsetAsync(w->d_red, 0, FSS_MAX_MATERIALS * 16, w->stream));
    return FSS_OK;
}

int fss_hash(fss_world* w, uint64_t* digest) {
    if(!w || !digest) return fail(FSS_ERR_INVALID, "null argument");
    ReduceParams p;
    int rc = reduce_params(w, p, 0);
    if(rc) return rc;
    k_hash<<<148 * 8, 256, 0, w->stream>>>(p);
    w->stats.kernel_launches++;
    CU(cudaMemcpyAsync(digest, w->d_red, 8, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return FSS_OK;
}

int fss_histogram(fss_world* w, uint64_t* counts, int n_materials) {
    if(!w || !counts || n_materials < 1 || n_materials > FSS_MAX_MATERIALS) return fail(FSS_ERR_INVALID, "bad argument");
    ReduceParams p;
    int rc = reduce_params(w, p, n_materials);
    if(rc) return rc;
    k_histogram<<<148 * 16, 256, 0, w->stream>>>(p);
    w->stats.kernel_launches++;
    CU(cudaMemcpyAsync(counts, w->d_red, 8 * (size_t)n_materials, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return FSS_OK;
}

int fss_fluid_totals(fss_world* w, double* totals, int n_materials) {
    if(!w || !totals || n_materials < 1 || n_materials > FSS_MAX_MATERIALS) return fail(FSS_ERR_INVALID, "bad argument");
    ReduceParams p;
    int rc = reduce_params(w, p, n_materials);
    if(rc) return rc;
    p.out_f64 = (double*)(w->d_red + FSS_MAX_MATERIALS);
    k_histogram<<<148 * 16, 256, 0, w->stream>>>(p);
    w->stats.kernel_launches++;
    CU(cudaMemcpyAsync(totals, p.out_f64, 8 * (size_t)n_materials, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return FSS_OK;
}

int fss_fill_synthetic(fss_world* w, const fss_fill_spec* spec) {
    if(!w || !spec) return fail(FSS_ERR_INVALID, "null argument");
    if(!w->mats_set) return fail(FSS_ERR_STATE, "fss_fill_synthetic before fss_set_materials");
    if(spec->n_entries < 0 || spec->n_entries > 16 || spec->n_boxes < 0 || spec->n_boxes > 256) return fail(FSS_ERR_INVALID, "bad fill spec");
    if(use_device(w)) return FSS_ERR_CUDA;
    FillParams p;
    p.g = w->g;
    p.mats = w->d_mats;
    p.tex = w->d_tex;
    p.n_mat = w->n_mat;
    p.width = w->cfg.width;
    p.height = w->cfg.height;
    p.spec = *spec;
    k_fill<<<dim3(w->g.rows, w->g.pitch / 256), 256, 0, w->stream>>>(p);
    w->stats.kernel_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(w->temp2, w->g.temp, plane_cells(w) * 4, cudaMemcpyDeviceToDevice, w->stream));
    auto mark_present = [&](int m) { if(m >= 0 && m < w->n_mat) w->present[m] = true; };
    mark_present(spec->border_material);
    for(int e = 0; e < spec->n_entries; e++) mark_present(spec->material[e]);
    if(spec->solid_from_row) mark_present(spec->solid_material);
    if(spec->n_boxes) mark_present(spec->sparse_material);
    recompute_liq(w); // (every strip of a sharded world is filled from the same spec)
    return dirty_chunks(w, 0, 0, -1, -1);
}

static int chunk_records(fss_world* w, int x0, int y0, int cw, int ch, fss_tile_record* recs, bool merge) {
    if(!w || !recs) return fail(FSS_ERR_INVALID, "null argument");
    if(!w->mats_set) return fail(FSS_ERR_STATE, "chunk records before fss_set_materials");
    if(cw <= 0 || ch <= 0 || (size_t)cw * ch * sizeof(fss_tile_record) > w->stage_bytes) return fail(FSS_ERR_INVALID, "chunk %dx%d", cw, ch);
    static_assert(sizeof(fss_tile_record) == 12, "MaterialInstanceData is 12 bytes (Chunk.hpp:27-31)");
    if(use_device(w)) return FSS_ERR_CUDA;
    const size_t bytes = (size_t)cw * ch * sizeof(fss_tile_record);
    ChunkParams p;
    p.g = w->g;
    p.mats = w->d_mats;
    p.n_mat = w->n_mat;
    p.recs = (uint32_t*)w->d_stage;
    p.x0 = x0;
    p.y0 = y0;
    p.cw = cw;
    p.ch = ch;
    p.width = (int)w->cfg.width;
    p.y_lo = w->g.sy0;
    p.y_hi = w->g.sy0 + w->g.rows;
    p.present = w->d_present;
    const int blocks = (cw * ch + 255) / 256;
    if(merge) {
        CU(cudaMemcpyAsync(w->d_stage, recs, bytes, cudaMemcpyHostToDevice, w->stream));
        k_merge_records<<<blocks, 256, 0, w->stream>>>(p);
        w->stats.kernel_launches++;
        CU(cudaGetLastError());
        const int cx0 = std::max(x0, 0), cy0 = std::max(y0, 0), cx1 = std::min(x0 + cw, (int)w->cfg.width), cy1 = std::min(y0 + ch, (int)w->cfg.height);
        if(cx1 > cx0 && cy1 > cy0 && dirty_chunks(w, cx0, cy0, cx1 - cx0, cy1 - cy0)) return FSS_ERR_CUDA;
        uint8_t seen[FSS_MAX_MATERIALS];
        CU(cudaMemcpyAsync(seen, w->d_present, FSS_MAX_MATERIALS, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        for(int m = 0; m < FSS_MAX_MATERIALS; m++) w->present[m] = w->present[m] || seen[m];
        recompute_liq(w);
    } else {
        k_read_records<<<blocks, 256, 0, w->stream>>>(p);
        w->stats.kernel_launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(recs, w->d_stage, bytes, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
    }
    return FSS_OK;
}

int fss_merge_chunk_records(fss_world* w, int x0, int y0, int chunk_w, int chunk_h, const fss_tile_record* records) {
    return chunk_records(w, x0, y0, chunk_w, chunk_h, (fss_tile_record*)records, true);
}
int fss_read_chunk_records(fss_world* w, int x0, int y0, int chunk_w, int chunk_h, fss_tile_record* records) {
    return chunk_records(w, x0, y0, chunk_w, chunk_h, records, false);
}

// ---- the chunk file of Chunk::read / Chunk::write (Chunk.cpp:47-281): host code, no world needed ---------------------------
#define FSS_FILE_CELLS ((size_t)FSS_CHUNK_FILE_W * FSS_CHUNK_FILE_H)

int fss_lz4_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_size) {
    if((!src && n) || !dst || !out_size) return fail(FSS_ERR_INVALID, "null argument");
    if(cap < fss_lz4::compress_bound(n)) return fail(FSS_ERR_RANGE, "LZ4 output buffer of %zu bytes, %zu needed in the worst case", cap, fss_lz4::compress_bound(n));
    *out_size = fss_lz4::compress(src, n, dst, cap);
    return FSS_OK;
}

int fss_lz4_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_size) {
    if(!src || (!dst && cap) || !out_size) return fail(FSS_ERR_INVALID, "null argument");
    const long got = fss_lz4::decompress(src, n, dst, cap);
    if(got < 0) return fail(FSS_ERR_INVALID, "malformed LZ4 block (or the output does not fit %zu bytes)", cap);
    *out_size = (size_t)got;
    return FSS_OK;
}

size_t fss_chunk_file_bound(void) {
    return 1 + 4 * sizeof(int32_t) + fss_lz4::compress_bound(FSS_FILE_CELLS * 2 * sizeof(fss_tile_record)) + fss_lz4::compress_bound(FSS_FILE_CELLS * sizeof(uint32_t));
}

int fss_chunk_file_encode(int8_t generation_phase, const fss_tile_record* tiles, const fss_tile_record* layer2, const uint32_t* background, uint8_t* out,
                          size_t cap, size_t* size) {
    if(!tiles || !layer2 || !background || !out || !size) return fail(FSS_ERR_INVALID, "null argument");
    if(cap < fss_chunk_file_bound()) return fail(FSS_ERR_RANGE, "chunk file buffer of %zu bytes, fss_chunk_file_bound() = %zu", cap, fss_chunk_file_bound());
    // Chunk::write, Chunk.cpp:208-212: the tiles layer then layer2 in one buffer of MaterialInstanceData
    std::vector<fss_tile_record> buf(FSS_FILE_CELLS * 2);
    memcpy(buf.data(), tiles, FSS_FILE_CELLS * sizeof(fss_tile_record));
    memcpy(buf.data() + FSS_FILE_CELLS, layer2, FSS_FILE_CELLS * sizeof(fss_tile_record));
    const int32_t src_size = (int32_t)(FSS_FILE_CELLS * 2 * sizeof(fss_tile_record)), src_size2 = (int32_t)(FSS_FILE_CELLS * sizeof(uint32_t));
    uint8_t* body = out + 1 + 4 * sizeof(int32_t);
    const int32_t c1 = (int32_t)fss_lz4::compress((const uint8_t*)buf.data(), (size_t)src_size, body, fss_lz4::compress_bound((size_t)src_size));
    const int32_t c2 = (int32_t)fss_lz4::compress((const uint8_t*)background, (size_t)src_size2, body + c1, fss_lz4::compress_bound((size_t)src_size2));
    out[0] = (uint8_t)generation_phase; // :262-270
    memcpy(out + 1, &src_size, 4);
    memcpy(out + 5, &c1, 4);
    memcpy(out + 9, &src_size2, 4);
    memcpy(out + 13, &c2, 4);
    *size = 17 + (size_t)c1 + (size_t)c2;
    return FSS_OK;
}

int fss_chunk_file_decode(const uint8_t* file, size_t size, int8_t* generation_phase, fss_tile_record* tiles, fss_tile_record* layer2, uint32_t* background) {
    if(!file || !tiles || !layer2 || !background) return fail(FSS_ERR_INVALID, "null argument");
    if(size < 17) return fail(FSS_ERR_INVALID, "chunk file of %zu bytes: shorter than its header", size);
    int32_t src_size, c1, src_size2, c2;
    memcpy(&src_size, file + 1, 4);
    memcpy(&c1, file + 5, 4);
    memcpy(&src_size2, file + 9, 4);
    memcpy(&c2, file + 13, 4);
    // Chunk::read throws on these two (Chunk.cpp:99, :109)
    if(src_size != (int32_t)(FSS_FILE_CELLS * 2 * sizeof(fss_tile_record)))
        return fail(FSS_ERR_INVALID, "Chunk src_size was different from expected: %d vs %zu", src_size, FSS_FILE_CELLS * 2 * sizeof(fss_tile_record));
    if(src_size2 != (int32_t)(FSS_FILE_CELLS * sizeof(uint32_t)))
        return fail(FSS_ERR_INVALID, "Chunk src_size2 was different from expected: %d vs %zu", src_size2, FSS_FILE_CELLS * sizeof(uint32_t));
    if(c1 < 0 || c2 < 0 || (size_t)c1 + (size_t)c2 > size - 17) return fail(FSS_ERR_INVALID, "chunk file of %zu bytes is shorter than its two blocks (%d + %d)", size, c1, c2);
    std::vector<fss_tile_record> buf(FSS_FILE_CELLS * 2);
    // (the reference logs these two and carries on with whatever was decoded, :129-133, :172-176; here the caller is told)
    if(fss_lz4::decompress(file + 17, (size_t)c1, (uint8_t*)buf.data(), (size_t)src_size) != src_size) return fail(FSS_ERR_INVALID, "Decompressed chunk tile data is corrupt");
    if(fss_lz4::decompress(file + 17 + c1, (size_t)c2, (uint8_t*)background, (size_t)src_size2) != src_size2)
        return fail(FSS_ERR_INVALID, "Decompressed chunk background data is corrupt");
    memcpy(tiles, buf.data(), FSS_FILE_CELLS * sizeof(fss_tile_record));
    memcpy(layer2, buf.data() + FSS_FILE_CELLS, FSS_FILE_CELLS * sizeof(fss_tile_record));
    if(generation_phase) *generation_phase = (int8_t)file[0];
    return FSS_OK;
}

int fss_download_flow(fss_world* w, int x, int y, int width, int height, float* flow_x, float* flow_y, size_t pitch) {
    if(!w || !flow_x || !flow_y) return fail(FSS_ERR_INVALID, "null argument");
    if(!w->g.fx) return fail(FSS_ERR_STATE, "the world was created without FSS_TRACK_FLOW");
    if(pitch < (size_t)width) return fail(FSS_ERR_INVALID, "pitch %zu < width %d", pitch, width);
    if(use_device(w)) return FSS_ERR_CUDA;
    size_t skip = 0;
    int rc = rect_clip(w, x, y, width, height, skip);
    if(rc) return rc;
    const int rows_per_pass = (int)std::max<size_t>(1, w->stage_bytes / ((size_t)width * 4));
    for(int pass = 0; pass < 2; pass++) {
        const float* plane = pass ? w->g.fy : w->g.fx;
        float* host = (pass ? flow_y : flow_x) + skip * pitch;
        for(int r0 = 0; r0 < height; r0 += rows_per_pass) {
            const int nr = std::min(rows_per_pass, height - r0);
            k_download_plane<<<(width * nr + 255) / 256, 256, 0, w->stream>>>(plane, (float*)w->d_stage, w->g.pitch, w->g.sy0, x, y + r0, width, nr);
            w->stats.kernel_launches++;
            CU(cudaMemcpy2DAsync(host + (size_t)r0 * pitch, pitch * 4, w->d_stage, (size_t)width * 4, (size_t)width * 4, nr, cudaMemcpyDeviceToHost, w->stream));
            CU(cudaStreamSynchronize(w->stream)); // the staging buffer is reused by the next pass
        }
    }
    return FSS_OK;
}

int fss_clear_flow(fss_world* w) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(!w->g.fx) return fail(FSS_ERR_STATE, "the world was created without FSS_TRACK_FLOW");
    if(use_device(w)) return FSS_ERR_CUDA;
    CU(cudaMemsetAsync(w->g.fx, 0, plane_cells(w) * 4, w->stream));
    CU(cudaMemsetAsync(w->g.fy, 0, plane_cells(w) * 4, w->stream));
    return FSS_OK;
}

static void plane_ptrs(fss_world* w, void* ptr[5]);

// Strip-sharded worlds: only the rows a strip owns are current in its planes (the halo exchange refreshes just the ghost rows
// the rules look at), so the sources of its stored rows [s0, s1) that lie outside [own_lo, own_hi) come from the neighbours:
// `u` from the upper strip, `l` from the lower one; `gd` / `gu` = what the lower / upper neighbour needs of this strip's rows.
struct ShiftRows {
    int u_lo, u_hi, l_lo, l_hi;
    int gd_lo, gd_hi, gu_lo, gu_hi;
};
static ShiftRows shift_rows(const fss_world* w, int change_y, bool has_upper, bool has_lower) {
    ShiftRows r = {0, 0, 0, 0, 0, 0, 0, 0};
    const int s0 = w->g.sy0, s1 = w->g.sy0 + w->g.rows, height = (int)w->cfg.height;
    if(has_upper) {
        r.u_lo = std::max(0, s0 - change_y), r.u_hi = w->own_lo;
        r.gu_lo = w->own_lo, r.gu_hi = std::min(height, w->own_lo + FSS_HALO_ROWS - change_y); // the upper strip's storage ends at own_lo + 16
    }
    if(has_lower) {
        r.l_lo = w->own_hi, r.l_hi = std::min(height, s1 - change_y);
        r.gd_lo = std::max(0, w->own_hi - FSS_HALO_ROWS - change_y), r.gd_hi = w->own_hi; // the lower strip's storage starts at own_hi - 16
    }
    if(r.u_hi < r.u_lo) r.u_hi = r.u_lo;
    if(r.l_hi < r.l_lo) r.l_hi = r.l_lo;
    if(r.gu_hi < r.gu_lo) r.gu_hi = r.gu_lo;
    if(r.gd_hi < r.gd_lo) r.gd_hi = r.gd_lo;
    return r;
}

// every plane goes through the spare temperature plane: shift into it, then swap the two pointers.  `foreign`: the five planes of
// the rows r.u of the upper strip, then the five planes of the rows r.l of the lower one (nullptr: an undivided world)
static int shift_planes(fss_world* w, int change_x, int change_y, const uint32_t* foreign, const ShiftRows& r) {
    uint32_t** planes[5] = {&w->g.mf, &w->g.col, (uint32_t**)&w->g.temp, (uint32_t**)&w->g.amt, (uint32_t**)&w->g.dif};
    uint32_t* spare = (uint32_t*)w->temp2;
    const size_t u_words = (size_t)(r.u_hi - r.u_lo) * w->g.pitch, l_words = (size_t)(r.l_hi - r.l_lo) * w->g.pitch;
    ShiftParams p;
    p.pitch = w->g.pitch;
    p.rows = w->g.rows;
    p.width = (int)w->cfg.width;
    p.change_x = change_x;
    p.change_y = change_y;
    p.sy0 = w->g.sy0;
    p.height = (int)w->cfg.height;
    p.own_lo = foreign ? w->own_lo : w->g.sy0;
    p.own_hi = foreign ? w->own_hi : w->g.sy0 + w->g.rows;
    p.u_lo = r.u_lo, p.u_hi = foreign ? r.u_hi : r.u_lo;
    p.l_lo = r.l_lo, p.l_hi = foreign ? r.l_hi : r.l_lo;
    for(int k = 0; k < 5; k++) {
        p.fu = foreign ? foreign + k * u_words : nullptr;
        p.fl = foreign ? foreign + 5 * u_words + k * l_words : nullptr;
        k_shift<<<dim3(w->g.rows, w->g.pitch / 256), 256, 0, w->stream>>>(*planes[k], spare, p);
        w->stats.kernel_launches++;
        std::swap(*planes[k], spare);
    }
    w->temp2 = (int32_t*)spare;
    CU(cudaGetLastError());
    return dirty_chunks(w, 0, 0, -1, -1);
}

int fss_shift_grid(fss_world* w, int change_x, int change_y) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(change_x == 0 && change_y == 0) return FSS_OK;
    if(use_device(w)) return FSS_ERR_CUDA;
    if(w->sharded) {
        // rows cross strip boundaries: collective over the communicator (every rank makes the same call).  Each strip first
        // fetches the rows of its neighbours that its stored rows take their new content from, then shifts.
        if(!w->comm) return fail(FSS_ERR_STATE, "fss_shift_grid on a strip-sharded world needs a communicator (or fss_shift_grid_peer inside one process)");
        const int d = std::abs(change_y);
        if(d > (int)w->cfg.strip_rows - FSS_HALO_ROWS)
            return fail(FSS_ERR_RANGE, "fss_shift_grid: |change_y| = %d exceeds the strip's %u rows less %d (rows would have to skip a strip)", d, w->cfg.strip_rows, FSS_HALO_ROWS);
        const ShiftRows r = shift_rows(w, change_y, w->rank > 0, w->rank + 1 < w->n_ranks);
        uint32_t* foreign = nullptr;
        const size_t u_words = (size_t)(r.u_hi - r.u_lo) * w->g.pitch, l_words = (size_t)(r.l_hi - r.l_lo) * w->g.pitch;
        CU(cudaMalloc(&foreign, std::max<size_t>(1, u_words + l_words) * 5 * sizeof(uint32_t)));
        int rc = FSS_OK;
        {
            void* ptr[5];
            plane_ptrs(w, ptr);
            const int up = w->rank - 1, down = w->rank + 1;
            if(g_nccl.group_start() != 0) rc = fail(FSS_ERR_COMM, "ncclGroupStart");
            for(int k = 0; k < 5 && !rc; k++) {
                auto rows_at = [&](int lo) { return (char*)ptr[k] + (size_t)(lo - w->g.sy0) * w->g.pitch * 4; };
                if(r.gd_hi > r.gd_lo && g_nccl.send(rows_at(r.gd_lo), (size_t)(r.gd_hi - r.gd_lo) * w->g.pitch, NCCL_UINT32, down, w->comm, w->stream) != 0) rc = fail(FSS_ERR_COMM, "ncclSend");
                if(!rc && r.gu_hi > r.gu_lo && g_nccl.send(rows_at(r.gu_lo), (size_t)(r.gu_hi - r.gu_lo) * w->g.pitch, NCCL_UINT32, up, w->comm, w->stream) != 0) rc = fail(FSS_ERR_COMM, "ncclSend");
                if(!rc && u_words && g_nccl.recv(foreign + k * u_words, u_words, NCCL_UINT32, up, w->comm, w->stream) != 0) rc = fail(FSS_ERR_COMM, "ncclRecv");
                if(!rc && l_words && g_nccl.recv(foreign + 5 * u_words + k * l_words, l_words, NCCL_UINT32, down, w->comm, w->stream) != 0) rc = fail(FSS_ERR_COMM, "ncclRecv");
            }
            const std::string first_error = rc ? g_err : std::string(); // (the group is always closed: see fss_exchange_halo)
            const int end = g_nccl.group_end();
            if(rc) fail(rc, "%s", first_error.c_str());
            else if(end != 0) rc = fail(FSS_ERR_COMM, "ncclGroupEnd: %s", g_nccl.err_string ? g_nccl.err_string(end) : "nccl error");
            w->stats.kernel_launches++;
        }
        if(!rc) rc = shift_planes(w, change_x, change_y, foreign, r);
        cudaStreamSynchronize(w->stream);
        cudaFree(foreign);
        return rc;
    }
    if(int rc = shift_planes(w, change_x, change_y, nullptr, ShiftRows{0, 0, 0, 0, 0, 0, 0, 0})) return rc;
    if(w->pl_used) { // world.cpp:2729-2731: the particles move with the grid (what the ticks so far spawned is in the list by then)
        if(int rc = particles_adopt(w)) return rc;
        if(w->pl_n) {
            k_particles_shift<<<(w->pl_n + 255) / 256, 256, 0, w->stream>>>(w->pl_list, w->pl_n, change_x, change_y);
            CU(cudaGetLastError());
            w->stats.kernel_launches++;
        }
    }
    return FSS_OK;
}

// the same for the strips of ONE process (top to bottom), rows fetched by peer copies
int fss_shift_grid_peer(fss_world** strips, int n, int change_x, int change_y) {
    if(!strips || n < 1) return fail(FSS_ERR_INVALID, "null argument");
    for(int k = 0; k < n; k++) {
        if(!strips[k] || !strips[k]->sharded) return fail(FSS_ERR_INVALID, "strip %d is not a strip-sharded world", k);
        if(k && (strips[k - 1]->own_hi != strips[k]->own_lo || strips[k - 1]->g.pitch != strips[k]->g.pitch)) return fail(FSS_ERR_INVALID, "worlds are not adjacent strips");
        if(std::abs(change_y) > (int)strips[k]->cfg.strip_rows - FSS_HALO_ROWS)
            return fail(FSS_ERR_RANGE, "|change_y| = %d exceeds the %u rows of strip %d less %d", std::abs(change_y), strips[k]->cfg.strip_rows, k, FSS_HALO_ROWS);
    }
    if(change_x == 0 && change_y == 0) return FSS_OK;
    // every strip's foreign rows are copied out of the neighbours BEFORE any strip shifts
    std::vector<uint32_t*> foreign((size_t)n, nullptr);
    std::vector<ShiftRows> rows((size_t)n);
    int rc = FSS_OK;
    for(int k = 0; k < n; k++) cudaStreamSynchronize(strips[k]->stream); // what the strips still have queued comes first
    for(int k = 0; k < n && !rc; k++) {
        fss_world* w = strips[k];
        const ShiftRows& r = rows[k] = shift_rows(w, change_y, k > 0, k + 1 < n);
        const size_t u_words = (size_t)(r.u_hi - r.u_lo) * w->g.pitch, l_words = (size_t)(r.l_hi - r.l_lo) * w->g.pitch;
        if(cudaSetDevice(w->device) != cudaSuccess || cudaMalloc(&foreign[k], std::max<size_t>(1, u_words + l_words) * 5 * sizeof(uint32_t)) != cudaSuccess) {
            rc = fail(FSS_ERR_CUDA, "fss_shift_grid_peer: staging rows");
            break;
        }
        for(int side = 0; side < 2 && !rc; side++) {
            const size_t words = side ? l_words : u_words;
            if(!words) continue;
            fss_world* src = strips[side ? k + 1 : k - 1];
            const int lo = side ? r.l_lo : r.u_lo;
            void* ps[5];
            plane_ptrs(src, ps);
            for(int q = 0; q < 5; q++) {
                const char* a = (const char*)ps[q] + (size_t)(lo - src->g.sy0) * src->g.pitch * 4;
                uint32_t* dst = foreign[k] + (side ? 5 * u_words : 0) + q * words;
                if(cudaMemcpyPeerAsync(dst, w->device, a, src->device, words * 4, w->stream) != cudaSuccess) rc = fail(FSS_ERR_CUDA, "peer copy");
            }
        }
        cudaStreamSynchronize(w->stream);
    }
    for(int k = 0; k < n && !rc; k++) {
        if(use_device(strips[k])) rc = FSS_ERR_CUDA;
        else rc = shift_planes(strips[k], change_x, change_y, foreign[k], rows[k]);
    }
    for(int k = 0; k < n; k++)
        if(foreign[k]) {
            cudaSetDevice(strips[k]->device);
            cudaStreamSynchronize(strips[k]->stream);
            cudaFree(foreign[k]);
        }
    return rc;
}

int fss_get_stats(fss_world* w, fss_stats* out) {
    if(!w || !out) return fail(FSS_ERR_INVALID, "null argument");
    *out = w->stats;
    return FSS_OK;
}

// ---- row-strip sharding --------------------------------------------------------------------------------
// Boundary B between an upper strip (rows < B) and the lower strip (rows >= B); B = zone.y + k*32.
//   after tk 0,1 (the chunk row just above B was active: it wrote rows B, B+1 and changed B-2, B-1)
//        upper -> lower : rows [B-2, B+2)
//   after tk 2,3 (the chunk row just below B was active: it wrote rows B-2, B-1 and changed [B, B+18))
//        lower -> upper : rows [B-2, B+10)   (10 = pillar probe depth, world.cpp:1700-1702)
static const int DOWN_LO = -2, DOWN_HI = 2, UP_LO = -2, UP_HI = 10;

static void plane_ptrs(fss_world* w, void* ptr[5]) {
    ptr[0] = w->g.mf;
    ptr[1] = w->g.col;
    ptr[2] = w->g.temp;
    ptr[3] = w->g.amt;
    ptr[4] = w->g.dif;
}

// World::dirty on strip-sharded worlds.  A strip's kernels also mark cells of the two ghost rows next to its boundary (moves reach
// one row, fire ignition two): those marks belong to the neighbour that owns the rows.  They travel with the halo exchange that
// carries the cells they stand for -- after tk 1 the upper strip's marks of rows [B, B+2) go down, after tk 3 the lower strip's
// marks of rows [B-2, B) go up -- are OR-ed into the owner's plane, and are cleared where they came from (sent once: the owner
// may have rendered and cleared its plane before the next exchange).
static const int DIRTY_HALO_ROWS = 2;
static int dirty_halo_buffer(fss_world* w) {
    if(!w->d_dirty_halo) CU(cudaMalloc(&w->d_dirty_halo, (size_t)DIRTY_HALO_ROWS * w->g.pitch));
    return FSS_OK;
}
static uint8_t* dirty_rows(fss_world* w, int lo) { return w->g.dirty + (size_t)(lo - w->g.sy0) * w->g.pitch; }
// after the transfer: the received marks join the owner's, the sent ones are dropped
static int dirty_halo_finish(fss_world* w, bool received, int recv_lo, bool sent, int sent_lo, cudaStream_t st) {
    const size_t bytes = (size_t)DIRTY_HALO_ROWS * w->g.pitch;
    if(received) {
        k_or_words<<<(unsigned)((bytes / 4 + 255) / 256), 256, 0, st>>>((uint32_t*)dirty_rows(w, recv_lo), (const uint32_t*)w->d_dirty_halo, (int)(bytes / 4));
        CU(cudaGetLastError());
        w->stats.kernel_launches++;
    }
    if(sent) CU(cudaMemsetAsync(dirty_rows(w, sent_lo), 0, bytes, st));
    return FSS_OK;
}

int fss_exchange_halo_peer(fss_world* upper, fss_world* lower, int tk) {
    if(!upper || !lower) return fail(FSS_ERR_INVALID, "null world");
    if(upper->own_hi != lower->own_lo || upper->g.pitch != lower->g.pitch) return fail(FSS_ERR_INVALID, "worlds are not adjacent strips");
    if(tk != 1 && tk != 3) return FSS_OK;
    const int B = upper->own_hi;
    fss_world* src = tk == 1 ? upper : lower;
    fss_world* dst = tk == 1 ? lower : upper;
    const int lo = B + (tk == 1 ? DOWN_LO : UP_LO), hi = B + (tk == 1 ? DOWN_HI : UP_HI);
    void *ps[5], *pd[5];
    plane_ptrs(src, ps);
    plane_ptrs(dst, pd);
    // order the copy after the producer's sub-step and before the consumer's next one
    cudaEvent_t ev;
    CU(cudaSetDevice(src->device));
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, src->stream));
    CU(cudaSetDevice(dst->device));
    CU(cudaStreamWaitEvent(dst->stream, ev, 0));
    for(int k = 0; k < 5; k++) {
        const char* s = (const char*)ps[k] + (size_t)(lo - src->g.sy0) * src->g.pitch * 4;
        char* d = (char*)pd[k] + (size_t)(lo - dst->g.sy0) * dst->g.pitch * 4;
        CU(cudaMemcpyPeerAsync(d, dst->device, s, src->device, (size_t)(hi - lo) * src->g.pitch * 4, dst->stream));
    }
    if(dirty_chunks(dst, 0, lo, (int)dst->cfg.width, hi - lo)) return FSS_ERR_CUDA;
    const bool marks = src->g.dirty && dst->g.dirty; // World::dirty: the marks the source left in the rows the destination owns
    const int mlo = tk == 1 ? B : B - DIRTY_HALO_ROWS;
    if(marks) {
        if(int rc = dirty_halo_buffer(dst)) return rc;
        CU(cudaMemcpyPeerAsync(dst->d_dirty_halo, dst->device, dirty_rows(src, mlo), src->device, (size_t)DIRTY_HALO_ROWS * src->g.pitch, dst->stream));
        if(int rc = dirty_halo_finish(dst, true, mlo, false, 0, dst->stream)) return rc;
    }
    CU(cudaEventRecord(ev, dst->stream));
    CU(cudaSetDevice(src->device));
    CU(cudaStreamWaitEvent(src->stream, ev, 0)); // the source must not overwrite the rows while they are copied
    CU(cudaEventDestroy(ev));
    if(marks)
        if(int rc = dirty_halo_finish(src, false, 0, true, mlo, src->stream)) return rc;
    return FSS_OK;
}

int fss_exchange_halo_temperature_peer(fss_world* upper, fss_world* lower) {
    if(!upper || !lower) return fail(FSS_ERR_INVALID, "null world");
    if(upper->own_hi != lower->own_lo || upper->g.pitch != lower->g.pitch) return fail(FSS_ERR_INVALID, "worlds are not adjacent strips");
    const int B = upper->own_hi;
    CU(cudaSetDevice(upper->device));
    CU(cudaStreamSynchronize(upper->stream));
    CU(cudaSetDevice(lower->device));
    CU(cudaStreamSynchronize(lower->stream));
    const size_t rowb = (size_t)upper->g.pitch * 4;
    // upper's rows [B-FSS_HALO_ROWS, B) -> lower's ghost; lower's rows [B, B+FSS_HALO_ROWS) -> upper's ghost
    CU(cudaMemcpyPeer((char*)lower->g.temp + (size_t)(B - FSS_HALO_ROWS - lower->g.sy0) * rowb, lower->device,
                      (char*)upper->g.temp + (size_t)(B - FSS_HALO_ROWS - upper->g.sy0) * rowb, upper->device, FSS_HALO_ROWS * rowb));
    CU(cudaMemcpyPeer((char*)upper->g.temp + (size_t)(B - upper->g.sy0) * rowb, upper->device,
                      (char*)lower->g.temp + (size_t)(B - lower->g.sy0) * rowb, lower->device, FSS_HALO_ROWS * rowb));
    return FSS_OK;
}

int fss_comm_unique_id(void* id128) {
    if(!id128) return fail(FSS_ERR_INVALID, "null argument");
    int rc = nccl_bind();
    if(rc) return rc;
    NC(g_nccl.get_unique_id((NcclId*)id128));
    return FSS_OK;
}

int fss_comm_attach(fss_world* w, const void* id128, int rank, int n_ranks) {
    if(!w || !id128 || rank < 0 || rank >= n_ranks) return fail(FSS_ERR_INVALID, "bad argument");
    int rc = nccl_bind();
    if(rc) return rc;
    if(use_device(w)) return FSS_ERR_CUDA;
    NcclId id;
    memcpy(&id, id128, sizeof id);
    NC(g_nccl.comm_init_rank(&w->comm, n_ranks, id, rank));
    w->rank = rank;
    w->n_ranks = n_ranks;
    recompute_liq(w); // (from here on the strips tell each other which materials they hold: share_present)
    return FSS_OK;
}

// Strip-sharded worlds over a communicator: which kernel specialisation is exact depends on the materials that can be present, and a
// material uploaded into one strip reaches the others through the halo rows.  Once per tick the strips OR their sets (a 256-byte
// all-reduce and a host read: tens of microseconds against ticks of tens of milliseconds); collective like fss_tick itself.
static int share_present(fss_world* w) {
    if(!w->comm) return FSS_OK;
    if(!w->d_present_mask) {
        CU(cudaMalloc(&w->d_present_mask, FSS_MAX_MATERIALS));
        CU(cudaHostAlloc((void**)&w->h_present_mask, FSS_MAX_MATERIALS, cudaHostAllocDefault));
    }
    for(int m = 0; m < FSS_MAX_MATERIALS; m++) w->h_present_mask[m] = (w->present_all || w->present[m]) ? 1 : 0;
    CU(cudaMemcpyAsync(w->d_present_mask, w->h_present_mask, FSS_MAX_MATERIALS, cudaMemcpyHostToDevice, w->stream));
    NC(g_nccl.all_reduce(w->d_present_mask, w->d_present_mask, FSS_MAX_MATERIALS, NCCL_UINT8, NCCL_MAX, w->comm, w->stream));
    CU(cudaMemcpyAsync(w->h_present_mask, w->d_present_mask, FSS_MAX_MATERIALS, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->stats.kernel_launches++;
    bool grew = false;
    for(int m = 0; m < w->n_mat; m++)
        if(w->h_present_mask[m] && !w->present[m]) w->present[m] = grew = true;
    if(grew) recompute_liq(w);
    return FSS_OK;
}

static int nccl_rows(fss_world* w, bool send, int peer, int lo, int hi, bool temp_only, cudaStream_t st) {
    void* ptr[5];
    plane_ptrs(w, ptr);
    const size_t count = (size_t)(hi - lo) * w->g.pitch;
    for(int k = 0; k < 5; k++) {
        if(temp_only && k != 2) continue;
        char* a = (char*)ptr[k] + (size_t)(lo - w->g.sy0) * w->g.pitch * 4;
        if(send) NC(g_nccl.send(a, count, NCCL_UINT32, peer, w->comm, st));
        else NC(g_nccl.recv(a, count, NCCL_UINT32, peer, w->comm, st));
    }
    return FSS_OK;
}

static int exchange_halo_on(fss_world* w, int tk, cudaStream_t st);
static int halo_received(fss_world* w, int tk);

int fss_exchange_halo(fss_world* w, int tk) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(!w->comm) return fail(FSS_ERR_STATE, "no communicator attached");
    if(tk != 1 && tk != 3) return FSS_OK;
    if(use_device(w)) return FSS_ERR_CUDA;
    if(int rc = exchange_halo_on(w, tk, w->stream)) return rc;
    return halo_received(w, tk);
}

// received rows replace cells that settled chunks may have looked at (on the world's stream: after the exchange has been joined)
static int halo_received(fss_world* w, int tk) {
    const bool has_upper = w->rank > 0, has_lower = w->rank + 1 < w->n_ranks;
    if(tk == 1 && has_upper) return dirty_chunks(w, 0, w->own_lo + DOWN_LO, (int)w->cfg.width, DOWN_HI - DOWN_LO);
    if(tk == 3 && has_lower) return dirty_chunks(w, 0, w->own_hi + UP_LO, (int)w->cfg.width, UP_HI - UP_LO);
    return FSS_OK;
}

static int exchange_halo_on(fss_world* w, int tk, cudaStream_t st) {
    const bool has_upper = w->rank > 0, has_lower = w->rank + 1 < w->n_ranks;
    const bool marks = w->g.dirty != nullptr;
    const size_t mark_words = (size_t)DIRTY_HALO_ROWS * w->g.pitch / 4;
    if(marks)
        if(int rc0 = dirty_halo_buffer(w)) return rc0;
    int rc = FSS_OK;
    NC(g_nccl.group_start());
    if(tk == 1) { // data flows down
        if(has_lower) rc = nccl_rows(w, true, w->rank + 1, w->own_hi + DOWN_LO, w->own_hi + DOWN_HI, false, st);
        if(!rc && has_upper) rc = nccl_rows(w, false, w->rank - 1, w->own_lo + DOWN_LO, w->own_lo + DOWN_HI, false, st);
        if(!rc && marks && has_lower && g_nccl.send(dirty_rows(w, w->own_hi), mark_words, NCCL_UINT32, w->rank + 1, w->comm, st) != 0) rc = fail(FSS_ERR_COMM, "ncclSend");
        if(!rc && marks && has_upper && g_nccl.recv(w->d_dirty_halo, mark_words, NCCL_UINT32, w->rank - 1, w->comm, st) != 0) rc = fail(FSS_ERR_COMM, "ncclRecv");
    } else { // data flows up
        if(has_upper) rc = nccl_rows(w, true, w->rank - 1, w->own_lo + UP_LO, w->own_lo + UP_HI, false, st);
        if(!rc && has_lower) rc = nccl_rows(w, false, w->rank + 1, w->own_hi + UP_LO, w->own_hi + UP_HI, false, st);
        if(!rc && marks && has_upper && g_nccl.send(dirty_rows(w, w->own_lo - DIRTY_HALO_ROWS), mark_words, NCCL_UINT32, w->rank - 1, w->comm, st) != 0) rc = fail(FSS_ERR_COMM, "ncclSend");
        if(!rc && marks && has_lower && g_nccl.recv(w->d_dirty_halo, mark_words, NCCL_UINT32, w->rank + 1, w->comm, st) != 0) rc = fail(FSS_ERR_COMM, "ncclRecv");
    }
    // the group is always closed, also after a failed send / recv: an open group would leave the communicator unusable and the
    // neighbour blocked.  A communicator error is fatal for the world (the strips are out of step): destroy and re-create it.
    const std::string first_error = rc ? g_err : std::string();
    const int end = g_nccl.group_end();
    if(rc) return fail(rc, "%s", first_error.c_str());
    if(end != 0) return fail(FSS_ERR_COMM, "ncclGroupEnd: %s", g_nccl.err_string ? g_nccl.err_string(end) : "nccl error");
    w->stats.kernel_launches++; // one fused NCCL send/recv kernel per group
    if(marks) {
        if(tk == 1) rc = dirty_halo_finish(w, has_upper, w->own_lo, has_lower, w->own_hi, st);
        else rc = dirty_halo_finish(w, has_lower, w->own_hi - DIRTY_HALO_ROWS, has_upper, w->own_lo - DIRTY_HALO_ROWS, st);
    }
    return rc;
}

int fss_exchange_halo_temperature(fss_world* w) {
    if(!w) return fail(FSS_ERR_INVALID, "null world");
    if(!w->comm) return fail(FSS_ERR_STATE, "no communicator attached");
    if(use_device(w)) return FSS_ERR_CUDA;
    const bool has_upper = w->rank > 0, has_lower = w->rank + 1 < w->n_ranks;
    int rc = FSS_OK;
    NC(g_nccl.group_start());
    if(has_lower) {
        rc = nccl_rows(w, true, w->rank + 1, w->own_hi - FSS_HALO_ROWS, w->own_hi, true, w->stream);
        if(!rc) rc = nccl_rows(w, false, w->rank + 1, w->own_hi, w->own_hi + FSS_HALO_ROWS, true, w->stream);
    }
    if(!rc && has_upper) {
        rc = nccl_rows(w, true, w->rank - 1, w->own_lo, w->own_lo + FSS_HALO_ROWS, true, w->stream);
        if(!rc) rc = nccl_rows(w, false, w->rank - 1, w->own_lo - FSS_HALO_ROWS, w->own_lo, true, w->stream);
    }
    const std::string first_error = rc ? g_err : std::string(); // (the group is always closed: see fss_exchange_halo)
    const int end = g_nccl.group_end();
    if(rc) return fail(rc, "%s", first_error.c_str());
    if(end != 0) return fail(FSS_ERR_COMM, "ncclGroupEnd: %s", g_nccl.err_string ? g_nccl.err_string(end) : "nccl error");
    w->stats.kernel_launches++;
    return rc;
}

} // extern "C"
