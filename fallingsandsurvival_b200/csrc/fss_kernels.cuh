// fss_kernels.cuh -- the CUDA kernels of the falling-sand core (sm_100a).
//   k_substep       K1  one (iter, tk) checkerboard sub-step of World::tick        world.cpp:1108-2011
//   k_temperature   K2  World::tickTemperature as a register-tiled 9-point stencil  world.cpp:2043-2143
//   k_upload / k_download   host exchange format (fss_cell, 20 B AoS) <-> permuted SoA planes
//   k_fill          synthetic worlds generated on the device
//   k_rederive / k_hash / k_histogram / k_fluid_totals   bookkeeping
#pragma once
#include "fss_rules.cuh"
#include "fss_rules_liq.cuh"

#ifndef K1_THREADS
#define K1_THREADS 128
#endif
#ifndef K1_MIN_BLOCKS
#define K1_MIN_BLOCKS 7
#endif

// Code-placement pad (see profiles/r01_tuning_log.md "code placement"): a never-launched kernel of K1_CODE_PAD x 16 bytes
// in front of k_substep.  The same k_substep SASS runs iters 0/1 in 13.6 or 15.7 ms depending on where the module places it.
#ifndef K1_CODE_PAD
#define K1_CODE_PAD 0
#endif
#if K1_CODE_PAD > 0
__global__ void k_code_pad(int* out) {
    int v = threadIdx.x;
#pragma unroll
    for(int k = 0; k < K1_CODE_PAD; k++) asm volatile("add.s32 %0, %0, 1;" : "+r"(v));
    if(out) *out = v;
}
#endif

// ---- K1: movement sub-step ---------------------------------------------------------------------------
// grid = (ceil(nx / 128), ny); thread (bx*128 + t, by) owns the micro-chunk at
// (x_begin + 8*(bx*128+t), y_begin + 32*by).  A warp therefore owns 32 micro-chunks that are 8 columns
// apart: with the permuted layout its lanes read consecutive words.
#ifndef K1L_MIN_BLOCKS
#define K1L_MIN_BLOCKS 9
#endif
#ifndef K1N_MIN_BLOCKS
#define K1N_MIN_BLOCKS 7
#endif
// What one warp works on: 32 horizontally adjacent chunks of one active chunk row ("warp tile" wx, by).  In a plain launch
// the warps of CTA (blockIdx.x, blockIdx.y) own the warp tiles 4 * blockIdx.x .. + 3 of chunk row blockIdx.y; with a worklist
// every warp of the fixed grid strides over the list on its own (no CTA-wide step inside the loop), so a list that names one
// live warp tile in every CTA-sized stretch of the world (a 256-cell-wide active pocket) still fills whole CTAs.
#define K1_WARPS (K1_THREADS / 32)
struct TileLoop {
    uint32_t t, n, stride;
    const uint32_t* work;
    __device__ __forceinline__ TileLoop(const SubstepParams& p) : work(p.work) {
        const uint32_t warp = threadIdx.x >> 5;
        t = work ? blockIdx.x * K1_WARPS + warp : 0u;
        n = work ? *p.work_n : 1u;
        stride = gridDim.x * K1_WARPS;
        if(work && blockIdx.x == 0 && threadIdx.x == 0) {
            *p.work_n_host = n;
            *p.work_n_next = 0u;
        }
    }
    // gi = this thread's chunk in its chunk row, by = the chunk row
    __device__ __forceinline__ bool next(int& gi, int& by) {
        if(t >= n) return false;
        const uint32_t e = work ? work[t] : ((uint32_t)blockIdx.y << 14 | (blockIdx.x * K1_WARPS + (threadIdx.x >> 5)));
        gi = (int)(e & 0x3FFFu) * 32 + (int)(threadIdx.x & 31u);
        by = (int)(e >> 14);
        t += work ? stride : 1u;
        return true;
    }
};

template <int MODE, bool FLOW, bool DIRTY>
__global__ void __launch_bounds__(K1_THREADS, MODE == K1_LIQUID ? K1L_MIN_BLOCKS : (MODE == K1_NOGAS ? K1N_MIN_BLOCKS : K1_MIN_BLOCKS))
    k_substep(const __grid_constant__ SubstepParams p) {
    __shared__ uint32_t sinfo[FSS_MAX_MATERIALS];
    for(int m = threadIdx.x; m < p.n_mat; m += K1_THREADS) {
        const DMat& d = p.mats[m];
        uint32_t stab = d.max_stability > 255 ? 255u : (uint32_t)d.max_stability;
        sinfo[m] = ((uint32_t)d.slip & 0xFFFFu) | (stab << 16) | (((uint32_t)d.n_react & 3u) << 24) |
                   ((d.flags & FSS_MAT_IS_STEAM) ? HI_STEAM : 0u);
    }
    __syncthreads();
    TileLoop tiles(p);
    int gi, by;
    while(tiles.next(gi, by)) {
        bool live = gi < p.nx;
        uint8_t* qown = nullptr;
#ifndef K1_NO_EARLY
        if(live && p.q) {
            qown = p.q + (size_t)(2 * by + p.ofs_y + 1) * p.q_w + (2 * gi + p.ofs_x + 1);
            if(*qown <= (uint8_t)p.iter) { // verified settled: a warp whose 32 chunks are all settled has nothing to do
                if(DIRTY) {
                    // the visit that is skipped would have been the identity on the cells, but not on World::dirty: scan 2 marks
                    // every liquid cell it reaches (world.cpp:1823).  The render job adds those marks from this note.
                    uint8_t* s = p.qskip + (qown - p.q);
                    if(*s > (uint8_t)p.iter) *s = (uint8_t)p.iter;
                }
                live = false;
            }
        }
#endif
        if(live) {
            ChunkWalk<MODE, FLOW, DIRTY> w(p, sinfo, p.x_begin + 2 * FSS_CHUNK_W * gi, p.y_begin + 2 * FSS_CHUNK_H * by, __activemask());
            w.run(qown);
        }
    }
}

// ---- K1w: sub-step with the chunk rows in a register window (fss_rules_liq.cuh) -------------------------------------
// Same grid and chunk ownership as k_substep; launched instead of k_substep<K1_LIQUID> (SAND = false) or k_substep<K1_NOGAS>
// (SAND = true) where liq_window_applies().
#ifndef K1W_MIN_BLOCKS
#define K1W_MIN_BLOCKS (K1_NROW_SMEM ? 6 : 5)
#endif
#ifndef K1WS_MIN_BLOCKS
#define K1WS_MIN_BLOCKS 4
#endif
#ifndef K1WG_MIN_BLOCKS
#define K1WG_MIN_BLOCKS 4
#endif
template <bool SAND, bool ITER0, int GEN = 0>
__global__ void __launch_bounds__(K1_THREADS, GEN ? K1WG_MIN_BLOCKS : (SAND ? K1WS_MIN_BLOCKS : K1W_MIN_BLOCKS)) k_substep_win(const __grid_constant__ SubstepParams p) {
    __shared__ uint32_t sinfo[SAND ? FSS_MAX_MATERIALS : 1];
    // WinWalk::pay: colour / temperature of the window rows (SAND), or the two buffers of the row requested ahead (K1_NROW_SMEM)
    __shared__ uint32_t spay[(SAND || K1_NROW_SMEM) ? WIN_SMEM_WORDS(SAND) * K1_THREADS : 1];
    if(SAND) {
        for(int m = threadIdx.x; m < p.n_mat; m += K1_THREADS) {
            const DMat& d = p.mats[m];
            uint32_t stab = d.max_stability > 255 ? 255u : (uint32_t)d.max_stability;
            sinfo[m] = ((uint32_t)d.slip & 0xFFFFu) | (stab << 16) | (((uint32_t)d.n_react & 3u) << 24) |
                       ((d.flags & FSS_MAT_IS_STEAM) ? HI_STEAM : 0u);
        }
        __syncthreads();
    }
    TileLoop tiles(p);
    int gi, by;
    while(tiles.next(gi, by)) {
        bool live = gi < p.nx;
        uint8_t* qown = nullptr;
#ifndef K1_NO_EARLY
        if(live && p.q) {
            qown = p.q + (size_t)(2 * by + p.ofs_y + 1) * p.q_w + (2 * gi + p.ofs_x + 1);
            if(*qown <= (uint8_t)p.iter) live = false; // verified settled
        }
#endif
        if(live) {
            WinWalk<SAND, ITER0, GEN> w(p, sinfo, p.x_begin + 2 * FSS_CHUNK_W * gi, p.y_begin + 2 * FSS_CHUNK_H * by, __activemask(), spay + threadIdx.x,
                                        K1_THREADS);
            w.run(qown);
        }
    }
}

// ---- fire for the window path: one thread per active chunk walks the chunk's active fire cells (fire_prepass_chunk) ----------
// Same chunk ownership and the same settled-chunk test as the kernel that follows; always the plain grid (a settled chunk has no
// active fire cell: a fire cell's visit is never idle).
__global__ void __launch_bounds__(128) k_fire_prepass(const __grid_constant__ SubstepParams p) {
    const int gi = (int)(blockIdx.x * 128u + threadIdx.x), by = (int)blockIdx.y;
    bool live = gi < p.nx;
#ifndef K1_NO_EARLY
    if(live && p.q && p.q[(size_t)(2 * by + p.ofs_y + 1) * p.q_w + (2 * gi + p.ofs_x + 1)] <= (uint8_t)p.iter) live = false;
#endif
    if(!__any_sync(0xFFFFFFFFu, live)) return;
    fire_prepass_chunk(p, p.x_begin + 2 * FSS_CHUNK_W * gi, p.y_begin + 2 * FSS_CHUNK_H * by, live, 0xFFFFFFFFu);
}

// ---- the worklist of a sub-step: which warp tiles (32 chunks of one active chunk row) hold a chunk that is not settled ----
// One THREAD per warp tile (its 32 flags are 64 consecutive bytes); the list slots of a warp's live tiles come from one atomicAdd.
__global__ void __launch_bounds__(128) k_worklist(const uint8_t* __restrict__ q, int q_w, int ofs_x, int ofs_y, int iter, int nx, int tiles_x, int ny,
                                                  uint32_t* __restrict__ work, uint32_t* __restrict__ work_n) {
    const int tile = (int)(blockIdx.x * 128u + threadIdx.x), lane = threadIdx.x & 31;
    bool active = false;
    uint32_t entry = 0u;
    if(tile < tiles_x * ny) {
        const int wx = tile % tiles_x, by = tile / tiles_x;
        const uint8_t* row = q + (size_t)(2 * by + ofs_y + 1) * q_w + (ofs_x + 1) + 2 * (size_t)wx * 32;
        const int n = min(32, nx - wx * 32);
        for(int l = 0; l < n; l++) active = active || row[2 * l] > (uint8_t)iter;
        entry = ((uint32_t)by << 14) | (uint32_t)wx;
    }
    const unsigned live = __ballot_sync(0xFFFFFFFFu, active);
    if(live == 0u) return;
    uint32_t base = 0u;
    if(lane == __ffs((int)live) - 1) base = atomicAdd(work_n, (uint32_t)__popc(live));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs((int)live) - 1);
    if(active) work[base + __popc(live & ((1u << lane) - 1u))] = entry;
}

__global__ void __launch_bounds__(256) k_or_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if(k < n) dst[k] |= src[k];
}

// a few words device -> mapped host memory (or anywhere) without the copy engines
__global__ void k_copy_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int n) {
    for(int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

// ---- K2: temperature stencil ---------------------------------------------------------------------------
// World::tickTemperature, world.cpp:2043-2143.  One thread owns four consecutive columns (half of one "group" x >> 3) and
// walks TEMP_ROWS stored rows top-down; lanes of a warp own consecutive groups.  The per-neighbour terms
//   factor = abs(T)/64.0f * conductionOther;  v += T*factor;  n += factor      (world.cpp:2072-2116)
// depend only on the neighbour, so they are computed once per loaded cell and kept in registers for the three rows
// they take part in; a cell with T == 0 contributes (+0.0f, +0.0f), which leaves the fp32 running sums bit-identical to
// the reference's `if(T != 0)` skip.  The nine terms are added in the reference's order (x offset outer, y offset inner).
// Output goes to the second temperature plane (Jacobi); cells outside tickZone are copied.  How the rows reach the
// threads: see k_temperature below (bulk copies into a shared-memory ring).
#define TEMP_ROWS 64
#define TEMP_THREADS 128
#define K2_COLS 4

struct TempParams {
    Grid g;
    int32_t* out;
    const DMat* mats;
    int n_mat;
    int zx, zy, zw, zh;
    int own_lo, own_hi; // global rows this world computes (owned rows); other stored rows are copied
};

struct TempRaw { // one stored row as loaded: the thread's columns plus one either side
    int32_t t[K2_COLS + 2];
    uint32_t m[K2_COLS + 2];
};
struct TempRow { // the conduction terms of one row, plus what the row needs when it is the centre row
    float f[K2_COLS + 2], tf[K2_COLS + 2];
    int32_t t[K2_COLS];
    uint32_t mpk[K2_COLS / 4]; // material ids of the own columns, one byte each
};

__device__ __forceinline__ void temp_terms(const float* s_co, const TempRaw& r, TempRow& o) {
#pragma unroll
    for(int q = 0; q < K2_COLS / 4; q++) o.mpk[q] = 0u;
#pragma unroll
    for(int c = 0; c < K2_COLS + 2; c++) {
        const int32_t t = r.t[c];
        const uint32_t m = mf_mat(r.m[c]);
        // `if(T != 0) { factor = abs(T) / 64.0f * conductionOther; v += T * factor; n += factor; }` without the test: for
        // T == 0 the two terms come out as zeros (conductionOther is finite: fss_set_materials), and adding a zero of either
        // sign leaves the running sums as the skipped statement does (v and n are never -0: n starts at 0.01, v at +0 and
        // x + (-x) rounds to +0).  float(abs(T)) == fabsf(float(T)): the conversion rounds symmetrically.
        const float ft = (float)t;
        const float fa = fabsf(ft) / 64.0f * s_co[m];
        const float ta = ft * fa;
        o.f[c] = fa;
        o.tf[c] = ta;
        if(c >= 1 && c <= K2_COLS) {
            o.mpk[(c - 1) / 4] |= m << (8 * ((c - 1) & 3));
            o.t[c - 1] = t;
        }
    }
}

// the outputs of the row whose terms are in B (A = the row above, C = the row below)
// s_self[m] = {conductionSelf, 1 - conductionSelf, float(addTemp), addTemp bits}
__device__ __forceinline__ void temp_outputs(const float4* s_self, const TempRow& A, const TempRow& B, const TempRow& C, int32_t* out, bool store = true) {
    // The two running sums of every cell are chains of nine dependent adds each (the reference's order is part of the result):
    // all the chains of the row are written out first, without a branch between them, so that they can be issued
    // interleaved instead of each waiting four cycles for its predecessor.
    float v[K2_COLS], n[K2_COLS];
#pragma unroll
    for(int r = 0; r < K2_COLS; r++) {
        // neighbour order of world.cpp:2072-2116: x offset -1, 0, +1 outer; y offset -1, 0, +1 inner
        float nn = 0.01f, vv = 0.0f;
#pragma unroll
        for(int c = r; c < r + 3; c++) {
            vv += A.tf[c];
            nn += A.f[c];
            vv += B.tf[c];
            nn += B.f[c];
            vv += C.tf[c];
            nn += C.f[c];
        }
        v[r] = vv;
        n[r] = nn;
    }
#pragma unroll
    for(int r = 0; r < K2_COLS; r++) {
        const uint32_t m = (B.mpk[r / 4] >> (8 * (r & 3))) & 0xFFu;
        const float4 self = s_self[m];
        // :2128-2132; addTemp is Uint32 (Material.hpp:40): `addTemp + float` converts it to float.  Both forms are computed and
        // one is selected (no branch: v != 0 almost everywhere a temperature is not zero)
        const int32_t warm = (int32_t)(self.z + (v[r] / n[r] * self.x) + (B.t[r] * self.y));
        const int32_t cold = (int32_t)(__float_as_uint(self.w) + (uint32_t)B.t[r]);
        if(store) out[32 * r] = v[r] != 0 ? warm : cold;
    }
}

// ---- K2, TMA-staged (sm_90+ bulk copies; the shipped form on sm_100a) -------------------------------------------------
// In the permuted
// layout the 512 columns a CTA owns (two aligned 256-column blocks) are 2 KB of CONTIGUOUS words per row and plane, and the
// two halo columns are the words right before and right after them.  So one elected thread fetches a whole row of both planes
// with two `cp.async.bulk` copies of 2 KB + 2 x 16 B into a ring of K2_STAGES shared-memory stages, each guarded by a "full"
// mbarrier (the copy's complete_tx) and an "empty" one (every consumer arrives after it has read the stage); the consumers read
// their 12 words per row from shared memory (lanes hit consecutive banks).  No thread issues a global load in the row loop, no
// register is held for words in flight, and K2_STAGES - 1 rows are on their way while one is computed: the round-2 kernel had
// one row in flight and was waiting for it a third of the time (ncu: long-scoreboard 3.7 per issue at 18 warps per SM).
#define K2_STAGES 3 /* = the period of the three-row register window: stage numbers and window roles are static in the unrolled loop */
#ifndef K2T_MIN_BLOCKS
#define K2T_MIN_BLOCKS 7 /* 72 registers; 5 / 6 / 7 CTAs per SM: 0.855 / 0.851 / 0.818 ms at 16384^2 */
#endif
#define K2_STAGE_WORDS (4 + 512 + 4) /* 16 B of left halo, two 256-column blocks, 16 B of right halo */

__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "K2_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra K2_DONE;\n"
        "bra K2_WAIT;\n"
        "K2_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(TEMP_THREADS, K2T_MIN_BLOCKS) k_temperature(const __grid_constant__ TempParams p) {
    static_assert(K2_COLS == 4 && TEMP_THREADS == 128, "a CTA owns two 256-column blocks: 128 threads x 4 columns");
    __shared__ float s_co[FSS_MAX_MATERIALS];
    __shared__ float4 s_self[FSS_MAX_MATERIALS];
    __shared__ __align__(128) uint32_t s_stage[K2_STAGES][2][K2_STAGE_WORDS]; // [stage][temperature, mf][words]
    __shared__ __align__(8) uint64_t s_full[K2_STAGES], s_empty[K2_STAGES];
    for(int m = threadIdx.x; m < FSS_MAX_MATERIALS; m += TEMP_THREADS) {
        const bool in = m < p.n_mat;
        const float cs = in ? p.mats[m].cond_self : 1.0f;
        const uint32_t add = in ? p.mats[m].add_temp : 0u;
        s_co[m] = in ? p.mats[m].cond_other : 0.0f;
        s_self[m] = make_float4(cs, 1 - cs, (float)add, __uint_as_float(add));
    }
    // thread -> (group of 8 columns, half of the group), as in k_temperature
    const uint32_t slot = blockIdx.x * TEMP_THREADS + threadIdx.x;
    const uint32_t grp = (slot / 64u) * 32u + (slot & 31u);
    const int part = (int)((slot >> 5) & 1u);
    const bool active = grp < p.g.pitch / 8;
    const int x0 = (int)grp * 8 + part * K2_COLS;
    const int sr0 = blockIdx.y * TEMP_ROWS;
    const int sr1 = min(sr0 + TEMP_ROWS, p.g.rows);
    const size_t pitch = p.g.pitch;
    const uint32_t base = fss_perm((uint32_t)x0);
    const uint32_t cta_base = blockIdx.x * 512u; // storage offset of the CTA's first column in a row
    const bool in_zone_x = x0 >= p.zx && x0 + K2_COLS <= p.zx + p.zw;
    // stored rows [c0, c1) of this band are computed (the same for every thread of the CTA), the others copied
    const int y_lo = max(p.zy, p.own_lo), y_hi = min(p.zy + p.zh, p.own_hi);
    const int c0 = max(sr0, y_lo - p.g.sy0), c1 = min(sr1, y_hi - p.g.sy0);
    const bool consumer = active && in_zone_x && c0 < c1;
    if(active) {
        for(int sr = sr0; sr < sr1; sr++) {
            if(consumer && sr >= c0 && sr < c1) continue;
            const size_t rb = (size_t)sr * pitch + base;
#pragma unroll
            for(int r = 0; r < K2_COLS; r++) p.out[rb + r * 32] = p.g.temp[rb + r * 32];
        }
    }
    // Thread 0 issues the copies.  It always walks the row loop like a consumer (its stores are switched off when its own columns
    // lie outside the zone), so that the loop body has no role tests in it.
    const bool producer = threadIdx.x == 0;
    const int n_cons = __syncthreads_count(consumer || producer); // (also orders the table stores above)
    if(__syncthreads_count(consumer) == 0) return;
    if(producer) {
#pragma unroll
        for(int s = 0; s < K2_STAGES; s++) {
            mbar_init(&s_full[s], 1u);
            mbar_init(&s_empty[s], (uint32_t)n_cons);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if(!consumer && !producer) return;

    // row j of the pipeline = stored row c0 - 1 + j, j = 0 .. n_rows - 1 (rows c0-1 and c1 exist: the zone lies FSS_ZONE_MARGIN
    // rows inside the world and a strip stores FSS_HALO_ROWS ghost rows); row j lives in stage j % 3, whose k-th use has parity k & 1
    const int n_rows = c1 - c0 + 2;
    const uint32_t n_blk = min(2u, p.g.pitch / 256u - 2u * blockIdx.x);
    const uint32_t bytes = n_blk * 1024u + 32u;
    const int32_t* src_t = p.g.temp + (size_t)(c0 - 1) * pitch + cta_base - 4;
    const uint32_t* src_m = p.g.mf + (size_t)(c0 - 1) * pitch + cta_base - 4;
    auto issue = [&](int stage, int j) {
        mbar_expect_tx(&s_full[stage], 2u * bytes);
        tma_load_1d(&s_stage[stage][0][0], src_t + (size_t)j * pitch, bytes, &s_full[stage]);
        tma_load_1d(&s_stage[stage][1][0], src_m + (size_t)j * pitch, bytes, &s_full[stage]);
    };
    if(producer) {
        issue(0, 0);
        issue(1, 1);
        if(n_rows > 2) issue(2, 2);
    }
    // word offsets inside a stage: own column k at w0 + 32 k, the columns left / right of the thread's four at wl / wr
    const int w0 = 4 + (int)(base - cta_base);
    const int wl = 4 + (int)(fss_perm((uint32_t)(x0 - 1)) - cta_base), wr = 4 + (int)(fss_perm((uint32_t)(x0 + K2_COLS)) - cta_base);
    // the terms of pipeline row j (in stage `stage`, use parity `par`); the stage is handed back as soon as its words are in
    // registers, and the stage read one row ago (`prev`, parity `ppar`) is sent for row j + 2 once every consumer has arrived there
    auto take = [&](int stage, uint32_t par, int prev, uint32_t ppar, int j, TempRow& o) {
        mbar_wait(&s_full[stage], par);
        TempRaw r;
        const uint32_t* st = &s_stage[stage][0][0];
        const uint32_t* sm = &s_stage[stage][1][0];
        r.t[0] = (int32_t)st[wl];
        r.m[0] = sm[wl];
#pragma unroll
        for(int k = 0; k < K2_COLS; k++) {
            r.t[k + 1] = (int32_t)st[w0 + 32 * k];
            r.m[k + 1] = sm[w0 + 32 * k];
        }
        r.t[K2_COLS + 1] = (int32_t)st[wr];
        r.m[K2_COLS + 1] = sm[wr];
        mbar_arrive(&s_empty[stage]);
        temp_terms(s_co, r, o);
        if(producer && j + 2 < n_rows) {
            mbar_wait(&s_empty[prev], ppar);
            issue(prev, j + 2);
        }
    };
    int32_t* out = p.out + (size_t)c0 * pitch + base;
    TempRow A, B, C;
    {   // rows 0 and 1: nothing to send behind row 0 (stages 0 .. 2 were sent above)
        mbar_wait(&s_full[0], 0u);
        TempRaw r;
        r.t[0] = (int32_t)s_stage[0][0][wl];
        r.m[0] = s_stage[0][1][wl];
#pragma unroll
        for(int k = 0; k < K2_COLS; k++) {
            r.t[k + 1] = (int32_t)s_stage[0][0][w0 + 32 * k];
            r.m[k + 1] = s_stage[0][1][w0 + 32 * k];
        }
        r.t[K2_COLS + 1] = (int32_t)s_stage[0][0][wr];
        r.m[K2_COLS + 1] = s_stage[0][1][wr];
        mbar_arrive(&s_empty[0]);
        temp_terms(s_co, r, A);
    }
    take(1, 0u, 0, 0u, 1, B); // sends row 3 into stage 0
    // loop turn i handles rows j = 2 + 3i (stage 2, use i), j + 1 (stage 0, use i + 1), j + 2 (stage 1, use i + 1)
    int j = 2;
    uint32_t par = 0u;
    for(; j + 2 < n_rows; j += 3, par ^= 1u) {
        take(2, par, 1, par, j, C);
        temp_outputs(s_self, A, B, C, out, consumer);
        take(0, par ^ 1u, 2, par, j + 1, A);
        temp_outputs(s_self, B, C, A, out + pitch, consumer);
        take(1, par ^ 1u, 0, par ^ 1u, j + 2, B);
        temp_outputs(s_self, C, A, B, out + 2 * pitch, consumer);
        out += 3 * pitch;
    }
    if(j < n_rows) {
        take(2, par, 1, par, j, C);
        temp_outputs(s_self, A, B, C, out, consumer);
        if(j + 1 < n_rows) {
            take(0, par ^ 1u, 2, par, j + 1, A);
            temp_outputs(s_self, B, C, A, out + pitch, consumer);
        }
    }
}

// ---- host exchange format <-> planes ---------------------------------------------------------------------
// One CTA converts one row of one aligned 256-column block.  The AoS side is read / written as a
// contiguous run of words through shared memory, the plane side in permuted order: both coalesced.
struct RectParams {
    Grid g;
    const DMat* mats;
    int n_mat;
    uint8_t* present; // upload: present[m] = 1 for every material id seen (which materials can be in the world)
    uint32_t* aos; // rw * rh cells of 5 words, row-major, packed
    int rx, ry, rw, rh; // rectangle in global cell coordinates (already clipped to storage)
};

__global__ void __launch_bounds__(256) k_upload(const __grid_constant__ RectParams p) {
    __shared__ uint32_t sw[256 * 5];
    const int xb = (p.rx & ~255) + 256 * (int)blockIdx.x;
    const int xs = max(xb, p.rx), xe = min(xb + 256, p.rx + p.rw);
    const int row = blockIdx.y;
    const int n = (xe - xs) * 5;
    const uint32_t* src = p.aos + ((size_t)row * p.rw + (xs - p.rx)) * 5;
    for(int k = threadIdx.x; k < n; k += 256) sw[k] = src[k];
    __syncthreads();
    const int x = xb + (int)fss_unperm(threadIdx.x);
    if(x < xs || x >= xe) return;
    const uint32_t* c = sw + (x - xs) * 5;
    const size_t i = p.g.idx(x, p.ry + row);
    const uint32_t w0 = c[0];
    uint32_t mat = w0 & 0xFFFFu;
    if((int)mat >= p.n_mat) mat = 0; // unknown material ids become Tiles::NOTHING's material
    p.present[mat] = 1;
    p.g.mf[i] = mf_pack(mat, (w0 >> 16) & 0xFFu, w0 >> 24, p.mats[mat].bits);
    p.g.col[i] = c[1];
    p.g.temp[i] = (int32_t)c[2];
    p.g.amt[i] = __uint_as_float(c[3]);
    p.g.dif[i] = __uint_as_float(c[4]);
}

__global__ void __launch_bounds__(256) k_download(const __grid_constant__ RectParams p) {
    __shared__ uint32_t sw[256 * 5];
    const int xb = (p.rx & ~255) + 256 * (int)blockIdx.x;
    const int xs = max(xb, p.rx), xe = min(xb + 256, p.rx + p.rw);
    const int row = blockIdx.y;
    const int x = xb + (int)fss_unperm(threadIdx.x);
    if(x >= xs && x < xe) {
        uint32_t* c = sw + (x - xs) * 5;
        const size_t i = p.g.idx(x, p.ry + row);
        const uint32_t mf = p.g.mf[i];
        c[0] = (mf & MF_MAT_MASK) | ((mf & MF_MOVED) ? 0x10000u : 0u) | (mf & 0xFF000000u);
        c[1] = p.g.col[i];
        c[2] = (uint32_t)p.g.temp[i];
        c[3] = __float_as_uint(p.g.amt[i]);
        c[4] = __float_as_uint(p.g.dif[i]);
    }
    __syncthreads();
    const int n = (xe - xs) * 5;
    uint32_t* dst = p.aos + ((size_t)row * p.rw + (xs - p.rx)) * 5;
    for(int k = threadIdx.x; k < n; k += 256) dst[k] = sw[k];
}

// material table changed: refresh the cached type / iterations / fire bits of every stored cell
__global__ void k_rederive(Grid g, const DMat* mats, int n_mat, size_t n) {
    for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t mf = g.mf[i];
        uint32_t mat = mf_mat(mf);
        if((int)mat >= n_mat) mat = 0;
        g.mf[i] = (mf & ~(MF_DERIVED_MASK | MF_MAT_MASK)) | mat | mats[mat].bits;
    }
}

// ---- synthetic worlds (same definition as fallingsandsurvival_b200/worlds.py) -------------------------------
struct FillParams {
    Grid g;
    const DMat* mats;
    const DTex* tex;
    int n_mat;
    uint32_t width, height; // global
    fss_fill_spec spec;
};

__device__ __forceinline__ bool fill_in_box(const fss_fill_spec& s, uint32_t width, uint32_t height, uint32_t x, uint32_t y) {
    // box b has its corner at a position hashed from (seed, b), snapped to 32 cells
    for(int b = 0; b < s.n_boxes; b++) {
        const uint32_t h1 = fss_fmix32(s.seed ^ (0xB0C5u + 2u * (uint32_t)b));
        const uint32_t h2 = fss_fmix32(s.seed ^ (0xB0C6u + 2u * (uint32_t)b) ^ 0x5bd1e995u);
        const uint32_t span_x = width > s.box_size + 2 * s.border ? width - s.box_size - 2 * s.border : 1u;
        const uint32_t span_y = height > s.box_size + 2 * s.border ? height - s.box_size - 2 * s.border : 1u;
        const uint32_t bx = s.border + ((h1 % span_x) & ~31u), by = s.border + ((h2 % span_y) & ~31u);
        if(x - bx < s.box_size && y - by < s.box_size) return true;
    }
    return false;
}

__global__ void __launch_bounds__(256) k_fill(const __grid_constant__ FillParams p) {
    const uint32_t pcol = blockIdx.y * 256u + threadIdx.x; // stored (permuted) column
    const uint32_t sr = blockIdx.x;                        // stored row (grid.x: a 65536-row world exceeds gridDim.y)
    if(pcol >= p.g.pitch) return;
    const uint32_t x = fss_unperm(pcol), y = (uint32_t)p.g.sy0 + sr;
    const size_t i = (size_t)sr * p.g.pitch + pcol;
    const fss_fill_spec& s = p.spec;
    int mat = 0;
    const uint32_t key = fss_rng_cell_key(s.seed, x, y);
    if(x < p.width) {
        const bool boxed = s.n_boxes != 0 && fill_in_box(s, p.width, p.height, x, y);
        if(x < s.border || y < s.border || x >= p.width - s.border || y >= p.height - s.border) {
            mat = s.border_material;
        } else if(s.solid_from_row && y >= s.solid_from_row && !boxed) { // (a box inside the solid region is a pocket of the mix)
            mat = s.solid_material;
        } else {
            const uint32_t r = key % 100u;
            if(s.n_boxes == 0 || boxed) {
                for(int e = 0; e < s.n_entries; e++) {
                    if(r < (uint32_t)s.cumulative_percent[e]) {
                        mat = s.material[e];
                        break;
                    }
                }
            } else if((key >> 8) % 1000u < s.sparse_per_mille) {
                mat = s.sparse_material;
            }
        }
    }
    if(mat < 0 || mat >= p.n_mat) mat = 0;
    const DMat& d = p.mats[mat];
    uint32_t col;
    if(d.create_kind == FSS_CREATE_RAND) {
        col = d.create_base + ((fss_rand(key, d.create_site, 0) % d.create_mod) << d.create_shift);
    } else if(d.create_kind == FSS_CREATE_TEXTURE && d.create_tex >= 0 && p.tex[d.create_tex].px) {
        const DTex t = p.tex[d.create_tex];
        col = t.px[(y % t.h) * t.w + (x % t.w)];
    } else {
        col = d.create_base;
    }
    p.g.mf[i] = (uint32_t)mat | d.bits;
    p.g.col[i] = col;
    p.g.temp[i] = d.create_temp;
    p.g.amt[i] = 2.0f;
    p.g.dif[i] = 0.0f;
}

// ---- grid shift when the camera moves: World::tickChunks, world.cpp:2617-2630 ---------------------------------
// tiles[new] = tiles[new - change] wherever the source lies inside the grid; every other cell keeps its content
// (the reference's in-place loop never clears the vacated band).  One plane per launch, out of place.
// Strip-sharded worlds: `rows` stored rows starting at global row sy0 of a world `height` rows high.  Only the rows a strip OWNS
// are current in its planes (of its ghost rows the halo exchange refreshes the 2 above and the 10 below that the rules look
// at), so a source row outside [own_lo, own_hi) comes from the neighbour that owns it: `fu` = rows [u_lo, u_hi) of the upper
// strip, `fl` = rows [l_lo, l_hi) of the lower one, fetched before the shift (fss_abi.cu).  Every stored row, ghost rows
// included, is current afterwards.
struct ShiftParams {
    uint32_t pitch;
    int rows, width, change_x, change_y, sy0, height, own_lo, own_hi;
    const uint32_t *fu, *fl;
    int u_lo, u_hi, l_lo, l_hi;
};
__global__ void __launch_bounds__(256) k_shift(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const ShiftParams p) {
    const uint32_t pcol = blockIdx.y * 256u + threadIdx.x;
    const int y = blockIdx.x;
    if(pcol >= p.pitch) return;
    const int x = (int)fss_unperm(pcol);
    const int ox = x - p.change_x, oy = p.sy0 + y - p.change_y; // global source row
    const size_t i = (size_t)y * p.pitch + pcol;
    uint32_t v = in[i];
    if(x < p.width && ox >= 0 && ox < p.width && oy >= 0 && oy < p.height) {
        const uint32_t po = fss_perm((uint32_t)ox);
        if(oy >= p.own_lo && oy < p.own_hi) v = in[(size_t)(oy - p.sy0) * p.pitch + po];
        else if(oy >= p.u_lo && oy < p.u_hi) v = p.fu[(size_t)(oy - p.u_lo) * p.pitch + po];
        else if(oy >= p.l_lo && oy < p.l_hi) v = p.fl[(size_t)(oy - p.l_lo) * p.pitch + po];
    }
    out[i] = v;
}

// ---- chunk records <-> planes (Chunk.cpp:139-156, :189-281; world.cpp:2526-2546) -----------------------------
struct ChunkParams {
    Grid g;
    const DMat* mats;
    int n_mat;
    uint32_t* recs; // chunk_w * chunk_h records of 3 words {index (low 16 bits), color, temperature}
    int x0, y0, cw, ch;
    int width, y_lo, y_hi; // world width; stored rows [y_lo, y_hi)
    uint8_t* present;
};

__global__ void __launch_bounds__(256) k_merge_records(const __grid_constant__ ChunkParams p) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if(k >= p.cw * p.ch) return;
    const int x = p.x0 + k % p.cw, y = p.y0 + k / p.cw;
    if(x < 0 || x >= p.width || y < p.y_lo || y >= p.y_hi) return;
    const uint32_t* r = p.recs + 3 * (size_t)k;
    uint32_t mat = r[0] & 0xFFFFu;
    if((int)mat >= p.n_mat) mat = 0;
    const size_t i = p.g.idx(x, y);
    p.g.mf[i] = mat | p.mats[mat].bits;
    p.g.col[i] = r[1];
    p.g.temp[i] = (int32_t)r[2];
    p.g.amt[i] = 2.0f; // MaterialInstance.hpp:21
    p.g.dif[i] = 0.0f;
    p.present[mat] = 1;
}

__global__ void __launch_bounds__(256) k_read_records(const __grid_constant__ ChunkParams p) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if(k >= p.cw * p.ch) return;
    const int x = p.x0 + k % p.cw, y = p.y0 + k / p.cw;
    uint32_t* r = p.recs + 3 * (size_t)k;
    if(x < 0 || x >= p.width || y < p.y_lo || y >= p.y_hi) {
        r[0] = r[1] = r[2] = 0u;
        return;
    }
    const size_t i = p.g.idx(x, y);
    r[0] = mf_mat(p.g.mf[i]);
    r[1] = p.g.col[i];
    r[2] = (uint32_t)p.g.temp[i];
}

// one float plane of a rectangle, permuted storage -> row-major staging
__global__ void __launch_bounds__(256) k_download_plane(const float* __restrict__ plane, float* __restrict__ out, uint32_t pitch, int sy0,
                                                        int rx, int ry, int rw, int rh) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if(k >= rw * rh) return;
    const int x = rx + k % rw, y = ry + k / rw;
    out[k] = plane[(size_t)(y - sy0) * pitch + fss_perm((uint32_t)x)];
}

// ---- reductions over the owned rows ---------------------------------------------------------------------
struct ReduceParams {
    Grid g;
    uint32_t width;
    int own_lo, own_hi; // global rows
    int n_mat;
    unsigned long long* out_u64; // hash: [0]; histogram: [n_mat]
    double* out_f64;             // fluid totals: [n_mat]
    const DMat* mats;
};

__global__ void __launch_bounds__(256) k_hash(const __grid_constant__ ReduceParams p) {
    unsigned long long sum = 0;
    const size_t rows = (size_t)(p.own_hi - p.own_lo);
    const size_t n = rows * p.g.pitch;
    for(size_t k = blockIdx.x * (size_t)256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) {
        const uint32_t pcol = (uint32_t)(k % p.g.pitch);
        const uint32_t x = fss_unperm(pcol);
        if(x >= p.width) continue;
        const int y = p.own_lo + (int)(k / p.g.pitch);
        const size_t i = (size_t)(y - p.g.sy0) * p.g.pitch + pcol;
        const uint32_t mf = p.g.mf[i];
        sum += fss_cell_digest(x, (uint32_t)y, mf & ~MF_DERIVED_MASK, p.g.col[i], (uint32_t)p.g.temp[i], __float_as_uint(p.g.amt[i]),
                               __float_as_uint(p.g.dif[i]));
    }
    for(int o = 16; o; o >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, o);
    __shared__ unsigned long long s[8];
    if((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = sum;
    __syncthreads();
    if(threadIdx.x == 0) {
        unsigned long long t = 0;
        for(int q = 0; q < 8; q++) t += s[q];
        atomicAdd(p.out_u64, t);
    }
}

__global__ void __launch_bounds__(256) k_histogram(const __grid_constant__ ReduceParams p) {
    __shared__ unsigned int h[FSS_MAX_MATERIALS];
    __shared__ double fl[FSS_MAX_MATERIALS];
    for(int m = threadIdx.x; m < FSS_MAX_MATERIALS; m += 256) {
        h[m] = 0;
        fl[m] = 0.0;
    }
    __syncthreads();
    const size_t rows = (size_t)(p.own_hi - p.own_lo);
    const size_t n = rows * p.g.pitch;
    // bounded trip count per block keeps the 32-bit shared counters from overflowing
    for(size_t k = blockIdx.x * (size_t)256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) {
        const uint32_t pcol = (uint32_t)(k % p.g.pitch);
        if(fss_unperm(pcol) >= p.width) continue;
        const size_t i = ((size_t)(p.own_lo - p.g.sy0) + k / p.g.pitch) * p.g.pitch + pcol;
        const uint32_t mf = p.g.mf[i];
        const uint32_t m = mf_mat(mf);
        atomicAdd(&h[m], 1u);
        if(p.out_f64 && MF_IS(mf, FSS_PHYS_SOUP)) atomicAdd(&fl[m], (double)p.g.amt[i] + (double)p.g.dif[i]);
    }
    __syncthreads();
    for(int m = threadIdx.x; m < p.n_mat; m += 256) {
        if(h[m]) atomicAdd(&p.out_u64[m], (unsigned long long)h[m]);
        if(p.out_f64 && fl[m] != 0.0) atomicAdd(&p.out_f64[m], fl[m]);
    }
}
