// fss_device.cuh -- device-side data layout of the world.
//
// HBM layout (DESIGN.md "data layout"): structure-of-arrays, five 4-byte planes
//   mf    u32  material id | cached density rank | moved | cached physics type / iterations / is-fire | settleCount
//   color u32  MaterialInstance::color
//   temp  i32  MaterialInstance::temperature   (two planes, ping-pong for the Jacobi stencil)
//   amt   f32  MaterialInstance::fluidAmount
//   dif   f32  MaterialInstance::fluidAmountDiff
// Rows are stored with a pitch that is a multiple of 256 cells, and inside each aligned block of
// 256 columns the columns are PERMUTED so that the 32 columns with the same (x mod 8) are
// contiguous:   perm(x) = (x & ~255) | ((x & 7) << 5) | ((x >> 3) & 31).
// The movement kernel gives lane i of a warp the micro-chunk at columns 8*i + {0..3 | 4..7} of a
// block, so when the lanes of a warp touch "their" column dx of row y they touch ONE 128-byte line.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fss.h"
#include "../../include/fss_rng.h"

#define FSS_BLOCK_COLS 256
#define FSS_HALO_ROWS 16 /* ghost rows stored above / below a strip */

// ---- mf word -----------------------------------------------------------------------------------
#define MF_MAT_MASK 0x000000FFu /* FSS_MAX_MATERIALS = 256 */
#define MF_RANK_MASK 0x0000FF00u /* cached rank of the material's density among the table's distinct densities */
#define MF_RANK_SHIFT 8
#define MF_MOVED 0x00010000u
#define MF_TYPE_SHIFT 17
#define MF_TYPE_MASK 0x000E0000u
#define MF_ITERS_SHIFT 20
#define MF_ITERS_MASK 0x00700000u
#define MF_FIRE 0x00800000u
#define MF_SETTLE_SHIFT 24
#define MF_SETTLE_ONE 0x01000000u
#define MF_DERIVED_MASK (MF_RANK_MASK | MF_TYPE_MASK | MF_ITERS_MASK | MF_FIRE)

__host__ __device__ __forceinline__ uint32_t fss_perm(uint32_t x) {
    return (x & ~255u) | ((x & 7u) << 5) | ((x >> 3) & 31u);
}
__host__ __device__ __forceinline__ uint32_t fss_unperm(uint32_t p) {
    return (p & ~255u) | ((p & 31u) << 3) | ((p >> 5) & 7u);
}
__host__ __device__ __forceinline__ uint32_t mf_mat(uint32_t mf) { return mf & MF_MAT_MASK; }
__host__ __device__ __forceinline__ int mf_type(uint32_t mf) { return (int)((mf & MF_TYPE_MASK) >> MF_TYPE_SHIFT); }
__host__ __device__ __forceinline__ int mf_iters(uint32_t mf) { return (int)((mf & MF_ITERS_MASK) >> MF_ITERS_SHIFT); }
__host__ __device__ __forceinline__ uint32_t mf_settle(uint32_t mf) { return mf >> MF_SETTLE_SHIFT; }
// the (type == T) tests compare the masked field against a constant: one LOP3 + one ISETP
#define MF_IS(mf, T) (((mf) & MF_TYPE_MASK) == ((uint32_t)(T) << MF_TYPE_SHIFT))

// ---- per-material record on the device (built by fss_set_materials) -----------------------------
struct DMat {
    float density;
    uint32_t bits;  // derived mf bits: type, clamped iterations, is-fire
    uint32_t flags; // FSS_MAT_IS_*
    int32_t slip;
    int32_t max_stability; // int(8 / sqrt(slipperyness) + 1), world.cpp:1710, precomputed in double on the host
    int32_t n_react;
    int32_t r_type[FSS_MAX_REACTIONS];
    int32_t r_thr[FSS_MAX_REACTIONS];
    int32_t r_target[FSS_MAX_REACTIONS];
    int32_t create_kind;
    uint32_t create_base;
    uint32_t create_mod;
    uint32_t create_shift;
    uint32_t create_site;
    int32_t create_tex;
    int32_t create_temp;
    uint32_t add_temp;
    float cond_self;
    float cond_other;
};

struct DTex {
    const uint32_t* px;
    int w, h;
};

// hot per-material word kept in shared memory by the movement kernel
//   bits 0..15 slipperyness | 16..23 maxStability (255 = never) | 24..25 n_reactions | 26 is-steam
#define HI_SLIP(i) ((i) & 0xFFFFu)
#define HI_STAB(i) (((i) >> 16) & 0xFFu)
#define HI_NREACT(i) (((i) >> 24) & 3u)
#define HI_STEAM 0x04000000u

// ---- the planes ----------------------------------------------------------------------------------
// Cell index inside one GPU's planes.  64 bits by default (a 65536 x 65536 world on one GPU has 2^32 cells);
// -DFSS_IDX32 was measured and is not faster (profiles/r01_tuning_log.md).
#ifdef FSS_IDX32
typedef uint32_t cidx_t;
#else
typedef size_t cidx_t;
#endif

struct Grid {
    uint32_t* mf;
    uint32_t* col;
    int32_t* temp;
    float* amt;
    float* dif;
    float* fx; // World::flowX / flowY, nullptr unless FSS_TRACK_FLOW
    float* fy;
    uint8_t* dirty; // World::dirty, nullptr unless FSS_TRACK_DIRTY; same indexing as the other planes
    uint32_t pitch; // cells per stored row (multiple of 256)
    int sy0;        // global row of stored row 0
    int rows;       // stored rows
    __device__ __forceinline__ cidx_t idx(int x, int y) const {
        return (cidx_t)(y - sy0) * pitch + fss_perm((uint32_t)x);
    }
};

struct Cell {
    uint32_t mf;
    uint32_t col;
    int32_t temp;
    float amt;
    float dif;
};

__device__ __forceinline__ Cell ld_cell(const Grid& g, cidx_t i) {
    Cell c;
    c.mf = g.mf[i];
    c.col = g.col[i];
    c.temp = g.temp[i];
    c.amt = g.amt[i];
    c.dif = g.dif[i];
    return c;
}
__device__ __forceinline__ void st_cell(const Grid& g, cidx_t i, const Cell& c) {
    g.mf[i] = c.mf;
    g.col[i] = c.col;
    g.temp[i] = c.temp;
    g.amt[i] = c.amt;
    g.dif[i] = c.dif;
}

// host exchange format <-> planes
__host__ __device__ __forceinline__ uint32_t mf_pack(uint32_t material, uint32_t moved, uint32_t settle, uint32_t derived) {
    return (material & MF_MAT_MASK) | (moved ? MF_MOVED : 0u) | (derived & MF_DERIVED_MASK) | (settle << MF_SETTLE_SHIFT);
}
