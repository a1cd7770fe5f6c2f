// fss_rules_liq.cuh -- sub-steps of World::tick with the chunk rows held in REGISTERS (round 2).
//
// Which sub-steps: the ones k_substep<K1_LIQUID> serves (only SOUP cells can act) and the ones k_substep<K1_NOGAS> serves
// (SAND and SOUP act, no GAS / fire material present), without FSS_PARTICLES_EMIT (the free-fall tests of world.cpp:1288
// and :1352 look four rows down) and without flow / dirty planes.  liq_window_applies() is the one definition of that
// condition; everything else keeps running ChunkWalk (fss_rules.cuh).
//
// Why: ChunkWalk walks its 64 cells through L1: every liquid cell visit re-loads the mf / amt / dif words of its four
// neighbours (13 loads), each behind 64-bit index arithmetic, and scan 2 walks the chunk a second time.  ncu (round 1):
// 448 warp instructions per chunk position, a third of them address arithmetic.  Here a thread keeps a sliding window of
// three chunk rows (the row being scanned, the one above, the one below: mf, amt, dif of its four columns) in registers,
// with the column loop unrolled so that every neighbour is a register with a static name:
//   * every word of the chunk is loaded ONCE (12 loads per row + 6 for the two outside columns) and stored at most once,
//     with constant offsets from three row pointers; the loads of row r-2 are requested before row r is scanned, so a
//     whole row of arithmetic covers their latency;
//   * scan 2's liquid step (`fluidAmount += fluidAmountDiff`, world.cpp:1808-1825) is applied to row r+1 in registers right
//     after scan 1 has finished row r, on its way back to memory.
// Exactness of the fused order (same argument as ChunkWalk::early_apply_row, here for ALL of scan 2 because nothing but
// SOUP acts in these sub-steps):
//   * scan 1 of row q reads and writes rows q-1 .. q+1 only (flows go one cell up / down / sideways; un-settling touches
//     the four neighbours), so once row r is done no scan-1 visit looks at row r+1 again;
//   * a scan-2 visit of a SOUP cell touches that cell only, and only unvisited SOUP cells are visited: the order among
//     scan-2 visits does not matter, and nothing of scan 2 is read by the scan-1 visits that follow it (rows <= r-1 ... they
//     look at rows <= r).
// Window registers are the truth for rows r-1 .. r+1 of the own four columns while they are in the window; the outside
// columns (one cell either side of the row being scanned) and the outside rows (-1 and 16) are written through.
#pragma once
#include "fss_rules.cuh"
#ifdef __CUDACC__
#include <cuda_pipeline.h>
#endif

// the one place that decides which sub-steps take the register-window path (kernel launch and tests/hostsim)
//   K1_LIQUID sub-steps: WinWalk<false>, scans 1 and 2 fused completely;
//   K1_NOGAS sub-steps (SAND + SOUP act): WinWalk<true> runs scan 1 (SAND falling straight, world.cpp:1276-1335, and the
//   liquid rule) and the liquid half of scan 2 on the window, then ChunkWalk<K1_NOGAS>::scan2_rest() runs scan 2's SAND rule
//   through L1 as before.  `sand_reacts`: a SAND material with reactions (world.cpp:1251-1274: reads the temperature plane and
//   re-creates the cell) can be present -- those worlds stay on ChunkWalk.
//   K1_GENERAL sub-steps and SAND materials with reactions (round 2, second half): WinWalk<true, *, GEN = true> adds GAS (rising
//   in scan 1 is an exchange with the row above; scans 2 and 3 by the visited bits) and SAND reactions (the temperature travels
//   with the cell in the shared-memory payload).  Fire: its rule is NOT in the window kernel (built there it was exact, but the
//   kernel grew to 105 KB of code and ran c3's iter 0 in 27.1 ms against 21.9 ms; profiles/r02_tuning_log.md section 10); where
//   the material table allows it (`fire_static`) the fire cells' visits are decided by k_fire_prepass before the kernel
//   (fire_prepass_chunk below), otherwise sub-steps in which fire acts stay on ChunkWalk<K1_GENERAL>.
// `fire_iters`: the largest `iterations` of a fire material that can be present (0: none); `fire_static`: the conditions under which
// the fire rule can be split off into k_fire_prepass hold (fire_prepass_chunk below)
__host__ __device__ __forceinline__ bool liq_window_applies(int mode, int iter, uint32_t cfg_flags, bool flow, bool dirty, bool sand_reacts = false,
                                                            int fire_iters = 0, bool fire_static = false) {
#ifdef K1_NO_LIQ_WINDOW
    return false;
#else
    if((cfg_flags & FSS_PARTICLES_EMIT) || flow || dirty) return false;
#ifdef K1_NO_SAND_WINDOW
    return mode == K1_LIQUID;
#elif defined(K1_NO_GEN_WINDOW)
    return mode == K1_LIQUID || (mode == K1_NOGAS && !sand_reacts);
#else
    return mode != K1_GENERAL || iter >= fire_iters || fire_static;
#endif
#endif
}
// the general form of the window walk is needed (else WinWalk<true, *, false> is exact)
__host__ __device__ __forceinline__ bool liq_window_general(int mode, bool sand_reacts) { return mode == K1_GENERAL || (mode == K1_NOGAS && sand_reacts); }

// x / 3.0f and x / 0.6f, correctly rounded, without the generic division's reciprocal + slow path: q = RN(x * RN(1/c)),
// r = x - c*q exactly (fma), RN(q + r * RN(1/c)).  Checked against `/` for every float (tests/test_liq_window.py restates
// the check on a sample): identical for all normal-range operands -- x/3 for every finite x except the sign of -0 / 3 --
// and the liquid rule only uses them on differences and sums of amounts (|x| between 1e-9 and 16, or exactly +0).
__host__ __device__ __forceinline__ float fdiv_const(float x, float c, float rc) {
    const float q = x * rc;
#ifdef __CUDA_ARCH__
    const float r = __fmaf_rn(-c, q, x);
    return __fmaf_rn(r, rc, q);
#else
    const float r = fmaf(-c, q, x);
    return fmaf(r, rc, q);
#endif
}

// ---- FIRE for the register-window path, world.cpp:1171-1216 -----------------------------------------------------------------
// The fire rule does not fit the window kernel (its code made the kernel a third longer and a quarter slower, tuning log 10), and
// it does not have to be there.  What a fire cell's visit does: draw (RNG keys of the cell), look which of its 25 neighbours are
// SOLID, set some of those alight (`tiles[...] = Tiles::createFire()`, marked visited), go out.  Under two conditions on the
// materials that can be present (fss_abi.cu, `fire_static`):
//   (a) no SAND material is denser than a fire material -- then nothing ever moves a fire cell (SAND displaces what is lighter,
//       liquids flow into AIR or their own kind, GAS rises into AIR), and no movement rule can tell a SOLID cell from a fire cell
//       (neither can be moved into);
//   (b) no SAND reaction produces a SOLID material -- then SOLID cells only change by being set alight;
// which neighbours are SOLID at the time of the visit depends on the grid at the START of the sub-step and on the ignitions of the
// fire cells of the same chunk that scan 1 reaches earlier, and WHEN within the sub-step a SOLID cell turns into fire makes no
// difference to anybody but later fire cells of the same chunk.  So one thread per active chunk walks the chunk's active fire cells
// in scan-1 order BEFORE the sub-step's kernel: it sets the chosen neighbours alight in memory right away (its later gathers
// see them), notes FIRE_NEW for those inside the chunk (visited: they do not act when scan 1 reaches them) and FIRE_DIES for a
// fire cell that goes out.  Going out has to happen at the cell's own visit (cells scanned later see AIR, earlier ones saw
// fire): the window kernel does that from the note.  Without FSS_PARTICLES_EMIT the recolour and the particle draws
// (:1172-1189) have no effect.
// `live`: this lane has a chunk to walk (every lane of the warp goes through the loops: the 25 neighbours of a fire cell are
// gathered by the whole warp, lane q loading neighbour q, like ChunkWalk::coop_fire_scan); `wmask`: the lanes of the warp.
__device__ __forceinline__ void fire_prepass_chunk(const SubstepParams& p, int cx, int cy, bool live, unsigned wmask) {
    const Grid& g = p.g;
    const int lane = (int)(threadIdx.x & 31u);
    const int rank = __popc(wmask & ((1u << lane) - 1u)), n = __popc(wmask);
#pragma unroll 1
    for(int r = FSS_CHUNK_H - 1; r >= 0; r--) {
        const cidx_t irow = live ? (cidx_t)(cy + r - g.sy0) * g.pitch + fss_perm((uint32_t)cx) : 0;
        uint32_t m[FSS_CHUNK_W];
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) m[k] = live ? g.mf[irow + 32 * k] : 0u;
#pragma unroll 1
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            const uint32_t mf = k == 0 ? m[0] : (k == 1 ? m[1] : (k == 2 ? m[2] : m[3]));
            const cidx_t i = irow + 32 * k;
            const int x = cx + k, y = cy + r;
            // an active fire cell that was not set alight a moment ago by this very thread
            bool fire = live && (mf & MF_FIRE) && p.iter < mf_iters(mf);
            if(fire) fire = p.fire_verdict[i] != FIRE_NEW;
            unsigned fire_lanes = __ballot_sync(wmask, fire);
            if(!fire_lanes) continue;
            // bit q = the neighbour (q/5 - 2, q%5 - 2) is SOLID, in the reference's loop order (xx outer, yy inner)
            uint32_t solid_nb = 0u;
            while(fire_lanes) {
                const int src = __ffs((int)fire_lanes) - 1;
                fire_lanes &= fire_lanes - 1;
                const int fx = __shfl_sync(wmask, x, src), fy = __shfl_sync(wmask, y, src);
                uint32_t part = 0u;
                for(int q = rank; q < 25; q += n)
                    if(MF_IS(g.mf[g.idx(fx + q / 5 - 2, fy + q % 5 - 2)], FSS_PHYS_SOLID)) part |= 1u << q;
                const uint32_t all = __reduce_or_sync(wmask, part);
                if(lane == src) solid_nb = all;
            }
            if(fire) {
                const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
                bool dies = fss_rand(key, 1191, 0) % 150 == 0; // :1191-1196
                if(!dies) {
                    uint32_t k1202 = 0u, kcreate = 0u;
                    for(uint32_t rest = solid_nb; rest; rest &= rest - 1) { // :1198-1208
                        if(fss_rand(key, 1202, k1202++) % 500 == 0) {
                            const int q = __ffs((int)rest) - 1, xx = q / 5 - 2, yy = q % 5 - 2;
                            const cidx_t j = g.idx(x + xx, y + yy);
                            st_cell(g, j, create_cell(p, p.fire_id, x + xx, y + yy, key, kcreate++)); // Tiles::createFire()
                            const unsigned dx = (unsigned)(x + xx - cx), dy = (unsigned)(y + yy - cy);
                            if(dx < (unsigned)FSS_CHUNK_W && dy < (unsigned)FSS_CHUNK_H) p.fire_verdict[j] = FIRE_NEW; // :1205, inside the chunk
                        }
                    }
                    dies = !solid_nb && fss_rand(key, 1210, 0) % 120 == 0; // :1210-1214
                }
                if(dies) p.fire_verdict[i] = FIRE_DIES;
            }
            __syncwarp(wmask); // what this lane has just set alight is read by the other lanes on its behalf in its later gathers
        }
    }
}

struct LRow { // one grid row of the chunk's four columns
    uint32_t mf[FSS_CHUNK_W];
    float amt[FSS_CHUNK_W];
    float dif[FSS_CHUNK_W];
    // bits 0-3: mf differs from memory; 4-7: dif differs; 8-11: amt differs; 16-19: tickVisited (world.cpp:1161-1166)
    uint32_t fl;
};
#define LF_MF 0x1u
#define LF_DIF 0x10u
#define LF_AMT 0x100u
#define LF_ALL 0x111u
#define LF_PAY 0x1000u /* bits 12-15: the colour / temperature words in shared memory differ from memory (SAND) */
#define LF_VIS 0x10000u
#define WIN_PAY_WORDS (3 * 2 * FSS_CHUNK_W) /* per thread: 3 window rows x {colour, temperature} x 4 columns */
#ifndef K1_NROW_SMEM
#define K1_NROW_SMEM 1 /* liquid-only kernels: the row requested ahead waits in shared memory (cp.async) instead of in 13 registers */
#endif
#define WIN_NB_WORDS (2 * 3 * FSS_CHUNK_W) /* per thread: two buffers of {mf, amt, dif} x 4 columns */
#define WIN_SMEM_WORDS(SAND_) ((SAND_) ? WIN_PAY_WORDS : WIN_NB_WORDS)

struct LSide { // the cell left / right of the row being scanned (outside the chunk)
    uint32_t mf;
    float amt, dif;
};

// SAND = false: liquid-only sub-steps, scans 1 and 2 fused completely.
// SAND = true:  SAND acts too: its scan-1 rule (falling straight down, world.cpp:1276-1335) runs on the window as well;
//               cells that scan 2's liquid step would delete are left to scan 2 proper (they change type, which a SAND
//               cell to their left tests first: ChunkWalk::early_apply_row), and `vis` is handed to ChunkWalk::scan2_rest().
// ITER0: the sub-step is an iter-0 one: only then can two different liquids change places (world.cpp:1406-1412, :1510-1516);
//        compiled out of the other five iters' kernels (their code is a tenth shorter: instruction-cache misses are a
//        measurable part of this kernel's stalls)
// GEN:   GAS and SAND reactions can be present (SAND must be true); fire materials too, as long as none of them can act in the
//        sub-step (liq_window_applies): a spent fire cell is an inert PASSABLE cell.
// GEN is a level: 0 = SAND + SOUP only; 1 = GAS and SAND reactions can be present but no GAS material can act in scan 1 any more
// (iter >= its `iterations`: GAS then only moves in scans 2 / 3, where it does whatever its `iterations`); 2 = everything.
// Level 1 exists because the kernel's speed follows its code size: it is 7 KB shorter than level 2.
template <bool SAND, bool ITER0 = true, int GEN = 0>
struct WinWalk {
    static_assert(SAND || !GEN, "the general window walk includes the SAND rules");
    static constexpr bool FIRE = GEN && ITER0; // the instantiation that also serves the sub-steps in which fire acts (fire_prepass_chunk)
    const SubstepParams& p;
    const uint32_t* sinfo;
    int cx, cy;
    // word offsets relative to the chunk's column 0: the own columns are at +32*k (one 8-column group), the outside columns at
    // lofs / rofs (fss_device.cuh, permuted layout)
    int lofs, rofs;
    cidx_t c0; // index of (cx, cy) in the planes
    bool busy; // this visit stored to the grid or drew a random number (else it was the identity); see SubstepParams::q
    unsigned long long vis; // SAND: tickVisited of the chunk as scan 2 finds it
    unsigned long long stopm; // SAND: unvisited SAND cells with moved == false (ChunkWalk::scan2_sparse decides their pillar draw without a load)
    unsigned long long gasm;  // GEN: positions holding a GAS cell when scan 1 was done with them (ChunkWalk::scan3_sparse)
    uint32_t lb_mf, rb_mf;  // SAND: mf of the outside cells of the row BELOW the one being scanned (a SAND cell's diagonals)
    unsigned wmask;
    // SAND: cells move, and their colour / temperature words move with them.  Those two planes of the three window rows live in
    // SHARED memory (requested with cp.async one row ahead, no registers, no global round trip in the middle of a row):
    // word (slot, plane, column) of this thread at pay[((slot * 2 + plane) * 4 + column) * pstride].  sA / sC / sB = the slots of
    // rows r-1 / r / r+1.  (Liquid-only sub-steps copy the two words through global memory in the rare cases they need them.)
    uint32_t* pay;
    int pstride;
    int sA, sC, sB;

    __device__ __forceinline__ WinWalk(const SubstepParams& p_, const uint32_t* si, int cx_, int cy_, unsigned wmask_, uint32_t* pay_ = nullptr,
                                       int pstride_ = 1)
        : p(p_), sinfo(si), cx(cx_), cy(cy_), busy(false), vis(0ull), stopm(0ull), gasm(0ull), lb_mf(0u), rb_mf(0u), wmask(wmask_), pay(pay_), pstride(pstride_), sA(0), sC(1), sB(2) {
        const uint32_t pb = fss_perm((uint32_t)cx);
        lofs = (int)fss_perm((uint32_t)(cx - 1)) - (int)pb;
        rofs = (int)fss_perm((uint32_t)(cx + FSS_CHUNK_W)) - (int)pb;
        c0 = (cidx_t)(cy - p.g.sy0) * p.g.pitch + pb;
    }

    // index of column 0 of chunk row r (r = -1 .. 16)
    __device__ __forceinline__ cidx_t row_index(int r) const { return c0 + (cidx_t)((long long)r * (long long)p.g.pitch); }

    __device__ __forceinline__ void load_row(LRow& R, cidx_t i) const {
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            R.mf[k] = p.g.mf[i + 32 * k];
            R.amt[k] = p.g.amt[i + 32 * k];
            R.dif[k] = p.g.dif[i + 32 * k];
        }
        R.fl = 0u;
    }
    __device__ __forceinline__ uint32_t& P(int slot, int plane, int k) const { return pay[((slot * 2 + plane) * FSS_CHUNK_W + k) * pstride]; }
    // colour / temperature of one grid row (column 0 at index i) into a slot: asynchronous on the device
    __device__ __forceinline__ void fetch_payload(int slot, cidx_t i) const {
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
#ifdef __CUDA_ARCH__
            __pipeline_memcpy_async(&P(slot, 0, k), p.g.col + i + 32 * k, 4);
            __pipeline_memcpy_async(&P(slot, 1, k), p.g.temp + i + 32 * k, 4);
#else
            P(slot, 0, k) = p.g.col[i + 32 * k];
            P(slot, 1, k) = (uint32_t)p.g.temp[i + 32 * k];
#endif
        }
#ifdef __CUDA_ARCH__
        __pipeline_commit();
#endif
    }
    // !SAND with K1_NROW_SMEM: the row requested one iteration ahead (mf, amt, dif of the four columns) lands in one of two
    // buffers of the thread's shared-memory slice instead of in the registers of `N`: 96 -> 80 registers, six CTAs per SM
    // instead of five, 3.24 -> 3.00 ms per liquid iter of c2.  (The same in the SAND kernel, with five CTAs instead of four, is
    // slower: 7.8 -> 8.05 ms; the outside cells through shared memory as well change nothing.  profiles/r02_tuning_log.md 15)
    static constexpr bool NSM = !SAND && K1_NROW_SMEM != 0;
    __device__ __forceinline__ uint32_t& NB(int buf, int plane, int k) const { return pay[((buf * 3 + plane) * FSS_CHUNK_W + k) * pstride]; }
    __device__ __forceinline__ void fetch_row_async(int buf, cidx_t i) const {
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
#ifdef __CUDA_ARCH__
            __pipeline_memcpy_async(&NB(buf, 0, k), p.g.mf + i + 32 * k, 4);
            __pipeline_memcpy_async(&NB(buf, 1, k), p.g.amt + i + 32 * k, 4);
            __pipeline_memcpy_async(&NB(buf, 2, k), p.g.dif + i + 32 * k, 4);
#else
            NB(buf, 0, k) = p.g.mf[i + 32 * k];
            memcpy(&NB(buf, 1, k), &p.g.amt[i + 32 * k], 4);
            memcpy(&NB(buf, 2, k), &p.g.dif[i + 32 * k], 4);
#endif
        }
#ifdef __CUDA_ARCH__
        __pipeline_commit();
#endif
    }
    __device__ __forceinline__ void take_row(LRow& R, int buf) const {
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            R.mf[k] = NB(buf, 0, k);
#ifdef __CUDA_ARCH__
            R.amt[k] = __uint_as_float(NB(buf, 1, k));
            R.dif[k] = __uint_as_float(NB(buf, 2, k));
#else
            memcpy(&R.amt[k], &NB(buf, 1, k), 4);
            memcpy(&R.dif[k], &NB(buf, 2, k), 4);
#endif
        }
        R.fl = 0u;
    }
    static __device__ __forceinline__ void wait_payload() {
#ifdef __CUDA_ARCH__
        __pipeline_wait_prior(0);
#endif
    }
    __device__ __forceinline__ void load_side(LSide& S, cidx_t i) const {
        S.mf = p.g.mf[i];
        S.amt = p.g.amt[i];
        S.dif = p.g.dif[i];
    }

    // `neighbour created (copy of mat/colour/temperature, amount 0) if AIR; neighbour.diff += flow`
    // (world.cpp:1399-1403, :1441-1445, :1471-1475, :1503-1507), into a window cell; is = own cell, in = the neighbour
    // (K = the giver's column in row slot sC; sx = the slot of row X)
    template <int J, int K>
    __device__ __forceinline__ void give(LRow& X, float flow, uint32_t self_mf, cidx_t is, cidx_t in, int sx) const {
        if(MF_IS(X.mf[J], FSS_PHYS_AIR)) {
            X.mf[J] = self_mf & (MF_MAT_MASK | MF_DERIVED_MASK);
            X.amt[J] = 0.0f;
            X.dif[J] = 0.0f + flow;
            X.fl |= LF_ALL << J;
            if(SAND) {
                wait_payload(); // (row r-1's words may still be on their way in)
                P(sx, 0, J) = P(sC, 0, K);
                P(sx, 1, J) = P(sC, 1, K);
                X.fl |= LF_PAY << J;
            } else {
                p.g.col[in] = p.g.col[is];
                p.g.temp[in] = p.g.temp[is];
            }
        } else {
            X.dif[J] = X.dif[J] + flow;
            X.fl |= LF_DIF << J;
        }
    }
    // the same into an outside column: written through
    template <int K>
    __device__ __forceinline__ void give_side(LSide& S, float flow, uint32_t self_mf, cidx_t is, cidx_t in) const {
        if(MF_IS(S.mf, FSS_PHYS_AIR)) {
            S.mf = self_mf & (MF_MAT_MASK | MF_DERIVED_MASK);
            p.g.mf[in] = S.mf;
            p.g.col[in] = SAND ? P(sC, 0, K) : p.g.col[is];
            p.g.temp[in] = SAND ? (int32_t)P(sC, 1, K) : p.g.temp[is];
            p.g.amt[in] = 0.0f;
            p.g.dif[in] = 0.0f + flow;
        } else {
            p.g.dif[in] = S.dif + flow;
        }
    }

    // two vertically adjacent cells of column K change places (X over Y in the window; ix / iy their indices); the colour and
    // temperature words, which the window does not hold, change places in memory.  xmf / xdif: what the cell that was in X
    // carries into Y (SAND: `moved` may be set on the way, :1300; SOUP :1511: its running fluidAmountDiff)
    template <int K>
    __device__ __forceinline__ void exchange(LRow& X, LRow& Y, cidx_t ix, cidx_t iy, int sx, int sy, uint32_t xmf, float xdif) const {
        if(SAND) {
            const uint32_t c = P(sx, 0, K), t = P(sx, 1, K);
            P(sx, 0, K) = P(sy, 0, K);
            P(sx, 1, K) = P(sy, 1, K);
            P(sy, 0, K) = c;
            P(sy, 1, K) = t;
            X.fl |= LF_PAY << K;
            Y.fl |= LF_PAY << K;
        }
        const uint32_t cx_ = SAND ? 0u : p.g.col[ix], cy_ = SAND ? 0u : p.g.col[iy];
        const int32_t tx = SAND ? 0 : p.g.temp[ix], ty = SAND ? 0 : p.g.temp[iy];
        const float xamt = X.amt[K];
        X.mf[K] = Y.mf[K];
        X.amt[K] = Y.amt[K];
        X.dif[K] = Y.dif[K];
        Y.mf[K] = xmf;
        Y.amt[K] = xamt;
        Y.dif[K] = xdif;
        X.fl |= LF_ALL << K;
        Y.fl |= LF_ALL << K;
        if(!SAND) {
            p.g.col[ix] = cy_;
            p.g.col[iy] = cx_;
            p.g.temp[ix] = ty;
            p.g.temp[iy] = tx;
        }
    }

    __device__ __forceinline__ bool can_move_into(uint32_t dst_mf, uint32_t self_mf) const { // world.cpp:1276-1285, on the density ranks
        return MF_IS(dst_mf, FSS_PHYS_AIR) || (!MF_IS(dst_mf, FSS_PHYS_SOLID) && (dst_mf & MF_RANK_MASK) < (self_mf & MF_RANK_MASK));
    }

    // ---- SAND, scan 1, world.cpp:1218-1337 (no reactions: liq_window_applies; no free fall: particles are off) ------------
    // (one way out of every rule body below: each `return` in the middle of a body is a merge point at which the compiler
    // shuffles the window registers into place again -- ncu counted 35 such moves per cell visit)
    template <int K>
    __device__ __forceinline__ void sand1(LRow& C, LRow& B, cidx_t irow, int r, uint32_t mf) {
        bool reacted = false;
        if(GEN && HI_NREACT(sinfo[mf_mat(mf)])) { // world.cpp:1251-1274; the cell's temperature is in the payload slot of row r
            busy = true;                           // (its behaviour depends on the temperature plane, which tickTemperature rewrites)
            Cell c;
            if(sand_react_cell(p, cx + K, cy + r, mf, (int32_t)P(sC, 1, K), c)) {
                C.mf[K] = c.mf;
                C.amt[K] = c.amt;
                C.dif[K] = c.dif;
                P(sC, 0, K) = c.col;
                P(sC, 1, K) = (uint32_t)c.temp;
                C.fl |= (LF_ALL | LF_PAY) << K;
                C.fl |= LF_VIS << K; // (a product is marked visited, so even a GAS product gets no scan 2 / 3 visit)
                reacted = true;
            }
        }
        if(!reacted) sand1_move<K>(C, B, irow, r, mf);
    }
    template <int K>
    __device__ __forceinline__ void sand1_move(LRow& C, LRow& B, cidx_t irow, int r, uint32_t mf) {
        const uint32_t bmf = B.mf[K];
        const uint32_t blmf = K > 0 ? B.mf[K > 0 ? K - 1 : 0] : lb_mf;
        const uint32_t brmf = K < FSS_CHUNK_W - 1 ? B.mf[K < FSS_CHUNK_W - 1 ? K + 1 : 0] : rb_mf;
        const bool can_b = can_move_into(bmf, mf), can_l = can_move_into(blmf, mf), can_r = can_move_into(brmf, mf);
        if(can_b) { // :1276, :1287
            busy = true;
            const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)(cx + K), (uint32_t)(cy + r));
#define SAND1_DRAW(bit, site, mod) (fss_rand(key, site, 0) % (mod) == 0)
            if(!((can_l || can_r) && SAND1_DRAW(0, 1287, 20))) {
                const cidx_t i = irow + 32 * K, ib = i + p.g.pitch;
                uint32_t tile_mf = mf;
                if(SAND1_DRAW(1, 1300, 2)) tile_mf |= MF_MOVED;
                exchange<K>(C, B, i, ib, sC, sB, tile_mf, C.dif[K]); // :1296-1306
                B.fl |= LF_VIS << K;                                 // :1308 tickVisited[x + (y + 1) * width]; row 16 is outside the chunk (ignored)
                if(SAND1_DRAW(2, 1313, 2)) { // :1313-1335 transmit movement
                    if(MF_IS(blmf, FSS_PHYS_SAND) && SAND1_DRAW(3, 1316, 2) && !(blmf & MF_MOVED)) {
                        if(K > 0) {
                            B.mf[K > 0 ? K - 1 : 0] = blmf | MF_MOVED;
                            B.fl |= LF_MF << (K > 0 ? K - 1 : 0);
                        } else {
                            lb_mf = blmf | MF_MOVED;
                            p.g.mf[irow + p.g.pitch + lofs] = lb_mf;
                        }
                    }
                    if(MF_IS(brmf, FSS_PHYS_SAND) && SAND1_DRAW(4, 1327, 2) && !(brmf & MF_MOVED)) {
                        if(K < FSS_CHUNK_W - 1) {
                            B.mf[K < FSS_CHUNK_W - 1 ? K + 1 : 0] = brmf | MF_MOVED;
                            B.fl |= LF_MF << (K < FSS_CHUNK_W - 1 ? K + 1 : 0);
                        } else {
                            rb_mf = brmf | MF_MOVED;
                            p.g.mf[irow + p.g.pitch + rofs] = rb_mf;
                        }
                    }
                }
            }
#undef SAND1_DRAW
        }
    }

    // ---- scan 1 of one cell, world.cpp:1154-1665, on the window: A / C / B = rows r-1 / r / r+1, L / R = the outside cells of row r
    template <int K>
    __device__ __forceinline__ void cell(LRow& A, LRow& C, LRow& B, LSide& L, LSide& R, const LSide& Ln, const LSide& Rn, cidx_t irow, int r) {
        const uint32_t mf = C.mf[K];
        // GEN: a GAS cell that rose into this position: already visited (:1161)
        const bool seen = GEN && (C.fl & (LF_VIS << K)) != 0u;
        const bool spent = seen || p.iter >= mf_iters(mf); // :1163-1166
        if(spent) C.fl |= LF_VIS << K;
        if(FIRE && !spent && (mf & MF_FIRE)) { // an active fire cell: k_fire_prepass has decided its visit (fire_prepass_chunk)
            busy = true;
            const uint32_t v = p.fire_verdict[irow + 32 * K];
            if(v) {
                p.fire_verdict[irow + 32 * K] = 0;
                if(v == FIRE_DIES) { // `tiles[index] = Tiles::NOTHING`, visited
                    const Cell n = cell_nothing(p.nothing_mf);
                    C.mf[K] = n.mf;
                    C.amt[K] = n.amt;
                    C.dif[K] = n.dif;
                    P(sC, 0, K) = n.col;
                    P(sC, 1, K) = (uint32_t)n.temp;
                    C.fl |= (LF_ALL | LF_PAY) << K;
                }
                C.fl |= LF_VIS << K; // (FIRE_NEW: set alight in this sub-step, :1205)
            }
        }
        if(SAND && !spent && MF_IS(mf, FSS_PHYS_SAND)) sand1<K>(C, B, irow, r, mf);
        if(GEN == 2 && !spent && MF_IS(mf, FSS_PHYS_GAS)) gas1<K>(A, C, Ln, Rn, irow, r, mf);
        // ---- SOUP, world.cpp:1338-1537
        const float amt = C.amt[K];
        const bool soup = !spent && MF_IS(mf, FSS_PHYS_SOUP) && amt != 0.0f; // :1344
        if(soup && amt < FLUID_MinValue) {                                   // :1346-1350
            busy = true;
            C.amt[K] = 0.0f;
            C.fl |= LF_AMT << K;
        } else if(soup && !(mf & MF_MOVED)) { // :1376 (moved == settled for liquids)
            busy = true;                      // an unsettled liquid cell always changes: it flows, or its settle counter advances
            soup1<K>(A, C, B, L, R, irow, r, mf, amt);
        }
    }

    // ---- GAS, scan 1, world.cpp:1647-1663: rises into AIR above (an exchange with row r-1 of the window)
    template <int K>
    __device__ __forceinline__ void gas1(LRow& A, LRow& C, const LSide& Ln, const LSide& Rn, cidx_t irow, int r, uint32_t mf) {
        const uint32_t amf = A.mf[K];
        const uint32_t almf = K > 0 ? A.mf[K > 0 ? K - 1 : 0] : Ln.mf;
        const uint32_t armf = K < FSS_CHUNK_W - 1 ? A.mf[K < FSS_CHUNK_W - 1 ? K + 1 : 0] : Rn.mf;
        if(MF_IS(amf, FSS_PHYS_AIR)) {
            busy = true;
            const bool open = MF_IS(almf, FSS_PHYS_AIR) || MF_IS(armf, FSS_PHYS_AIR);
            if(!(open && fss_rand(fss_rng_cell_key(p.tkey, (uint32_t)(cx + K), (uint32_t)(cy + r)), 1654, 0) % 2 == 0)) {
                const cidx_t i = irow + 32 * K;
                wait_payload(); // (row r-1's words may still be on their way in)
                exchange<K>(C, A, i, i - p.g.pitch, sC, sA, mf, C.dif[K]);
                A.fl |= LF_VIS << K; // :1661 tickVisited[x + (y - 1) * width]; row -1 is outside the chunk (ignored)
            }
        }
    }

    template <int K>
    __device__ __forceinline__ void soup1(LRow& A, LRow& C, LRow& B, LSide& L, LSide& R, cidx_t irow, int r, const uint32_t mf, const float amt) {
        const cidx_t pitch = p.g.pitch;
        const cidx_t i = irow + 32 * K;
        const uint32_t mat = mf_mat(mf);
        const float start = amt;
        float remaining = amt;
        float tdif = C.dif[K];
        uint32_t& lmf = K > 0 ? C.mf[K > 0 ? K - 1 : 0] : L.mf;
        uint32_t& rmf = K < FSS_CHUNK_W - 1 ? C.mf[K < FSS_CHUNK_W - 1 ? K + 1 : 0] : R.mf;
        const float lamt = K > 0 ? C.amt[K > 0 ? K - 1 : 0] : L.amt;
        const float ramt = K < FSS_CHUNK_W - 1 ? C.amt[K < FSS_CHUNK_W - 1 ? K + 1 : 0] : R.amt;
        const uint32_t bmf = B.mf[K], tmf = A.mf[K];
        // the by-value copies of the four neighbours decide the un-settling at the end (:1529-1535)
        const uint32_t lmf0 = lmf, rmf0 = rmf;
        bool swapped = false; // ITER0: the cell changed places with a different liquid below / above and its visit is over

        const bool air_below = MF_IS(bmf, FSS_PHYS_AIR); // :1381
        if((air_below && p.iter <= 2) || mf_mat(bmf) == mat) { // :1384-1405
            const float dst = MF_IS(bmf, FSS_PHYS_SOUP) ? B.amt[K] : 0.0f;
            float flow = vertical_flow_fast(start, dst) - dst;
            flow = clamp_flow(flow, start);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give<K, K>(B, flow, mf, i, i + pitch, sB);
            }
        } else if(ITER0 && p.iter == 0 && MF_IS(bmf, FSS_PHYS_SOUP)) { // :1406-1412 (a different liquid below)
            const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)(cx + K), (uint32_t)(cy + r));
            if(fss_rand(key, 1407, 0) % 10 == 0) {
                exchange<K>(C, B, i, i + pitch, sC, sB, mf, C.dif[K]);
                swapped = true;
            }
        }
        bool done = swapped || remaining < FLUID_MinValue; // :1414-1418
        if(!done) {
            // :1420-1424
            const bool can_left = (MF_IS(lmf0, FSS_PHYS_AIR) || mf_mat(lmf0) == mat) && !air_below;
            const bool can_right = (MF_IS(rmf0, FSS_PHYS_AIR) || mf_mat(rmf0) == mat) && !air_below;
            if(can_left) { // :1426-1448
                const float dst = MF_IS(lmf0, FSS_PHYS_SOUP) ? lamt : 0.0f;
                const float d = remaining - dst;
                float flow = can_right ? fdiv_const(d, 3.0f, 1.0f / 3.0f) : d / 2.0f;
                flow = clamp_flow(flow, remaining);
                if(flow != 0) {
                    remaining -= flow;
                    tdif -= flow;
                    if(K > 0) give<(K > 0 ? K - 1 : 0), K>(C, flow, mf, i, i - 32, sC);
                    else give_side<K>(L, flow, mf, i, irow + lofs);
                }
            }
            done = remaining < FLUID_MinValue;
            if(!done && can_right) { // :1456-1478 (`canMoveLeft ? 2.0f : 2.0f`)
                const float dst = MF_IS(rmf0, FSS_PHYS_SOUP) ? ramt : 0.0f;
                float flow = (remaining - dst) / 2.0f;
                flow = clamp_flow(flow, remaining);
                if(flow != 0) {
                    remaining -= flow;
                    tdif -= flow;
                    if(K < FSS_CHUNK_W - 1) give<(K < FSS_CHUNK_W - 1 ? K + 1 : 0), K>(C, flow, mf, i, i + 32, sC);
                    else give_side<K>(R, flow, mf, i, irow + rofs);
                }
                done = remaining < FLUID_MinValue;
            }
            // :1486-1516
            if(!done) {
                if(MF_IS(tmf, FSS_PHYS_AIR) || mf_mat(tmf) == mat) {
                    const float dst = MF_IS(tmf, FSS_PHYS_SOUP) ? A.amt[K] : 0.0f;
                    float flow = remaining - vertical_flow_fast(remaining, dst);
                    flow = clamp_flow(flow, remaining);
                    if(flow != 0) {
                        remaining -= flow;
                        tdif -= flow;
                        give<K, K>(A, flow, mf, i, i - pitch, sA);
                    }
                    done = remaining < FLUID_MinValue;
                } else if(ITER0 && p.iter == 0 && MF_IS(tmf, FSS_PHYS_SOUP)) { // :1510-1516
                    const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)(cx + K), (uint32_t)(cy + r));
                    if(fss_rand(key, 1511, 0) % 10 == 0) {
                        if(SAND) wait_payload(); // (row r-1's words may still be on their way in)
                        exchange<K>(C, A, i, i - pitch, sC, sA, mf, tdif);
                        swapped = true;
                    }
                }
            }
        }
        if(!swapped) {
            if(done) { // :1414-1418, :1450-1454, :1480-1484, :1518-1522
                C.dif[K] = tdif - remaining;
            } else {
                if(start == remaining) { // :1524-1528
                    const uint32_t settle = (mf_settle(mf) + 1u) & 0xFFu;
                    uint32_t nmf = (mf & 0x00FFFFFFu) | (settle << MF_SETTLE_SHIFT);
                    if(settle >= 10u) nmf |= MF_MOVED;
                    C.mf[K] = nmf;
                    C.fl |= LF_MF << K;
                } else { // :1529-1535: the by-value copies decide (a neighbour seeded above still counts as AIR)
                    if(MF_IS(tmf, FSS_PHYS_SOUP) && (tmf & MF_MOVED)) {
                        A.mf[K] = tmf & ~MF_MOVED;
                        A.fl |= LF_MF << K;
                    }
                    if(MF_IS(bmf, FSS_PHYS_SOUP) && (bmf & MF_MOVED)) {
                        B.mf[K] = bmf & ~MF_MOVED;
                        B.fl |= LF_MF << K;
                    }
                    if(MF_IS(lmf0, FSS_PHYS_SOUP) && (lmf0 & MF_MOVED)) {
                        lmf = lmf0 & ~MF_MOVED;
                        if(K > 0) C.fl |= LF_MF << (K > 0 ? K - 1 : 0);
                        else p.g.mf[irow + lofs] = lmf;
                    }
                    if(MF_IS(rmf0, FSS_PHYS_SOUP) && (rmf0 & MF_MOVED)) {
                        rmf = rmf0 & ~MF_MOVED;
                        if(K < FSS_CHUNK_W - 1) C.fl |= LF_MF << (K < FSS_CHUNK_W - 1 ? K + 1 : 0);
                        else p.g.mf[irow + rofs] = rmf;
                    }
                }
                C.dif[K] = tdif; // :1537 `tiles[index] = tile`
            }
            C.fl |= LF_DIF << K;
        }
    }

    // world.cpp:1075-1088, the three cases computed and selected (each case is the reference's own expression)
    static __device__ __forceinline__ float vertical_flow_fast(float remaining, float dest) {
        const float sum = remaining + dest;
        // `mid` is only selected for 0.5 < sum < 1.1, where its dividend lies in [0.3, 0.36]: fdiv_const is exact there
        const float mid = fdiv_const(FLUID_MaxValue * FLUID_MaxValue + sum * FLUID_MaxCompression, FLUID_MaxValue + FLUID_MaxCompression,
                                     1.0f / (FLUID_MaxValue + FLUID_MaxCompression));
        const float high = (sum + FLUID_MaxCompression) / 2.0f;
        const float value = sum < 2 * FLUID_MaxValue + FLUID_MaxCompression ? mid : high;
        return sum <= FLUID_MaxValue ? FLUID_MaxValue : value;
    }

    // A row leaves the window.  A chunk row (inside = true, r = 0 .. 15) first gets scan 2's SOUP step (world.cpp:1808-1825:
    // `fluidAmount += fluidAmountDiff; fluidAmountDiff = 0`, below FLUID_MinValue -> Tiles::NOTHING) on its unvisited SOUP cells;
    // then whatever differs from memory is stored.
    __device__ __forceinline__ void retire(const LRow& X, cidx_t irow, int r, bool inside, int slot) {
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            const cidx_t i = irow + 32 * k;
            const uint32_t mf = X.mf[k];
            const float a = X.amt[k], d = X.dif[k];
            const bool visited = (X.fl & (LF_VIS << k)) != 0u;
            const bool soup = inside && !visited && MF_IS(mf, FSS_PHYS_SOUP);
            const float s = a + d;
            const bool gone = soup && s < FLUID_MinValue;
            if(!SAND && gone) {
                busy = true;
                st_cell(p.g, i, cell_nothing(p.nothing_mf));
                continue;
            }
            const bool apply = soup && !gone && __float_as_uint(d) != 0u; // else the stores would rewrite the same bits
            if(apply) busy = true;
            if(SAND && (X.fl & (LF_PAY << k))) {
                p.g.col[i] = P(slot, 0, k);
                p.g.temp[i] = (int32_t)P(slot, 1, k);
            }
            if(X.fl & (LF_MF << k)) p.g.mf[i] = mf;
            if(apply || (X.fl & (LF_AMT << k))) p.g.amt[i] = apply ? s : a;
            if(apply || (X.fl & (LF_DIF << k))) p.g.dif[i] = apply ? 0.0f : d;
            if(SAND && inside && (visited || (soup && !gone))) vis |= 1ull << (r * FSS_CHUNK_W + k);
            if(SAND && inside && !visited && MF_IS(mf, FSS_PHYS_SAND) && !(mf & MF_MOVED)) stopm |= 1ull << (r * FSS_CHUNK_W + k);
            if(GEN && inside && MF_IS(mf, FSS_PHYS_GAS)) gasm |= 1ull << (r * FSS_CHUNK_W + k);
        }
    }

    __device__ __forceinline__ void scans() {
        LRow A, C, B, N;
        LSide L, R, Ln, Rn;
        cidx_t ir = row_index(FSS_CHUNK_H - 1); // column 0 of the row being scanned
        const cidx_t pitch = p.g.pitch;
        load_row(B, ir + pitch);
        load_row(C, ir);
        int nbuf = 0;
        if(NSM) fetch_row_async(0, ir - pitch);
        else load_row(N, ir - pitch);
        load_side(L, ir + lofs);
        load_side(R, ir + rofs);
        if(SAND) {
            lb_mf = p.g.mf[ir + pitch + lofs];
            rb_mf = p.g.mf[ir + pitch + rofs];
            fetch_payload(sB, ir + pitch);
            fetch_payload(sC, ir);
        }
        Ln = L;
        Rn = R;
#pragma unroll 1
        for(int r = FSS_CHUNK_H - 1; r >= 0; r--) {
            if(NSM) {
                wait_payload();
                take_row(A, nbuf);
                nbuf ^= 1;
                if(r >= 1) fetch_row_async(nbuf, ir - 2 * pitch);
            } else {
                A = N;
            }
            if(SAND) { // rows r and r+1 have arrived (requested an iteration ago); row r-1 is needed from the next iteration on
                wait_payload();
                fetch_payload(sA, ir - pitch);
            }
            if(r >= 1 && !NSM) load_row(N, ir - 2 * pitch); // requested a whole row of arithmetic ahead of their use
            if(r >= 1 || GEN == 2) { // (a GAS cell of row 0 looks at the outside cells of row -1)
                load_side(Ln, ir - pitch + lofs);
                load_side(Rn, ir - pitch + rofs);
            }
            cell<0>(A, C, B, L, R, Ln, Rn, ir, r);
            cell<1>(A, C, B, L, R, Ln, Rn, ir, r);
            cell<2>(A, C, B, L, R, Ln, Rn, ir, r);
            cell<3>(A, C, B, L, R, Ln, Rn, ir, r);
            retire(B, ir + pitch, r + 1, r < FSS_CHUNK_H - 1, sB);
            B = C;
            C = A;
            {
                const int freed = sB;
                sB = sC;
                sC = sA;
                sA = freed;
            }
            if(SAND) {
                lb_mf = L.mf;
                rb_mf = R.mf;
            }
            L = Ln;
            R = Rn;
            ir -= pitch;
        }
        // ir = column 0 of row -1; B = row 0, C = row -1
        if(SAND) wait_payload();
        retire(B, ir + pitch, 0, true, sB);
        retire(C, ir, -1, false, sC);
    }

    // early-out bookkeeping after the scans (same contract as ChunkWalk::run)
    __device__ __forceinline__ void run(uint8_t* qown) {
        scans();
        if(SAND) { // scan 2's SAND rule (and the liquid cells its order matters for), through L1
            ChunkWalk<GEN ? K1_GENERAL : K1_NOGAS, false, false> w(p, sinfo, cx, cy, wmask);
            w.vis = vis;
            w.busy = busy;
#ifdef K1_SCAN2_LOCKSTEP
            w.scan2_rest();
#else
            w.scan2_sparse(stopm);
            if(GEN) w.scan3_sparse(gasm);
#endif
            busy = w.busy;
        }
        if(!qown) return;
        if(!busy) {
            *qown = (uint8_t)p.iter;
        } else {
#pragma unroll
            for(int a = -1; a <= 1; a++)
#pragma unroll
                for(int b = -1; b <= 1; b++) qown[a * p.q_w + b] = 0xFF;
        }
    }
};
