// fss_rules_liq.cuh -- the liquid-only sub-step of World::tick with the chunk rows held in REGISTERS (round 2).
//
// Which sub-steps: the ones k_substep<K1_LIQUID> serves (only SOUP cells can act: no SAND / GAS / fire material with
// `iterations > iter` can be present), further restricted to iter >= 1 (no SOUP-SOUP swaps, world.cpp:1406-1412 and
// :1510-1516, which draw random numbers and move whole cells), no FSS_PARTICLES_EMIT (the free-fall test of
// world.cpp:1352-1374 looks four rows down), no flow / dirty planes.  liq_window_applies() is the one definition of that
// condition; everything else keeps running ChunkWalk<K1_LIQUID> (fss_rules.cuh).
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

// the one place that decides which sub-steps take the register-window path (kernel launch and tests/hostsim)
__host__ __device__ __forceinline__ bool liq_window_applies(int mode, int iter, uint32_t cfg_flags, bool flow, bool dirty) {
#ifdef K1_NO_LIQ_WINDOW
    return false;
#else
    return mode == K1_LIQUID && iter >= 1 && !(cfg_flags & FSS_PARTICLES_EMIT) && !flow && !dirty;
#endif
}

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
#define LF_VIS 0x10000u

struct LSide { // the cell left / right of the row being scanned (outside the chunk)
    uint32_t mf;
    float amt, dif;
};

struct LiqWalk {
    const SubstepParams& p;
    int cx, cy;
    // word offsets relative to the chunk's column 0: the own columns are at +32*k (one 8-column group), the outside columns at
    // lofs / rofs (fss_device.cuh, permuted layout)
    int lofs, rofs;
    cidx_t c0; // index of (cx, cy) in the planes
    bool busy; // this visit stored to the grid (else it was the identity); see SubstepParams::q

    __device__ __forceinline__ LiqWalk(const SubstepParams& p_, const uint32_t*, int cx_, int cy_, unsigned)
        : p(p_), cx(cx_), cy(cy_), busy(false) {
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
    __device__ __forceinline__ void load_side(LSide& S, cidx_t i) const {
        S.mf = p.g.mf[i];
        S.amt = p.g.amt[i];
        S.dif = p.g.dif[i];
    }

    // `neighbour created (copy of mat/colour/temperature, amount 0) if AIR; neighbour.diff += flow`
    // (world.cpp:1399-1403, :1441-1445, :1471-1475, :1503-1507), into a window cell; is = own cell, in = the neighbour
    template <int J>
    __device__ __forceinline__ void give(LRow& X, float flow, uint32_t self_mf, cidx_t is, cidx_t in) const {
        if(MF_IS(X.mf[J], FSS_PHYS_AIR)) {
            X.mf[J] = self_mf & (MF_MAT_MASK | MF_DERIVED_MASK);
            X.amt[J] = 0.0f;
            X.dif[J] = 0.0f + flow;
            X.fl |= (LF_MF | LF_DIF | LF_AMT) << J;
            p.g.col[in] = p.g.col[is];
            p.g.temp[in] = p.g.temp[is];
        } else {
            X.dif[J] = X.dif[J] + flow;
            X.fl |= LF_DIF << J;
        }
    }
    // the same into an outside column: written through
    __device__ __forceinline__ void give_side(LSide& S, float flow, uint32_t self_mf, cidx_t is, cidx_t in) const {
        if(MF_IS(S.mf, FSS_PHYS_AIR)) {
            S.mf = self_mf & (MF_MAT_MASK | MF_DERIVED_MASK);
            p.g.mf[in] = S.mf;
            p.g.col[in] = p.g.col[is];
            p.g.temp[in] = p.g.temp[is];
            p.g.amt[in] = 0.0f;
            p.g.dif[in] = 0.0f + flow;
        } else {
            p.g.dif[in] = S.dif + flow;
        }
    }

    // ---- SOUP, scan 1, world.cpp:1338-1537, on the window: A / C / B = rows r-1 / r / r+1, L / R = the outside cells of row r
    template <int K>
    __device__ __forceinline__ void cell(LRow& A, LRow& C, LRow& B, LSide& L, LSide& R, cidx_t irow) {
        const uint32_t mf = C.mf[K];
        if(p.iter >= mf_iters(mf)) { // :1163-1166
            C.fl |= LF_VIS << K;
            return;
        }
        if(!MF_IS(mf, FSS_PHYS_SOUP)) return;
        const float amt = C.amt[K];
        if(amt == 0.0f) return;    // :1344
        if(amt < FLUID_MinValue) { // :1346-1350
            busy = true;
            C.amt[K] = 0.0f;
            C.fl |= LF_AMT << K;
            return;
        }
        if(mf & MF_MOVED) return; // :1376 (moved == settled for liquids)
        busy = true;              // an unsettled liquid cell always changes: it flows, or its settle counter advances
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

        const bool air_below = MF_IS(bmf, FSS_PHYS_AIR); // :1381
        if((air_below && p.iter <= 2) || mf_mat(bmf) == mat) { // :1384-1405
            const float dst = MF_IS(bmf, FSS_PHYS_SOUP) ? B.amt[K] : 0.0f;
            float flow = vertical_flow_fast(start, dst) - dst;
            flow = clamp_flow(flow, start);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give<K>(B, flow, mf, i, i + pitch);
            }
        }
        bool done = remaining < FLUID_MinValue; // :1414-1418
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
                    if(K > 0) give<(K > 0 ? K - 1 : 0)>(C, flow, mf, i, i - 32);
                    else give_side(L, flow, mf, i, irow + lofs);
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
                    if(K < FSS_CHUNK_W - 1) give<(K < FSS_CHUNK_W - 1 ? K + 1 : 0)>(C, flow, mf, i, i + 32);
                    else give_side(R, flow, mf, i, irow + rofs);
                }
                done = remaining < FLUID_MinValue;
            }
            // :1486-1516 (the swap of :1510-1516 needs iter == 0)
            if(!done && (MF_IS(tmf, FSS_PHYS_AIR) || mf_mat(tmf) == mat)) {
                const float dst = MF_IS(tmf, FSS_PHYS_SOUP) ? A.amt[K] : 0.0f;
                float flow = remaining - vertical_flow_fast(remaining, dst);
                flow = clamp_flow(flow, remaining);
                if(flow != 0) {
                    remaining -= flow;
                    tdif -= flow;
                    give<K>(A, flow, mf, i, i - pitch);
                }
                done = remaining < FLUID_MinValue;
            }
        }
        if(done) { // :1414-1418, :1450-1454, :1480-1484, :1518-1522
            C.dif[K] = tdif - remaining;
            C.fl |= LF_DIF << K;
            return;
        }
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
        C.fl |= LF_DIF << K;
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

    // A row leaves the window.  A chunk row (inside = true) first gets scan 2's SOUP step (world.cpp:1808-1825:
    // `fluidAmount += fluidAmountDiff; fluidAmountDiff = 0`, below FLUID_MinValue -> Tiles::NOTHING) on its unvisited SOUP cells;
    // then whatever differs from memory is stored.
    __device__ __forceinline__ void retire(const LRow& X, cidx_t irow, bool inside) {
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            const cidx_t i = irow + 32 * k;
            const uint32_t mf = X.mf[k];
            const float a = X.amt[k], d = X.dif[k];
            const bool soup = inside && !(X.fl & (LF_VIS << k)) && MF_IS(mf, FSS_PHYS_SOUP);
            const float s = a + d;
            if(soup && s < FLUID_MinValue) {
                busy = true;
                st_cell(p.g, i, cell_nothing(p.nothing_mf));
                continue;
            }
            const bool apply = soup && __float_as_uint(d) != 0u; // else the stores would rewrite the same bits
            if(apply) busy = true;
            if(X.fl & (LF_MF << k)) p.g.mf[i] = mf;
            if(apply || (X.fl & (LF_AMT << k))) p.g.amt[i] = apply ? s : a;
            if(apply || (X.fl & (LF_DIF << k))) p.g.dif[i] = apply ? 0.0f : d;
        }
    }

    __device__ __forceinline__ void scans() {
        LRow A, C, B, N;
        LSide L, R, Ln, Rn;
        cidx_t ir = row_index(FSS_CHUNK_H - 1); // column 0 of the row being scanned
        const cidx_t pitch = p.g.pitch;
        load_row(B, ir + pitch);
        load_row(C, ir);
        load_row(N, ir - pitch);
        load_side(L, ir + lofs);
        load_side(R, ir + rofs);
        Ln = L;
        Rn = R;
#pragma unroll 1
        for(int r = FSS_CHUNK_H - 1; r >= 0; r--) {
            A = N;
            if(r >= 1) { // requested a whole row of arithmetic ahead of their use
                load_row(N, ir - 2 * pitch);
                load_side(Ln, ir - pitch + lofs);
                load_side(Rn, ir - pitch + rofs);
            }
            cell<0>(A, C, B, L, R, ir);
            cell<1>(A, C, B, L, R, ir);
            cell<2>(A, C, B, L, R, ir);
            cell<3>(A, C, B, L, R, ir);
            retire(B, ir + pitch, r < FSS_CHUNK_H - 1);
            B = C;
            C = A;
            L = Ln;
            R = Rn;
            ir -= pitch;
        }
        // ir = column 0 of row -1; B = row 0, C = row -1
        retire(B, ir + pitch, true);
        retire(C, ir, false);
    }

    // early-out bookkeeping after the scans (same contract as ChunkWalk::run)
    __device__ __forceinline__ void run(uint8_t* qown) {
        scans();
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
