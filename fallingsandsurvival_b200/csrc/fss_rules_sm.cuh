// fss_rules_sm.cuh -- K1b: the movement rules with the chunk window staged in shared memory.
//
// Same rules, same visiting order inside a micro-chunk, same results as fss_rules.cuh (K1a), different execution
// shape.  K1a keeps the 32 lanes of a warp in lock step over the 64 positions of their chunks so that global accesses
// coalesce; on a world where neighbouring chunks hold different materials every position then costs the SAND body plus
// the SOUP body (ncu r01b: 11.8 of 32 lanes active, 64 k instructions per warp and sub-step).  Here each lane first
// copies the part of its read window that the rules address at random -- 18 rows x 6 columns of the mf / amt / dif planes
// (rows -1..16, columns -1..4 around the chunk) -- into a PRIVATE slice of shared memory laid out [cell][lane], so that
// any lane may touch any of its cells in any order without a bank conflict (bank == lane).  The lanes then run
// DECOUPLED: each one steps through its own unvisited, active cells only, and the warp iterates max-over-lanes(active
// cells) times instead of 64.  The colour / temperature planes stay in HBM (they are only moved, never tested, except
// by the SAND reactions) and are accessed per lane when a cell moves.  Rows 17..25 below the chunk are only ever tested
// for AIR (pillar probe world.cpp:1700-1702, free-fall test :1288) and cannot change during the visit (the deepest
// write of a visit is a fire igniting a SOLID cell at row 17: not AIR before or after), so they are nine bits per
// column in two registers.  At the end the slice is written back.
//
// One warp per CTA (41.5 KB of shared memory + the material table), five CTAs per SM.
#pragma once
#include "fss_rules.cuh"

#define SW_ROWS 18 /* dy = -1 .. 16 */
#define SW_COLS 6  /* dx = -1 .. 4  */
#define SW_CELLS (SW_ROWS * SW_COLS)
#define SW_LOW_ROWS 9 /* dy = 17 .. 25: AIR bits only */

// What the rare, non-inlined rules need to reach this lane's cells: passed by value.
struct SliceView {
    uint32_t* smf;
    float* samt;
    float* sdif;
    int cx, cy;
    static __device__ __forceinline__ int H(int dx, int dy) { return (dy + 1) * SW_COLS + dx + 1; }
    static __device__ __forceinline__ bool in_slice(int dx, int dy) { return dx >= -1 && dx <= FSS_CHUNK_W && dy >= -1 && dy <= FSS_CHUNK_H; }
    // a whole cell at (dx, dy), inside or outside the slice: mf / amt / dif of slice cells live in shared memory
    __device__ __forceinline__ uint32_t mf_any(const Grid& g, int dx, int dy) const {
        return in_slice(dx, dy) ? smf[H(dx, dy) * 32] : g.mf[g.idx(cx + dx, cy + dy)];
    }
    __device__ __forceinline__ Cell ld_any(const Grid& g, int dx, int dy) const {
        const cidx_t j = g.idx(cx + dx, cy + dy);
        Cell c = ld_cell(g, j);
        if(in_slice(dx, dy)) {
            const int h = H(dx, dy) * 32;
            c.mf = smf[h];
            c.amt = samt[h];
            c.dif = sdif[h];
        }
        return c;
    }
    __device__ __forceinline__ void st_any(const Grid& g, int dx, int dy, const Cell& c) const {
        const cidx_t j = g.idx(cx + dx, cy + dy);
        if(in_slice(dx, dy)) {
            const int h = H(dx, dy) * 32;
            smf[h] = c.mf;
            samt[h] = c.amt;
            sdif[h] = c.dif;
            g.col[j] = c.col;
            g.temp[j] = c.temp;
        } else {
            st_cell(g, j, c);
        }
    }
    static __device__ __forceinline__ unsigned long long mark_bit(int dx, int dy) {
        if((unsigned)dx < (unsigned)FSS_CHUNK_W && (unsigned)dy < (unsigned)FSS_CHUNK_H) return 1ull << (dy * FSS_CHUNK_W + (FSS_CHUNK_W - 1 - dx));
        return 0ull;
    }
};

// FIRE, world.cpp:1171-1216; returns the tickVisited marks to add
__device__ __noinline__ unsigned long long fss_rare_fire(const SubstepParams& p, SliceView v, int dx, int dy) {
    const int x = v.cx + dx, y = v.cy + dy;
    const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
    unsigned long long marks = 0ull;
    Cell tile = v.ld_any(p.g, dx, dy);
    if(fss_rand(key, 1172, 0) % 10 == 0) tile.col = (((255u << 8) + 100 + fss_rand(key, 1174, 0) % 50) << 8) + 50;
    if(fss_rand(key, 1179, 0) % 10 == 0) {
        if(p.flags & FSS_PARTICLES_EMIT) emit_particle(p, x, y, 1180, key, tile);
    }
    if(fss_rand(key, 1191, 0) % 150 == 0) {
        v.st_any(p.g, dx, dy, cell_nothing(p.nothing_mf));
        return SliceView::mark_bit(dx, dy);
    }
    bool found = false;
    uint32_t k1202 = 0, kcreate = 0;
    for(int xx = -2; xx <= 2; xx++) {
        for(int yy = -2; yy <= 2; yy++) {
            if(MF_IS(v.mf_any(p.g, dx + xx, dy + yy), FSS_PHYS_SOLID)) {
                found = true;
                if(fss_rand(key, 1202, k1202++) % 500 == 0) {
                    v.st_any(p.g, dx + xx, dy + yy, create_cell(p, p.fire_id, x + xx, y + yy, key, kcreate++)); // Tiles::createFire()
                    marks |= SliceView::mark_bit(dx + xx, dy + yy);
                }
            }
        }
    }
    if(!found && fss_rand(key, 1210, 0) % 120 == 0) {
        v.st_any(p.g, dx, dy, cell_nothing(p.nothing_mf));
        marks |= SliceView::mark_bit(dx, dy);
    }
    return marks;
}

// SAND reactions, world.cpp:1251-1274
__device__ __noinline__ bool fss_rare_sand_react(const SubstepParams& p, SliceView v, int dx, int dy, uint32_t mf) {
    const int x = v.cx + dx, y = v.cy + dy;
    const DMat& m = p.mats[mf_mat(mf)];
    const int32_t temp = p.g.temp[p.g.idx(x, y)];
    const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
    bool react = false;
    uint32_t site_seen[FSS_MAX_REACTIONS];
    for(int r = 0; r < m.n_react && r < FSS_MAX_REACTIONS; r++) {
        site_seen[r] = 0xFFFFFFFFu;
        if((m.r_type[r] == FSS_REACT_TEMPERATURE_BELOW && temp < m.r_thr[r]) ||
           (m.r_type[r] == FSS_REACT_TEMPERATURE_ABOVE && temp > m.r_thr[r])) {
            const uint32_t site = p.mats[m.r_target[r]].create_site;
            uint32_t k = 0;
            for(int q = 0; q < r; q++) k += (site_seen[q] == site) ? 1u : 0u;
            site_seen[r] = site;
            Cell c = create_cell(p, m.r_target[r], x, y, key, k);
            c.temp = temp;
            v.st_any(p.g, dx, dy, c);
            react = true;
        }
    }
    return react;
}

// falling liquid leaves the grid as particles, world.cpp:1352-1374
__device__ __noinline__ void fss_rare_soup_free_fall(const SubstepParams& p, SliceView v, int dx, int dy, uint32_t mf, float amt) {
    const int x = v.cx + dx, y = v.cy + dy;
    const cidx_t g = p.g.idx(x, y);
    const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
    Cell nt;
    nt.mf = mf & (MF_MAT_MASK | MF_DERIVED_MASK);
    nt.col = p.g.col[g];
    nt.temp = p.g.temp[g];
    nt.dif = 0.0f;
    int n = (int)(amt / 4);
    if(n < 1) n = 1;
    nt.amt = amt / n;
    v.st_any(p.g, dx, dy, cell_nothing(p.nothing_mf));
    for(int k = 0; k < n; k++) emit_particle(p, x, y, 1366, key, nt);
}

struct ChunkWalkS {
    const SubstepParams& p;
    const float* sden;
    const uint32_t* sinfo;
    uint32_t* smf; // this lane's slice: cell h at [h * 32]
    float* samt;
    float* sdif;
    int cx, cy;
    uint32_t pb, pl, pr; // permuted storage columns: own column dx at pb + 32*dx, dx = -1 at pl, dx = 4 at pr
    cidx_t row0;         // storage offset of row cy
    unsigned long long marked; // explicit tickVisited marks, bit = scan position (dy*4 + 3 - dx)
    unsigned long long pv;     // cells found `iter >= iterations` at their scan-1 visit (world.cpp:1163-1166)
    unsigned long long airlow; // bit c*9 + q: column dx = c-1, row dy = 17+q is AIR
    bool busy, maybe_gas;

    __device__ __forceinline__ ChunkWalkS(const SubstepParams& p_, const float* sd, const uint32_t* si, uint32_t* m, float* a, float* d,
                                          int cx_, int cy_)
        : p(p_), sden(sd), sinfo(si), smf(m), samt(a), sdif(d), cx(cx_), cy(cy_), marked(0ull), pv(0ull), airlow(0ull), busy(false),
          maybe_gas(false) {
        pb = fss_perm((uint32_t)cx);
        pl = fss_perm((uint32_t)(cx - 1));
        pr = fss_perm((uint32_t)(cx + FSS_CHUNK_W));
        row0 = (cidx_t)(cy - p.g.sy0) * p.g.pitch;
    }

    // ---- addressing ------------------------------------------------------------------------------------------
    static __device__ __forceinline__ int H(int dx, int dy) { return (dy + 1) * SW_COLS + dx + 1; }
    __device__ __forceinline__ uint32_t colp(int dx) const { return dx < 0 ? pl : (dx >= FSS_CHUNK_W ? pr : pb + 32u * (uint32_t)dx); }
    __device__ __forceinline__ cidx_t G(int dx, int dy) const { return row0 + (cidx_t)((long long)dy * (long long)p.g.pitch) + colp(dx); }
    __device__ __forceinline__ uint32_t& MF(int h) const { return smf[h * 32]; }
    __device__ __forceinline__ float& AMT(int h) const { return samt[h * 32]; }
    __device__ __forceinline__ float& DIF(int h) const { return sdif[h * 32]; }
    static __device__ __forceinline__ int pos_of(int dx, int dy) { return dy * FSS_CHUNK_W + (FSS_CHUNK_W - 1 - dx); }
    __device__ __forceinline__ void mark(int dx, int dy) {
        if((unsigned)dx < (unsigned)FSS_CHUNK_W && (unsigned)dy < (unsigned)FSS_CHUNK_H) marked |= 1ull << pos_of(dx, dy);
    }
    // type test of (dx, dy) for dx in -1..4, dy in -1..25
    __device__ __forceinline__ bool is_air(int dx, int dy) const {
        if(dy <= FSS_CHUNK_H) return MF_IS(MF(H(dx, dy)), FSS_PHYS_AIR);
        return (airlow >> ((dx + 1) * SW_LOW_ROWS + (dy - (FSS_CHUNK_H + 1)))) & 1ull;
    }
    __device__ __forceinline__ bool can_move_into(uint32_t dst_mf, float self_density) const {
        if(MF_IS(dst_mf, FSS_PHYS_AIR)) return true;
        if(MF_IS(dst_mf, FSS_PHYS_SOLID)) return false;
        return sden[mf_mat(dst_mf)] < self_density;
    }

    // a whole cell: mf / amt / dif from the slice, colour / temperature from HBM
    __device__ __forceinline__ Cell ld(int h, cidx_t g) const {
        Cell c;
        c.mf = MF(h);
        c.col = p.g.col[g];
        c.temp = p.g.temp[g];
        c.amt = AMT(h);
        c.dif = DIF(h);
        return c;
    }
    __device__ __forceinline__ void st(int h, cidx_t g, const Cell& c) const {
        MF(h) = c.mf;
        p.g.col[g] = c.col;
        p.g.temp[g] = c.temp;
        AMT(h) = c.amt;
        DIF(h) = c.dif;
    }
    __device__ __forceinline__ void swap(int ha, cidx_t ga, int hb, cidx_t gb) const {
        Cell a = ld(ha, ga), b = ld(hb, gb);
        st(ha, ga, b);
        st(hb, gb, a);
    }

    // ---- staging ---------------------------------------------------------------------------------------------
    __device__ __forceinline__ void stage_in() {
        const Grid& g = p.g;
#pragma unroll 2
        for(int r = 0; r < SW_ROWS; r++) {
            const cidx_t rb = row0 + (cidx_t)((long long)(r - 1) * (long long)g.pitch);
            uint32_t m[SW_COLS];
            float a[SW_COLS], d[SW_COLS];
#pragma unroll
            for(int c = 0; c < SW_COLS; c++) {
                const cidx_t j = rb + colp(c - 1);
                m[c] = g.mf[j];
                a[c] = g.amt[j];
                d[c] = g.dif[j];
            }
#pragma unroll
            for(int c = 0; c < SW_COLS; c++) {
                const int h = r * SW_COLS + c;
                MF(h) = m[c];
                AMT(h) = a[c];
                DIF(h) = d[c];
            }
        }
        unsigned long long low = 0ull;
#pragma unroll
        for(int q = 0; q < SW_LOW_ROWS; q++) {
            const cidx_t rb = row0 + (cidx_t)(FSS_CHUNK_H + 1 + q) * g.pitch;
#pragma unroll
            for(int c = 0; c < SW_COLS; c++)
                if(MF_IS(g.mf[rb + colp(c - 1)], FSS_PHYS_AIR)) low |= 1ull << (c * SW_LOW_ROWS + q);
        }
        airlow = low;
    }
    __device__ __forceinline__ void stage_out() const {
        const Grid& g = p.g;
#pragma unroll 2
        for(int r = 0; r < SW_ROWS; r++) {
            const cidx_t rb = row0 + (cidx_t)((long long)(r - 1) * (long long)g.pitch);
#pragma unroll
            for(int c = 0; c < SW_COLS; c++) {
                const cidx_t j = rb + colp(c - 1);
                const int h = r * SW_COLS + c;
                g.mf[j] = MF(h);
                g.amt[j] = AMT(h);
                g.dif[j] = DIF(h);
            }
        }
    }

    // ---- rare rules live outside the struct (fss_rare_*): they are not inlined, and a member function would force
    //      the whole walk state into local memory through `this`
    __device__ __forceinline__ void fire(int dx, int dy, int h, uint32_t mf);
    __device__ __forceinline__ bool sand_react(int dx, int dy, int h, uint32_t mf);
    __device__ __forceinline__ void soup_free_fall(int dx, int dy, int h, uint32_t mf, float amt);

    // `neighbour created (copy of mat/colour/temperature, amount 0) if AIR; neighbour.diff += flow`
    __device__ __forceinline__ void give(int hn, int ndx, int ndy, uint32_t nmf, float ndif, float flow, uint32_t self_mf, cidx_t gself) const {
        if(MF_IS(nmf, FSS_PHYS_AIR)) {
            const cidx_t gn = G(ndx, ndy);
            MF(hn) = self_mf & (MF_MAT_MASK | MF_DERIVED_MASK);
            p.g.col[gn] = p.g.col[gself];
            p.g.temp[gn] = p.g.temp[gself];
            AMT(hn) = 0.0f;
            DIF(hn) = 0.0f + flow;
        } else {
            DIF(hn) = ndif + flow;
        }
    }

    // ---- SOUP, scan 1, world.cpp:1338-1537 -------------------------------------------------------------------
    __device__ __forceinline__ void soup1(int dx, int dy, int h, uint32_t mf) {
        const float amt = AMT(h);
        if(amt == 0.0f) return;
        if(amt < FLUID_MinValue) {
            busy = true;
            AMT(h) = 0.0f;
            return;
        }
        if((p.flags & FSS_PARTICLES_EMIT) && (double)amt > 0.005 && is_air(dx, dy + 1) && is_air(dx, dy + 2) && is_air(dx, dy + 3) &&
           is_air(dx, dy + 4)) {
            busy = true;
            soup_free_fall(dx, dy, h, mf, amt);
            return;
        }
        if(mf & MF_MOVED) return;
        busy = true;
        const float start = amt;
        float remaining = amt;
        const uint32_t mat = mf_mat(mf);
        const int iter = p.iter;
        const int hb = h + SW_COLS, ht = h - SW_COLS, hl = h - 1, hr = h + 1;
        float tdif = DIF(h);
        const uint32_t bmf = MF(hb), lmf = MF(hl), rmf = MF(hr), tmf = MF(ht);
        const cidx_t g = G(dx, dy);
        const int x = cx + dx, y = cy + dy;

        const bool air_below = MF_IS(bmf, FSS_PHYS_AIR);
        if((air_below && iter <= 2) || mf_mat(bmf) == mat) {
            const float dst = MF_IS(bmf, FSS_PHYS_SOUP) ? AMT(hb) : 0.0f;
            float flow = vertical_flow<false>(start, dst) - dst;
            flow = clamp_flow(flow, start);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(hb, dx, dy + 1, bmf, DIF(hb), flow, mf, g);
            }
        } else if(iter == 0 && MF_IS(bmf, FSS_PHYS_SOUP)) {
            const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
            if(fss_rand(key, 1407, 0) % 10 == 0) {
                swap(h, g, hb, G(dx, dy + 1));
                return;
            }
        }
        if(remaining < FLUID_MinValue) {
            DIF(h) = tdif - remaining;
            return;
        }
        const bool can_left = (MF_IS(lmf, FSS_PHYS_AIR) || mf_mat(lmf) == mat) && !air_below;
        const bool can_right = (MF_IS(rmf, FSS_PHYS_AIR) || mf_mat(rmf) == mat) && !air_below;
        if(can_left) {
            const float dst = MF_IS(lmf, FSS_PHYS_SOUP) ? AMT(hl) : 0.0f;
            float flow = (remaining - dst) / (can_right ? 3.0f : 2.0f);
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(hl, dx - 1, dy, lmf, DIF(hl), flow, mf, g);
            }
        }
        if(remaining < FLUID_MinValue) {
            DIF(h) = tdif - remaining;
            return;
        }
        if(can_right) {
            const float dst = MF_IS(rmf, FSS_PHYS_SOUP) ? AMT(hr) : 0.0f;
            float flow = (remaining - dst) / 2.0f;
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(hr, dx + 1, dy, rmf, DIF(hr), flow, mf, g);
            }
        }
        if(remaining < FLUID_MinValue) {
            DIF(h) = tdif - remaining;
            return;
        }
        if(MF_IS(tmf, FSS_PHYS_AIR) || mf_mat(tmf) == mat) {
            const float dst = MF_IS(tmf, FSS_PHYS_SOUP) ? AMT(ht) : 0.0f;
            float flow = remaining - vertical_flow<false>(remaining, dst);
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(ht, dx, dy - 1, tmf, DIF(ht), flow, mf, g);
            }
        } else if(iter == 0 && MF_IS(tmf, FSS_PHYS_SOUP)) {
            const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
            if(fss_rand(key, 1511, 0) % 10 == 0) {
                const cidx_t gt = G(dx, dy - 1);
                Cell top = ld(ht, gt);
                Cell me = ld(h, g);
                me.dif = tdif;
                st(h, g, top);
                st(ht, gt, me);
                return;
            }
        }
        if(remaining < FLUID_MinValue) {
            DIF(h) = tdif - remaining;
            return;
        }
        uint32_t nmf = mf;
        if(start == remaining) {
            const uint32_t settle = (mf_settle(mf) + 1u) & 0xFFu;
            nmf = (mf & 0x00FFFFFFu) | (settle << MF_SETTLE_SHIFT);
            if(settle >= 10u) nmf |= MF_MOVED;
        } else {
            if(MF_IS(tmf, FSS_PHYS_SOUP) && (tmf & MF_MOVED)) MF(ht) = tmf & ~MF_MOVED;
            if(MF_IS(bmf, FSS_PHYS_SOUP) && (bmf & MF_MOVED)) MF(hb) = bmf & ~MF_MOVED;
            if(MF_IS(lmf, FSS_PHYS_SOUP) && (lmf & MF_MOVED)) MF(hl) = lmf & ~MF_MOVED;
            if(MF_IS(rmf, FSS_PHYS_SOUP) && (rmf & MF_MOVED)) MF(hr) = rmf & ~MF_MOVED;
        }
        if(nmf != mf) MF(h) = nmf;
        DIF(h) = tdif;
    }

    // ---- SAND, scan 1, world.cpp:1218-1337 -------------------------------------------------------------------
    __device__ __forceinline__ void sand1(int dx, int dy, int h, uint32_t mf) {
        const uint32_t mat = mf_mat(mf);
        if(HI_NREACT(sinfo[mat])) {
            busy = true;
            if(sand_react(dx, dy, h, mf)) {
                mark(dx, dy);
                return;
            }
        }
        const int hb = h + SW_COLS;
        const uint32_t bmf = MF(hb), blmf = MF(hb - 1), brmf = MF(hb + 1);
        const float dself = sden[mat];
        if(!can_move_into(bmf, dself)) return;
        busy = true;
        const bool can_l = can_move_into(blmf, dself), can_r = can_move_into(brmf, dself);
        const int x = cx + dx, y = cy + dy;
        const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
        if((can_l || can_r) && fss_rand(key, 1287, 0) % 20 == 0) return;
        const cidx_t g = G(dx, dy), gb = g + p.g.pitch;
        Cell tile = ld(h, g);
        Cell below = ld(hb, gb);
        if(MF_IS(bmf, FSS_PHYS_GAS)) maybe_gas = true;
        st(h, g, below);
        if((p.flags & FSS_PARTICLES_EMIT) && MF_IS(bmf, FSS_PHYS_AIR) && is_air(dx, dy + 2) && is_air(dx, dy + 3) && is_air(dx, dy + 4)) {
            emit_particle(p, x, y, 1291, key, tile);
        } else {
            if(fss_rand(key, 1300, 0) % 2 == 0) tile.mf |= MF_MOVED;
            st(hb, gb, tile);
            mark(dx, dy + 1);
        }
        if(fss_rand(key, 1313, 0) % 2 == 0) {
            if(MF_IS(blmf, FSS_PHYS_SAND) && fss_rand(key, 1316, 0) % 2 == 0 && !(blmf & MF_MOVED)) MF(hb - 1) = blmf | MF_MOVED;
            if(MF_IS(brmf, FSS_PHYS_SAND) && fss_rand(key, 1327, 0) % 2 == 0 && !(brmf & MF_MOVED)) MF(hb + 1) = brmf | MF_MOVED;
        }
    }

    // ---- one active cell of scan 1, world.cpp:1167-1663 --------------------------------------------------------
    __device__ __forceinline__ void scan1_body(int dx, int dy, uint32_t mf) {
        const int h = H(dx, dy);
        if(MF_IS(mf, FSS_PHYS_GAS)) maybe_gas = true;
        if(mf & MF_FIRE) {
            busy = true;
            fire(dx, dy, h, mf);
            return;
        }
        if(MF_IS(mf, FSS_PHYS_SAND)) {
            sand1(dx, dy, h, mf);
        } else if(MF_IS(mf, FSS_PHYS_SOUP)) {
            soup1(dx, dy, h, mf);
        } else if(MF_IS(mf, FSS_PHYS_GAS)) { // :1647-1663
            const int ha = h - SW_COLS;
            const uint32_t amf = MF(ha), almf = MF(ha - 1), armf = MF(ha + 1);
            if(MF_IS(amf, FSS_PHYS_AIR)) {
                busy = true;
                const bool open = MF_IS(almf, FSS_PHYS_AIR) || MF_IS(armf, FSS_PHYS_AIR);
                if(!(open && fss_rand(fss_rng_cell_key(p.tkey, (uint32_t)(cx + dx), (uint32_t)(cy + dy)), 1654, 0) % 2 == 0)) {
                    const cidx_t g = G(dx, dy);
                    swap(h, g, ha, g - p.g.pitch);
                    mark(dx, dy - 1);
                }
            }
        }
    }

    // ---- SAND, scan 2, world.cpp:1682-1807 -------------------------------------------------------------------
    __device__ __forceinline__ void sand2(int dx, int dy, int h, uint32_t mf) {
        const uint32_t mat = mf_mat(mf);
        const int hbl = h + SW_COLS - 1, hbr = h + SW_COLS + 1;
        const uint32_t blmf = MF(hbl), brmf = MF(hbr);
        const float dself = sden[mat];
        const bool can_l = can_move_into(blmf, dself), can_r = can_move_into(brmf, dself);
        if(!(can_l || can_r)) {
            if(mf & MF_MOVED) {
                busy = true;
                MF(h) = mf & ~MF_MOVED;
            }
            return;
        }
        const uint32_t info = sinfo[mat];
        const uint32_t slip = HI_SLIP(info);
        const int x = cx + dx, y = cy + dy;
        const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
        bool stopped = !(mf & MF_MOVED);
        if(stopped) { // :1697-1725
            int drop = 0;
#pragma unroll
            for(int pil = 0; pil < 10; pil++) drop += (is_air(dx - 1, dy + 1 + pil) || is_air(dx + 1, dy + 1 + pil)) ? 1 : 0;
            const int unstable = drop + 1 - (int)HI_STAB(info);
            if(unstable > 0) {
                const int chance = 1000 / unstable;
                if(chance < 1000) busy = true;
                if(chance < 1000 && fss_rand(key, 1715, 0) % (uint32_t)chance == 0) stopped = false;
            }
            if(stopped) return;
        }
        busy = true;
        const bool should_move = fss_rand(key, 1736, 0) % (2u * slip) != 0;
        if(!should_move) {
            if(mf & MF_MOVED) MF(h) = mf & ~MF_MOVED;
            return;
        }
        if(fss_rand(key, 1741, 0) % 2 == 0) {
            const uint32_t bmf = MF(h + SW_COLS);
            if(MF_IS(bmf, FSS_PHYS_SAND) && fss_rand(key, 1744, 0) % 2 == 0 && !(bmf & MF_MOVED)) MF(h + SW_COLS) = bmf | MF_MOVED;
        }
        const bool go_left = can_l && (!can_r || fss_rand(key, 1755, 0) % 2 == 0);
        const int sx = go_left ? -1 : 1;
        const int hdiag = go_left ? hbl : hbr, hside = h + sx;
        const cidx_t g = G(dx, dy), gside = G(dx + sx, dy), gdiag = gside + p.g.pitch;
        Cell tile = ld(h, g);
        tile.mf = mf;
        const Cell diag = ld(hdiag, gdiag);
        const uint32_t smf_ = MF(hside);
        if(MF_IS(diag.mf, FSS_PHYS_GAS)) maybe_gas = true;
        if(MF_IS(smf_, FSS_PHYS_AIR)) {
            st(hside, gside, diag);
            if(go_left) mark(dx - 1, dy);
            st(h, g, cell_nothing(p.nothing_mf));
        } else {
            st(h, g, diag);
            mark(dx, dy);
        }
        if(fss_rand(key, go_left ? 1768u : 1791u, 0) % (20u * slip) == 0) tile.mf &= ~MF_MOVED;
        st(hdiag, gdiag, tile);
        mark(dx + sx, dy + 1);
    }

    // ---- one unvisited cell of scan 2, world.cpp:1676-1901 -------------------------------------------------------
    __device__ __forceinline__ void scan2_body(int dx, int dy, uint32_t mf) {
        const int h = H(dx, dy);
        if(MF_IS(mf, FSS_PHYS_SAND)) {
            sand2(dx, dy, h, mf);
        } else if(MF_IS(mf, FSS_PHYS_SOUP)) { // :1808-1825
            const float dif = DIF(h);
            const float amt = AMT(h) + dif;
            if(__float_as_uint(dif) != 0u || amt < FLUID_MinValue) busy = true;
            if(amt < FLUID_MinValue) {
                st(h, G(dx, dy), cell_nothing(p.nothing_mf));
            } else {
                AMT(h) = amt;
                DIF(h) = 0.0f;
            }
            mark(dx, dy);
        } else { // GAS :1879-1899
            const int hal = h - SW_COLS - 1, har = h - SW_COLS + 1;
            const bool al = MF_IS(MF(hal), FSS_PHYS_AIR), ar = MF_IS(MF(har), FSS_PHYS_AIR);
            if(al || ar) busy = true;
            const cidx_t g = G(dx, dy);
            if(al && !(ar && fss_rand(fss_rng_cell_key(p.tkey, (uint32_t)(cx + dx), (uint32_t)(cy + dy)), 1884, 0) % 2 == 0)) {
                swap(h, g, hal, G(dx - 1, dy - 1));
                mark(dx - 1, dy - 1);
            } else if(ar) {
                swap(h, g, har, G(dx + 1, dy - 1));
                mark(dx + 1, dy - 1);
            }
        }
    }

    // ---- one unvisited GAS cell of scan 3, world.cpp:1905-1974 -----------------------------------------------------
    __device__ __forceinline__ void scan3_body(int dx, int dy, uint32_t mf) {
        const int h = H(dx, dy);
        const bool l = MF_IS(MF(h - 1), FSS_PHYS_AIR), r = MF_IS(MF(h + 1), FSS_PHYS_AIR);
        const int x = cx + dx, y = cy + dy;
        const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
        if(l || r || ((sinfo[mf_mat(mf)] & HI_STEAM) && p.condense >= 0)) busy = true;
        const cidx_t g = G(dx, dy);
        if(l && !(r && fss_rand(key, 1950, 0) % 2 == 0)) {
            swap(h, g, h - 1, G(dx - 1, dy));
            mark(dx - 1, dy);
        } else if(r) {
            swap(h, g, h + 1, G(dx + 1, dy));
            mark(dx + 1, dy);
        } else if((sinfo[mf_mat(mf)] & HI_STEAM) && p.condense >= 0) {
            if(fss_rand(key, 1966, 0) % 10 == 0) st(h, g, create_cell(p, p.condense, x, y, key, 0));
        }
    }

    // ---- the three scans, decoupled: every lane steps through ITS cells; bit / scan position = dy*4 + 3 - dx, walked
    //      from 63 down to 0 (rows bottom-up, columns left to right, world.cpp:1154-1157) ---------------------------------
    __device__ __forceinline__ void scans(unsigned mask) {
        int pos = FSS_CHUNK_W * FSS_CHUNK_H - 1;
        while(true) { // scan 1
            uint32_t mf = 0;
            bool found = false;
            while(pos >= 0) {
                if(!((marked >> pos) & 1ull)) {
                    mf = MF(H(FSS_CHUNK_W - 1 - (pos & 3), pos >> 2));
                    if(p.iter >= mf_iters(mf)) {
                        pv |= 1ull << pos; // :1163-1166
                    } else {
                        found = true;
                        break;
                    }
                }
                pos--;
            }
            if(!__any_sync(mask, found)) break;
            if(found) {
                scan1_body(FSS_CHUNK_W - 1 - (pos & 3), pos >> 2, mf);
                pos--;
            }
        }
        if(__all_sync(mask, (marked | pv) == ~0ull)) return;
        pos = FSS_CHUNK_W * FSS_CHUNK_H - 1;
        while(true) { // scan 2
            uint32_t mf = 0;
            bool found = false;
            while(pos >= 0) {
                if(!(((marked | pv) >> pos) & 1ull)) {
                    mf = MF(H(FSS_CHUNK_W - 1 - (pos & 3), pos >> 2));
                    if(MF_IS(mf, FSS_PHYS_SAND) || MF_IS(mf, FSS_PHYS_SOUP) || MF_IS(mf, FSS_PHYS_GAS)) {
                        found = true;
                        break;
                    }
                }
                pos--;
            }
            if(!__any_sync(mask, found)) break;
            if(found) {
                scan2_body(FSS_CHUNK_W - 1 - (pos & 3), pos >> 2, mf);
                pos--;
            }
        }
        if(!__any_sync(mask, maybe_gas)) return;
        pos = FSS_CHUNK_W * FSS_CHUNK_H - 1;
        while(true) { // scan 3
            uint32_t mf = 0;
            bool found = false;
            while(pos >= 0) {
                if(!(((marked | pv) >> pos) & 1ull)) {
                    mf = MF(H(FSS_CHUNK_W - 1 - (pos & 3), pos >> 2));
                    if(MF_IS(mf, FSS_PHYS_GAS)) {
                        found = true;
                        break;
                    }
                }
                pos--;
            }
            if(!__any_sync(mask, found)) break;
            if(found) {
                scan3_body(FSS_CHUNK_W - 1 - (pos & 3), pos >> 2, mf);
                pos--;
            }
        }
    }

    __device__ __forceinline__ void run(uint8_t* qown, unsigned mask) {
        stage_in();
        scans(mask);
        if(busy) stage_out(); // an idle visit changed nothing
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

__device__ __forceinline__ void ChunkWalkS::fire(int dx, int dy, int h, uint32_t mf) {
    SliceView v = {smf, samt, sdif, cx, cy};
    marked |= fss_rare_fire(p, v, dx, dy);
}
__device__ __forceinline__ bool ChunkWalkS::sand_react(int dx, int dy, int h, uint32_t mf) {
    SliceView v = {smf, samt, sdif, cx, cy};
    return fss_rare_sand_react(p, v, dx, dy, mf);
}
__device__ __forceinline__ void ChunkWalkS::soup_free_fall(int dx, int dy, int h, uint32_t mf, float amt) {
    SliceView v = {smf, samt, sdif, cx, cy};
    fss_rare_soup_free_fall(p, v, dx, dy, mf, amt);
}
