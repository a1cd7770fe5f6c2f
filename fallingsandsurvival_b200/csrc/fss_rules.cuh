// fss_rules.cuh -- the movement rules of World::tick as device code.
//
// One thread owns one 4x16 micro-chunk of the current checkerboard phase and walks it exactly like
// the reference walks a chunk (FallingSandSurvival/world.cpp:1154-1974): three sequential scans,
// rows bottom-up, columns left-to-right.  All 32 lanes of a warp run the scans in lock step (same
// (dx,dy) at the same time), so with the permuted column layout of fss_device.cuh every global
// access of the warp falls into one 128-byte line per plane.  Cells are updated in place in HBM
// (through L1/L2); the checkerboard gap makes the read/write windows of concurrently active
// micro-chunks disjoint, so no two threads ever touch the same word in one launch.
//
// Every rule cites the reference lines it implements.  Copies vs. live reads follow the reference:
// where the reference works on a by-value copy of a cell (`MaterialInstance tile = tiles[index]`)
// the register copy is used, where it re-reads the grid the grid is re-read, unless no store of the
// same cell visit can have changed the value in between.
#pragma once
#include "fss_device.cuh"

#define FLUID_MaxValue 0.5f
#define FLUID_MinValue 0.0005f
#define FLUID_MaxCompression 0.1f
#define FLUID_MinFlow 0.05f
#define FLUID_MaxFlow 8.0f
#define FLUID_FlowSpeed 1.0f

// host-only path counters (tests/hostsim with -DK1_STATS): how often each branch of the SAND rules is taken
#if defined(K1_STATS) && !defined(__CUDA_ARCH__)
extern unsigned long long g_k1_stats[32];
#define K1_STAT(k) (g_k1_stats[k]++)
#define g_k1_add(k, v) (g_k1_stats[k] += (unsigned long long)(v))
#define __popcll_host(v) __builtin_popcountll(v)
#else
#define K1_STAT(k) ((void)0)
#define g_k1_add(k, v) ((void)0)
#define __popcll_host(v) 0
#endif

struct SubstepParams {
    Grid g;
    const DMat* mats;
    const DTex* tex;
    int n_mat;
    int x_begin, y_begin; // origin of the first active micro-chunk of this phase
    int nx, ny;           // active micro-chunks per row / rows of active micro-chunks
    uint32_t tkey;        // fss_rng_tick_key(seed, tickCt, iter)
    int iter;
    uint32_t tick_iter;   // tickCt * 8 + iter (particle records)
    uint32_t flags;       // FSS_PARTICLES_EMIT
    int condense;         // material created by condensing steam, -1 = off
    int fire_id;          // first material flagged FSS_MAT_IS_FIRE, -1 = none
    uint32_t nothing_mf;  // mf word of Tiles::NOTHING (material 0 + its derived bits)
    fss_particle_spawn* parts;
    unsigned long long* part_count;
    uint32_t part_cap;
    // settled-tile early-out (DESIGN.md "early-out"): one byte per micro-chunk of the zone, with a guard ring.
    //   0xFF  must be processed;   v in 0..5  the chunk was verified at iter v to be left unchanged, without any
    //   RNG draw, by a sub-step: while its read window stays untouched, every sub-step with iter >= v skips it.
    uint8_t* q;   // nullptr = early-out disabled
    uint8_t* qskip; // FSS_TRACK_DIRTY: per chunk the lowest iter whose visit was skipped since the last render (0xFF = none); same layout as q
    int q_w;      // flags per row of chunks (zone chunks + 2 guards)
    int ofs_x, ofs_y; // phase offsets in chunks: the chunk of thread (gi, by) is (2*gi + ofs_x, 2*by + ofs_y)
    // Compacted worklist (mostly settled worlds): `work` = the warp tiles (32 chunks of one chunk row: by << 14 | wx) that
    // hold at least one chunk that is not verified settled, built by k_worklist right before the launch; the warps of a fixed
    // grid then stride over it.  nullptr = one CTA per K1_THREADS chunks of the (ceil(nx / K1_THREADS), ny) grid.
    const uint32_t* work;
    const uint32_t* work_n;
    // with a worklist: thread 0 of CTA 0 reports the count to the host (mapped memory, a hint for the next launch's choice) and
    // zeroes the counter the NEXT list will be counted in -- instead of a memset and a copy kernel around every launch
    uint32_t* work_n_host;
    uint32_t* work_n_next;
    // Fire on the register-window path (fss_rules_liq.cuh, fire_prepass_chunk): one byte per stored cell, written by k_fire_prepass
    // right before the sub-step's kernel and consumed (zeroed again) by it.  0 = nothing to do, FIRE_DIES, FIRE_NEW.
    uint8_t* fire_verdict;
};
#define FIRE_DIES 1u /* the fire cell goes out at its scan-1 visit (world.cpp:1191-1196, :1210-1214) */
#define FIRE_NEW 2u  /* the cell was set alight in this sub-step: visited (world.cpp:1205), it does not act */

__device__ __forceinline__ Cell cell_nothing(uint32_t nothing_mf) {
    // Tiles::NOTHING, Tiles.cpp:7; fluidAmount keeps the constructor default 2.0f (MaterialInstance.hpp:21)
    Cell c;
    c.mf = nothing_mf;
    c.col = 0u;
    c.temp = 0;
    c.amt = 2.0f;
    c.dif = 0.0f;
    return c;
}

// world.cpp:1075-1088
template <bool BRANCH_FREE>
__device__ __forceinline__ float vertical_flow(float remaining, float dest) {
    const float sum = remaining + dest;
    if(BRANCH_FREE) {
        // The three cases computed and selected: lanes fall into different cases all the time and the branchy form
        // serialises them.  Each case is the reference's own expression, so the selected value is bit-identical.  Used by
        // the liquid-only kernel; in the general kernel the extra live values cost more (spills) than the branches.
        const float mid = (FLUID_MaxValue * FLUID_MaxValue + sum * FLUID_MaxCompression) / (FLUID_MaxValue + FLUID_MaxCompression);
        const float high = (sum + FLUID_MaxCompression) / 2.0f;
        const float value = sum < 2 * FLUID_MaxValue + FLUID_MaxCompression ? mid : high;
        return sum <= FLUID_MaxValue ? FLUID_MaxValue : value;
    }
    float value;
    if(sum <= FLUID_MaxValue) {
        value = FLUID_MaxValue;
    } else if(sum < 2 * FLUID_MaxValue + FLUID_MaxCompression) {
        value = (FLUID_MaxValue * FLUID_MaxValue + sum * FLUID_MaxCompression) / (FLUID_MaxValue + FLUID_MaxCompression);
    } else {
        value = (sum + FLUID_MaxCompression) / 2.0f;
    }
    return value;
}

// `flow = std::max(flow, 0.0f); if(flow > std::min(FLUID_MaxFlow, lim)) flow = std::min(...)`
__device__ __forceinline__ float clamp_flow(float flow, float lim) {
    if(flow < 0.0f) flow = 0.0f;
    float cap = lim < FLUID_MaxFlow ? lim : FLUID_MaxFlow;
    if(flow > cap) flow = cap;
    return flow;
}

// Tiles::create(mat, x, y), Tiles.cpp:190-268, from the table's recipe; k = draws already made at the
// recipe's RNG site during this cell visit (fss_rng.h)
__device__ __noinline__ Cell create_cell(const SubstepParams& p, int m, int x, int y, uint32_t key, uint32_t k) {
    const DMat& d = p.mats[m];
    Cell c = cell_nothing(0u);
    c.mf = (uint32_t)m | d.bits;
    c.temp = d.create_temp;
    if(d.create_kind == FSS_CREATE_RAND) {
        c.col = d.create_base + ((fss_rand(key, d.create_site, k) % d.create_mod) << d.create_shift);
    } else if(d.create_kind == FSS_CREATE_TEXTURE && d.create_tex >= 0 && p.tex[d.create_tex].px) {
        const DTex t = p.tex[d.create_tex];
        c.col = t.px[(y % t.h) * t.w + (x % t.w)];
    } else {
        c.col = d.create_base;
    }
    return c;
}

// Room for n spawn records (world.cpp:1180, :1291, :1366).  false = the buffer (fss_config.particle_capacity) is full: nothing is
// reserved, the refusal is counted in part_count[1] and the caller must leave the cell in the grid (it takes the branch the
// reference takes when the free-fall test fails), so that no matter is lost; fss_drain_particles / fss_get_stats report it.
__device__ __forceinline__ bool reserve_spawns(const SubstepParams& p, unsigned n, unsigned long long& slot) {
    slot = atomicAdd(p.part_count, (unsigned long long)n);
    if(slot + n <= (unsigned long long)p.part_cap) return true;
    atomicAdd(p.part_count, 0ull - (unsigned long long)n); // hand them back: part_count[0] ends up as the number of records written
    atomicAdd(p.part_count + 1, 1ull);
    return false;
}

// Particle spawn record (world.cpp:1180, :1291, :1366): the cell leaves the grid for the host
__device__ __noinline__ void emit_particle(const SubstepParams& p, unsigned long long slot, int x, int y, uint32_t site, uint32_t key, Cell c) {
    fss_particle_spawn* r = &p.parts[slot];
    r->x = (uint32_t)x;
    r->y = (uint32_t)y;
    r->site = site;
    r->cell_key = key;
    r->tick_iter = p.tick_iter;
    r->cell.material = (uint16_t)mf_mat(c.mf);
    r->cell.moved = (c.mf & MF_MOVED) ? 1 : 0;
    r->cell.settle_count = (uint8_t)mf_settle(c.mf);
    r->cell.color = c.col;
    r->cell.temperature = c.temp;
    r->cell.fluid_amount = c.amt;
    r->cell.fluid_amount_diff = c.dif;
}

// ---- FIRE, world.cpp:1171-1216 (matched by material id; its physics type is PASSABLE) ------------------
// returns the tickVisited bits to set inside the caller's micro-chunk at (cx, cy)
// `solid_nb`: bit q = the neighbour at (xx, yy) = (q/5 - 2, q%5 - 2) is SOLID, in the reference's loop order (xx outer, yy
// inner, world.cpp:1198-1199), gathered by the whole warp before the divergent part (ChunkWalk::coop_fire_scan): the
// reference reads each neighbour once, and an ignition only rewrites a cell that was already looked at.
__device__ __noinline__ unsigned long long rule_fire(const SubstepParams& p, int cx, int cy, cidx_t i, int x, int y,
                                                     unsigned long long bit, uint32_t solid_nb) {
    const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
    unsigned long long vis = 0ull;
    Cell tile = ld_cell(p.g, i);
    if(fss_rand(key, 1172, 0) % 10 == 0) { // recolours the local copy only (never stored)
        tile.col = (((255u << 8) + 100 + fss_rand(key, 1174, 0) % 50) << 8) + 50;
    }
    if(fss_rand(key, 1179, 0) % 10 == 0) {
        unsigned long long slot; // a temporary, purely visual particle: dropped when the buffer is full
        if((p.flags & FSS_PARTICLES_EMIT) && reserve_spawns(p, 1, slot)) emit_particle(p, slot, x, y, 1180, key, tile);
    }
    uint8_t* const dirty = p.g.dirty; // World::dirty, nullptr unless FSS_TRACK_DIRTY (a rare rule: tested at run time)
    if(fss_rand(key, 1191, 0) % 150 == 0) {
        st_cell(p.g, i, cell_nothing(p.nothing_mf));
        if(dirty) dirty[i] = 1; // :1194
        return bit;
    }
    const bool found = solid_nb != 0u;
    uint32_t k1202 = 0, kcreate = 0;
    for(uint32_t rest = solid_nb; rest; rest &= rest - 1) {
        if(fss_rand(key, 1202, k1202++) % 500 == 0) {
            const int q = __ffs(rest) - 1, xx = q / 5 - 2, yy = q % 5 - 2;
            st_cell(p.g, p.g.idx(x + xx, y + yy), create_cell(p, p.fire_id, x + xx, y + yy, key, kcreate++)); // Tiles::createFire()
            if(dirty) dirty[p.g.idx(x + xx, y + yy)] = 1; // :1204
            const unsigned dx = (unsigned)(x + xx - cx), dy = (unsigned)(y + yy - cy);
            if(dx < (unsigned)FSS_CHUNK_W && dy < (unsigned)FSS_CHUNK_H) vis |= 1ull << (dy * FSS_CHUNK_W + dx);
        }
    }
    if(!found && fss_rand(key, 1210, 0) % 120 == 0) {
        st_cell(p.g, i, cell_nothing(p.nothing_mf));
        if(dirty) dirty[i] = 1; // :1212
        vis |= bit;
    }
    return vis;
}

// ---- SAND reactions, world.cpp:1251-1274 (only SAND evaluates them); true = the cell reacted ----------
// sand_react_cell: what the cell turns into (`out`), given its temperature; rule_sand_react: the same through memory.
__device__ __noinline__ bool sand_react_cell(const SubstepParams& p, int x, int y, uint32_t mf, int32_t temp, Cell& out) {
    const DMat& m = p.mats[mf_mat(mf)];
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
            out = create_cell(p, m.r_target[r], x, y, key, k); // (each matching reaction overwrites the cell: the last one stays)
            out.temp = temp;
            react = true;
        }
    }
    return react;
}
__device__ __forceinline__ bool rule_sand_react(const SubstepParams& p, cidx_t i, int x, int y, uint32_t mf) {
    Cell c;
    if(!sand_react_cell(p, x, y, mf, p.g.temp[i], c)) return false;
    st_cell(p.g, i, c);
    if(p.g.dirty) p.g.dirty[i] = 1; // :1259, :1267
    return true;
}

// ---- falling liquid leaves the grid as particles, world.cpp:1352-1374; false = no room for the records: the cell stays --
__device__ __noinline__ bool rule_soup_free_fall(const SubstepParams& p, cidx_t i, int x, int y, uint32_t mf, float amt) {
    const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
    Cell nt;
    nt.mf = mf & (MF_MAT_MASK | MF_DERIVED_MASK);
    nt.col = p.g.col[i];
    nt.temp = p.g.temp[i];
    nt.dif = 0.0f;
    int n = (int)(amt / 4);
    if(n < 1) n = 1;
    nt.amt = amt / n;
    unsigned long long slot;
    if(!reserve_spawns(p, (unsigned)n, slot)) return false;
    st_cell(p.g, i, cell_nothing(p.nothing_mf)); // :1353 setTile, which sets dirty (:1061)
    if(p.g.dirty) p.g.dirty[i] = 1;
    for(int k = 0; k < n; k++) emit_particle(p, slot + k, x, y, 1366, key, nt);
    return true;
}

// per-thread state of one micro-chunk visit.
// MODE selects a specialisation that is exact for the sub-step it is launched for (fss_abi.cu tracks which materials can
// be present in the world and picks the mode per iter):
//   K1_GENERAL  every rule
//   K1_NOGAS    no GAS and no fire material with `iterations > iter` can be present: SAND + SOUP only, no scan 3
//   K1_LIQUID   additionally no SAND material: every cell that passes the `iter < iterations` gate is SOUP (or an inert
//               SOLID / PASSABLE); a third of the instructions and registers.  4 of the 6 iters of a sand + water world.
#define K1_GENERAL 0
#define K1_NOGAS 1
#define K1_LIQUID 2
// FLOW: also keep World::flowX / flowY (FSS_TRACK_FLOW); a compile-time switch because the four extra read-modify-writes
// sit in the hottest rule and cost 10 % of the liquid kernel even when a run-time test skips them.
// DIRTY: also keep World::dirty (FSS_TRACK_DIRTY): one byte store next to every store the reference accompanies with
// `dirty[...] = true`; compile-time for the same reason.
template <int MODE, bool FLOW = false, bool DIRTY = false>
struct ChunkWalk {
    static constexpr bool LIQ = MODE == K1_LIQUID;
    static constexpr bool GASFIRE = MODE == K1_GENERAL;
    // K1_NOGAS: a SAND cell's first three loads are requested before the warp meets in the scan-1 prologue (measured on c2: 1 %
    // off iters 0 / 1; not measured in the general kernel, which keeps them inside the rule)
    static constexpr bool HOIST_SAND = MODE == K1_NOGAS;
    const SubstepParams& p;
    const uint32_t* sinfo; // shared: HI_* word per material
    int cx, cy;
    // storage offsets: own column dx is at pb + 32*dx (the four own columns share one 8-column group, so
    // they are 32 words apart in the permuted layout); pl / pr are the columns left of dx = 0 / right of dx = 3
    uint32_t pb, pl, pr;
    cidx_t row0; // offset of stored row cy
    // the cell being visited: i = its index, il / ir = the indices of (x-1, y) / (x+1, y)
    cidx_t il, ir;
    unsigned wmask;         // the lanes of this warp that walk a chunk (the others left the kernel: out of range or settled)
    bool busy;              // this visit stored to the grid or drew a random number (else it was the identity)
    bool maybe_gas;         // an unvisited GAS cell may exist in the chunk (else scan 3 has nothing to do)
    unsigned long long vis; // tickVisited restricted to the chunk, bit dy*4+dx (flags set outside the chunk are
                            // never read back: the other buffer is cleared before reuse, world.cpp:1125-1131, :1998)

    __device__ __forceinline__ ChunkWalk(const SubstepParams& p_, const uint32_t* si, int cx_, int cy_, unsigned wmask_)
        : p(p_), sinfo(si), cx(cx_), cy(cy_), wmask(wmask_), busy(false), maybe_gas(false), vis(0ull) {
        pb = fss_perm((uint32_t)cx);
        pl = fss_perm((uint32_t)(cx - 1));
        pr = fss_perm((uint32_t)(cx + FSS_CHUNK_W));
        row0 = (cidx_t)(cy - p.g.sy0) * p.g.pitch;
        il = ir = 0;
    }
    // index of (cx+dx, cy+dy) for dx in 0..3; also sets il / ir
    __device__ __forceinline__ cidx_t locate(int dx, int dy) {
        const cidx_t rb = row0 + (cidx_t)dy * p.g.pitch;
        const uint32_t pc = pb + 32u * (uint32_t)dx;
        il = rb + (dx > 0 ? pc - 32u : pl);
        ir = rb + (dx < FSS_CHUNK_W - 1 ? pc + 32u : pr);
        return rb + pc;
    }

    __device__ __forceinline__ cidx_t I(int x, int y) const { return p.g.idx(x, y); }
    __device__ __forceinline__ void D(cidx_t i) const { // dirty[...] = true
        if(DIRTY) p.g.dirty[i] = 1;
    }
    __device__ __forceinline__ void mark(int x, int y) {
        unsigned dx = (unsigned)(x - cx), dy = (unsigned)(y - cy);
        if(dx < (unsigned)FSS_CHUNK_W && dy < (unsigned)FSS_CHUNK_H) vis |= 1ull << (dy * FSS_CHUNK_W + dx);
    }
    __device__ __forceinline__ bool can_move_into(uint32_t dst_mf, uint32_t self_mf) const {
        // world.cpp:1276-1285: AIR, or not SOLID and lighter.  `density < density` is decided on the density RANKS cached in
        // the two mf words (equal densities share a rank): no table lookup behind the neighbour's load.
        return MF_IS(dst_mf, FSS_PHYS_AIR) || (!MF_IS(dst_mf, FSS_PHYS_SOLID) && (dst_mf & MF_RANK_MASK) < (self_mf & MF_RANK_MASK));
    }
    __device__ __forceinline__ void swap_cells(cidx_t a, cidx_t b) const {
        Cell ca = ld_cell(p.g, a), cb = ld_cell(p.g, b);
        st_cell(p.g, a, cb);
        st_cell(p.g, b, ca);
    }

    // `neighbour created (copy of mat/colour/temperature, amount 0) if AIR; neighbour.diff += flow`
    // (world.cpp:1399-1403, :1441-1445, :1471-1475, :1503-1507)
    __device__ __forceinline__ void give(cidx_t in, uint32_t nmf, float ndif, float flow, uint32_t self_mf, cidx_t i) const {
        if(MF_IS(nmf, FSS_PHYS_AIR)) {
            Cell s;
            s.mf = self_mf & (MF_MAT_MASK | MF_DERIVED_MASK);
            s.col = p.g.col[i];
            s.temp = p.g.temp[i];
            s.amt = 0.0f;
            s.dif = 0.0f + flow;
            st_cell(p.g, in, s);
        } else {
            p.g.dif[in] = ndif + flow;
        }
    }

    // ---- SOUP, scan 1, world.cpp:1338-1537 -----------------------------------------------------------------
    __device__ __forceinline__ void soup1(cidx_t i, int x, int y, uint32_t mf) {
        const Grid& g = p.g;
        const cidx_t pitch = g.pitch;
        const float amt = g.amt[i];
        if(amt == 0.0f) return;      // :1344
        if(amt < FLUID_MinValue) {   // :1346-1350
            busy = true;
            g.amt[i] = 0.0f;
            return;
        }
        if((p.flags & FSS_PARTICLES_EMIT) && (double)amt > 0.005 && MF_IS(g.mf[i + pitch], FSS_PHYS_AIR) &&
           MF_IS(g.mf[i + 2 * pitch], FSS_PHYS_AIR) && MF_IS(g.mf[i + 3 * pitch], FSS_PHYS_AIR) &&
           MF_IS(g.mf[i + 4 * pitch], FSS_PHYS_AIR)) {
            if(rule_soup_free_fall(p, i, x, y, mf, amt)) {
                busy = true;
                return;
            }
        }
        if(mf & MF_MOVED) return;    // :1376 (moved == settled for liquids)
        busy = true; // an unsettled liquid cell always changes: it flows, or its settle counter advances
        const float start = amt;
        float remaining = amt;
        const uint32_t mat = mf_mat(mf);
        const int iter = p.iter;
        // Everything the four flows can need, issued as one batch of independent loads (one exposed memory
        // latency instead of a dozen chained ones).  None of these words is written by this visit before its use:
        // each flow only stores to the neighbour it feeds.
        const cidx_t ib = i + pitch, it = i - pitch;
        float tdif = g.dif[i];
        const uint32_t bmf = g.mf[ib], lmf = g.mf[il], rmf = g.mf[ir], tmf = g.mf[it];
        const float bamt = g.amt[ib], lamt = g.amt[il], ramt = g.amt[ir], tamt = g.amt[it];
        const float bdif = g.dif[ib], ldif = g.dif[il], rdif = g.dif[ir], updif = g.dif[it];

        const bool air_below = MF_IS(bmf, FSS_PHYS_AIR); // :1381
        if((air_below && iter <= 2) || mf_mat(bmf) == mat) { // :1384-1405
            const float dst = MF_IS(bmf, FSS_PHYS_SOUP) ? bamt : 0.0f;
            float flow = vertical_flow<!GASFIRE>(start, dst) - dst;
            flow = clamp_flow(flow, start);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(ib, bmf, bdif, flow, mf, i);
            }
            if(FLOW) g.fy[i] = g.fy[i] + flow; // :1405 flowY[index] += flow
        } else if(iter == 0 && MF_IS(bmf, FSS_PHYS_SOUP)) { // :1406-1412 (different material: the branch above failed)
            const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
            if(fss_rand(key, 1407, 0) % 10 == 0) {
                swap_cells(i, ib);
                return;
            }
        }
        if(remaining < FLUID_MinValue) { // :1414-1418
            g.dif[i] = tdif - remaining;
            return;
        }

        // :1420-1424
        const bool can_left = (MF_IS(lmf, FSS_PHYS_AIR) || mf_mat(lmf) == mat) && !air_below;
        const bool can_right = (MF_IS(rmf, FSS_PHYS_AIR) || mf_mat(rmf) == mat) && !air_below;
        if(can_left) { // :1426-1448
            const float dst = MF_IS(lmf, FSS_PHYS_SOUP) ? lamt : 0.0f;
            float flow = (remaining - dst) / (can_right ? 3.0f : 2.0f);
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(il, lmf, ldif, flow, mf, i);
            }
            if(FLOW) g.fx[i] = g.fx[i] - flow; // :1447
        }
        if(remaining < FLUID_MinValue) {
            g.dif[i] = tdif - remaining;
            return;
        }
        if(can_right) { // :1456-1478 (`canMoveLeft ? 2.0f : 2.0f`)
            const float dst = MF_IS(rmf, FSS_PHYS_SOUP) ? ramt : 0.0f;
            float flow = (remaining - dst) / 2.0f;
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(ir, rmf, rdif, flow, mf, i);
            }
            if(FLOW) g.fx[i] = g.fx[i] + flow; // :1477
        }
        if(remaining < FLUID_MinValue) {
            g.dif[i] = tdif - remaining;
            return;
        }
        // :1486-1516
        if(MF_IS(tmf, FSS_PHYS_AIR) || mf_mat(tmf) == mat) {
            const float dst = MF_IS(tmf, FSS_PHYS_SOUP) ? tamt : 0.0f;
            float flow = remaining - vertical_flow<!GASFIRE>(remaining, dst);
            flow = clamp_flow(flow, remaining);
            if(flow != 0) {
                remaining -= flow;
                tdif -= flow;
                give(it, tmf, updif, flow, mf, i);
            }
            if(FLOW) g.fy[i] = g.fy[i] - flow; // :1509
        } else if(iter == 0 && MF_IS(tmf, FSS_PHYS_SOUP)) {
            const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
            if(fss_rand(key, 1511, 0) % 10 == 0) {
                Cell top = ld_cell(g, it);
                Cell me = ld_cell(g, i);
                me.dif = tdif;
                st_cell(g, i, top);
                st_cell(g, it, me);
                return;
            }
        }
        if(remaining < FLUID_MinValue) {
            g.dif[i] = tdif - remaining;
            return;
        }
        uint32_t nmf = mf;
        if(start == remaining) { // :1524-1528
            const uint32_t settle = (mf_settle(mf) + 1u) & 0xFFu;
            nmf = (mf & 0x00FFFFFFu) | (settle << MF_SETTLE_SHIFT);
            if(settle >= 10u) nmf |= MF_MOVED;
        } else {                 // :1529-1535: the by-value copies decide (a neighbour seeded above still counts as AIR)
            D(i);                // :1530
            if(MF_IS(tmf, FSS_PHYS_SOUP) && (tmf & MF_MOVED)) g.mf[it] = tmf & ~MF_MOVED;
            if(MF_IS(bmf, FSS_PHYS_SOUP) && (bmf & MF_MOVED)) g.mf[ib] = bmf & ~MF_MOVED;
            if(MF_IS(lmf, FSS_PHYS_SOUP) && (lmf & MF_MOVED)) g.mf[il] = lmf & ~MF_MOVED;
            if(MF_IS(rmf, FSS_PHYS_SOUP) && (rmf & MF_MOVED)) g.mf[ir] = rmf & ~MF_MOVED;
        }
        if(nmf != mf) g.mf[i] = nmf; // :1537 `tiles[index] = tile`
        g.dif[i] = tdif;
    }

    // ---- SAND, scan 1, world.cpp:1218-1337 -----------------------------------------------------------------
    __device__ __forceinline__ void sand1(cidx_t i, int x, int y, uint32_t mf, unsigned long long bit, const uint32_t* pre = nullptr) {
        const Grid& g = p.g;
        const uint32_t mat = mf_mat(mf);
        // :1223-1249 interactions: `interact` is false for every material (Materials.cpp:128, :144)
        if(HI_NREACT(sinfo[mat])) {
            busy = true; // its behaviour depends on the temperature plane, which tickTemperature rewrites
            if(rule_sand_react(p, i, x, y, mf)) {
                vis |= bit; // (a product is marked visited, so even a GAS product gets no scan 2 / 3 visit)
                return;
            }
        }
        const cidx_t ib = i + g.pitch;
        const cidx_t ibl = il + p.g.pitch, ibr = ir + p.g.pitch;
        const uint32_t bmf = HOIST_SAND ? pre[0] : g.mf[ib], blmf = HOIST_SAND ? pre[1] : g.mf[ibl], brmf = HOIST_SAND ? pre[2] : g.mf[ibr];
        const bool can_b = can_move_into(bmf, mf), can_l = can_move_into(blmf, mf), can_r = can_move_into(brmf, mf);
        K1_STAT(10); // sand1 visits
        if(!can_b) return; // :1276, :1287
        K1_STAT(11); // can move below
        busy = true;
        // the cell will move in 19 of 20 cases: its words and those of the cell below are requested now, so that the loads
        // are in flight while the draw is computed
        Cell tile = ld_cell(g, i);
        Cell below = ld_cell(g, ib);
        const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
        if((can_l || can_r) && fss_rand(key, 1287, 0) % 20 == 0) return;
        if(GASFIRE && MF_IS(bmf, FSS_PHYS_GAS)) maybe_gas = true;
        unsigned long long slot = 0;
        const bool free_fall = (p.flags & FSS_PARTICLES_EMIT) && MF_IS(bmf, FSS_PHYS_AIR) && MF_IS(g.mf[ib + g.pitch], FSS_PHYS_AIR) &&
                               MF_IS(g.mf[ib + 2 * g.pitch], FSS_PHYS_AIR) && MF_IS(g.mf[ib + 3 * g.pitch], FSS_PHYS_AIR) &&
                               reserve_spawns(p, 1, slot); // (no room for the record: the cell keeps falling in the grid)
        st_cell(g, i, below); // :1289 / :1296
        D(i);                 // :1289 setTile sets dirty (:1061); :1297
        if(free_fall) {
            emit_particle(p, slot, x, y, 1291, key, tile); // :1288-1294 free fall: the cell leaves the grid
        } else {
            if(fss_rand(key, 1300, 0) % 2 == 0) tile.mf |= MF_MOVED;
            st_cell(g, ib, tile); // :1306
            D(ib);                // :1307
            mark(x, y + 1);
        }
        if(fss_rand(key, 1313, 0) % 2 == 0) { // :1313-1335 transmit movement (x-1, x+1 are inside the world: zone margin)
            if(MF_IS(blmf, FSS_PHYS_SAND) && fss_rand(key, 1316, 0) % 2 == 0 && !(blmf & MF_MOVED)) g.mf[ibl] = blmf | MF_MOVED;
            if(MF_IS(brmf, FSS_PHYS_SAND) && fss_rand(key, 1327, 0) % 2 == 0 && !(brmf & MF_MOVED)) g.mf[ibr] = brmf | MF_MOVED;
        }
    }

    // ---- scan 1, world.cpp:1154-1665 -------------------------------------------------------------------------
    // The 5x5 neighbourhood of every active fire cell of the warp, gathered by all lanes together: lane k loads the k-th
    // neighbour of one fire cell at a time (25 independent loads in flight) instead of the fire lane walking 25
    // dependent loads on its own while the other 31 lanes wait (ncu on c3: 20 % of all stall samples at 1.3 lanes).
    __device__ __forceinline__ uint32_t coop_fire_scan(unsigned fire_lanes, int x, int y) const {
        const int lane = threadIdx.x & 31;
        const int rank = __popc(wmask & ((1u << lane) - 1u)), n = __popc(wmask);
        uint32_t mine = 0u;
        while(fire_lanes) {
            const int src = __ffs(fire_lanes) - 1;
            fire_lanes &= fire_lanes - 1;
            const int fx = __shfl_sync(wmask, x, src), fy = __shfl_sync(wmask, y, src);
            uint32_t part = 0u;
            for(int q = rank; q < 25; q += n)
                if(MF_IS(p.g.mf[p.g.idx(fx + q / 5 - 2, fy + q % 5 - 2)], FSS_PHYS_SOLID)) part |= 1u << q;
            const uint32_t all = __reduce_or_sync(wmask, part);
            if(lane == src) mine = all;
        }
        return mine;
    }

    __device__ __forceinline__ void scan1_cell(int dx, int dy) {
        const unsigned long long bit = 1ull << (dy * FSS_CHUNK_W + dx);
        if(!LIQ) {
            // All lanes of the warp go through the prologue together and meet again (ballot / syncwarp) before the
            // divergent rule bodies: measured 14 % faster than letting early-returning lanes drift ahead inside a row
            // (they then issue their loads separately and the lock step that makes accesses coalesce is lost).
            const int x = cx + dx, y = cy + dy;
            const bool seen = (vis & bit) != 0ull; // :1161
            const cidx_t i = locate(dx, dy);
            const uint32_t mf = seen ? 0u : p.g.mf[i];
            const bool active = !seen && p.iter < mf_iters(mf);
            uint32_t solid_nb = 0u;
            // the three cells a SAND cell looks at first, requested before the warp meets: every rule body that follows finds
            // its first loads on the way instead of each divergent group exposing its own latency
            uint32_t pre[3] = {0u, 0u, 0u};
            if(HOIST_SAND && active && MF_IS(mf, FSS_PHYS_SAND)) {
                pre[0] = p.g.mf[i + p.g.pitch];
                pre[1] = p.g.mf[il + p.g.pitch];
                pre[2] = p.g.mf[ir + p.g.pitch];
            }
            if(GASFIRE) {
                const unsigned fire_lanes = __ballot_sync(wmask, active && (mf & MF_FIRE));
                if(fire_lanes) solid_nb = coop_fire_scan(fire_lanes, x, y);
            } else {
                __syncwarp(wmask);
            }
            if(seen) return;
            if(!active) { // :1163-1166
                vis |= bit;
                return;
            }
            if(HOIST_SAND && MF_IS(mf, FSS_PHYS_SAND)) {
                sand1(i, x, y, mf, bit, pre);
                return;
            }
            scan1_body(i, x, y, mf, bit, solid_nb);
        } else {
            if(vis & bit) return; // :1161
            const int x = cx + dx, y = cy + dy;
            const cidx_t i = locate(dx, dy);
            const uint32_t mf = p.g.mf[i];
            if(p.iter >= mf_iters(mf)) { // :1163-1166
                vis |= bit;
                return;
            }
            if(MF_IS(mf, FSS_PHYS_SOUP)) soup1(i, x, y, mf);
        }
    }

    __device__ __forceinline__ void scan1_body(cidx_t i, int x, int y, uint32_t mf, unsigned long long bit, uint32_t solid_nb) {
        if(LIQ) {
            if(MF_IS(mf, FSS_PHYS_SOUP)) soup1(i, x, y, mf);
            return;
        }
        if(GASFIRE && MF_IS(mf, FSS_PHYS_GAS)) maybe_gas = true;
        if(GASFIRE && (mf & MF_FIRE)) { // fss_set_materials guarantees a fire material is PASSABLE: no other branch follows
            busy = true;
            vis |= rule_fire(p, cx, cy, i, x, y, bit, solid_nb);
            return;
        }
        if(MF_IS(mf, FSS_PHYS_SAND)) {
            sand1(i, x, y, mf, bit);
        } else if(MF_IS(mf, FSS_PHYS_SOUP)) {
            soup1(i, x, y, mf);
        } else if(GASFIRE && MF_IS(mf, FSS_PHYS_GAS)) { // :1647-1663
            const Grid& g = p.g;
            const cidx_t ia = i - g.pitch;
            const uint32_t amf = g.mf[ia], almf = g.mf[il - g.pitch], armf = g.mf[ir - g.pitch];
            if(MF_IS(amf, FSS_PHYS_AIR)) {
                busy = true;
                const bool open = MF_IS(almf, FSS_PHYS_AIR) || MF_IS(armf, FSS_PHYS_AIR);
                if(!(open && fss_rand(fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y), 1654, 0) % 2 == 0)) {
                    swap_cells(i, ia);
                    D(i); // :1656, :1659
                    D(ia);
                    mark(x, y - 1);
                }
            }
        }
    }

    // ---- SAND, scan 2, world.cpp:1682-1807 -----------------------------------------------------------------
    __device__ __forceinline__ void sand2(cidx_t i, int x, int y, uint32_t mf, unsigned long long bit) {
        const Grid& g = p.g;
        const cidx_t pitch = g.pitch;
        const uint32_t mat = mf_mat(mf);
        const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
        K1_STAT(0); // sand2 visits
        if(!(mf & MF_MOVED)) K1_STAT(1); // ... of stopped cells
        // A stopped cell (moved == false) only changes if its pillar draw hits: `rand() % chance == 0`, chance = 1000 / unstable,
        // unstable in 2..10 (:1710-1716); with no open diagonal it stays as it is as well.  So when the visit is already known
        // not to be idle (busy), the draw is looked at first: if it is not a multiple of ANY of the nine possible chances the
        // cell stays stopped whatever the probe would count, and the visit ends before anything is loaded (19 visits in 20).
        // (While busy is still false everything below runs, so that "no draw could matter here" stays an exact statement
        // for the settled-chunk flags.)
        if(!(mf & MF_MOVED) && busy) {
            const uint32_t r = fss_rand(key, 1715, 0);
            if(r % 500u && r % 333u && r % 250u && r % 200u && r % 166u && r % 142u && r % 125u && r % 111u && r % 100u) return;
            K1_STAT(2); // stopped, the draw could hit
        }
        const cidx_t ibl = il + p.g.pitch, ibr = ir + p.g.pitch;
        const uint32_t blmf = g.mf[ibl], brmf = g.mf[ibr];
        const bool can_l = can_move_into(blmf, mf), can_r = can_move_into(brmf, mf);
        if(!(can_l || can_r)) {
            K1_STAT(3); // no open diagonal
            // :1697-1734 with no open diagonal: whatever the pillar probe and its draw decide, `moved` ends up false
            if(mf & MF_MOVED) {
                busy = true;
                g.mf[i] = mf & ~MF_MOVED;
            }
            return;
        }
        const uint32_t info = sinfo[mat];
        const uint32_t slip = HI_SLIP(info);
        bool stopped = !(mf & MF_MOVED);
        if(stopped) { // :1697-1725: the 18 probe loads are independent, issue them together
            int drop = (MF_IS(blmf, FSS_PHYS_AIR) || MF_IS(brmf, FSS_PHYS_AIR)) ? 1 : 0;
            uint32_t pl[9], pr[9];
#pragma unroll
            for(int pil = 1; pil < 10; pil++) {
                pl[pil - 1] = g.mf[ibl + (cidx_t)pil * pitch];
                pr[pil - 1] = g.mf[ibr + (cidx_t)pil * pitch];
            }
#pragma unroll
            for(int pil = 0; pil < 9; pil++) drop += (MF_IS(pl[pil], FSS_PHYS_AIR) || MF_IS(pr[pil], FSS_PHYS_AIR)) ? 1 : 0;
            const int unstable = drop + 1 - (int)HI_STAB(info);
            if(unstable > 0) {
                const int chance = 1000 / unstable;
                if(chance < 1000) busy = true; // a draw decides
                if(chance < 1000 && fss_rand(key, 1715, 0) % (uint32_t)chance == 0) {
                    // tiles[index].moved = true while the by-value copy keeps moved == false; every path below
                    // either overwrites the grid cell or resets its flag (:1802), so the store is never observable
                    stopped = false;
                }
            }
            if(stopped) return; // :1727-1734, moved is already false
        }
        busy = true;
        K1_STAT(4); // moving cell with an open diagonal (or un-stuck)
        Cell tile; // requested before the draws below: the cell moves unless `should_move` fails (1 in 2*slip)
        tile.mf = mf;
        tile.col = g.col[i];
        tile.temp = g.temp[i];
        tile.amt = g.amt[i];
        tile.dif = g.dif[i];
        const bool should_move = fss_rand(key, 1736, 0) % (2u * slip) != 0;
        if(should_move) K1_STAT(5); // slides
        if(!should_move) { // :1802 `tiles[index].moved = false`
            if(mf & MF_MOVED) g.mf[i] = mf & ~MF_MOVED;
            return;
        }
        if(fss_rand(key, 1741, 0) % 2 == 0) { // :1738-1753
            const uint32_t bmf = g.mf[i + pitch];
            if(MF_IS(bmf, FSS_PHYS_SAND) && fss_rand(key, 1744, 0) % 2 == 0 && !(bmf & MF_MOVED)) g.mf[i + pitch] = bmf | MF_MOVED;
        }
        // :1755-1800, both directions through one code path.  Left: `canMoveBelowLeft && (!canMoveBelowRight || rand()%2==0)`,
        // marks (x-1, y) visited; right: marks nothing at (x+1, y) (the reference's asymmetry, :1759 vs :1780-1784)
        const bool go_left = can_l && (!can_r || fss_rand(key, 1755, 0) % 2 == 0);
        const int sx = go_left ? -1 : 1;
        const cidx_t idiag = go_left ? ibl : ibr, iside = go_left ? il : ir;
        Cell diag;
        diag.mf = go_left ? blmf : brmf;
        diag.col = g.col[idiag];
        diag.temp = g.temp[idiag];
        diag.amt = g.amt[idiag];
        diag.dif = g.dif[idiag];
        const uint32_t smf = g.mf[iside];
        if(GASFIRE && MF_IS(diag.mf, FSS_PHYS_GAS)) maybe_gas = true;
        if(MF_IS(smf, FSS_PHYS_AIR)) {
            st_cell(g, iside, diag);
            D(iside); // :1758 / :1782
            if(go_left) mark(x - 1, y);
            st_cell(g, i, cell_nothing(p.nothing_mf));
        } else {
            st_cell(g, i, diag);
            vis |= bit;
        }
        D(i); // :1761, :1764 / :1784, :1787
        if(fss_rand(key, go_left ? 1768u : 1791u, 0) % (20u * slip) == 0) tile.mf &= ~MF_MOVED;
        st_cell(g, idiag, tile);
        D(idiag); // :1775 / :1798
        mark(x + sx, y + 1);
    }

    // ---- scan 2, world.cpp:1669-1901 -------------------------------------------------------------------------
    __device__ __forceinline__ void scan2_cell(int dx, int dy) {
        const unsigned long long bit = 1ull << (dy * FSS_CHUNK_W + dx);
        const Grid& g = p.g;
        const int x = cx + dx, y = cy + dy;
        if(vis & bit) return; // :1676 (meeting the warp here like scan1_cell does was measured: no gain)
        const cidx_t i = locate(dx, dy);
        const uint32_t mf = g.mf[i];
        if(MF_IS(mf, FSS_PHYS_SAND)) {
#ifndef EXP_SKIP_SAND2 /* timing experiment only: wrong results */
            sand2(i, x, y, mf, bit);
#endif
        } else if(MF_IS(mf, FSS_PHYS_SOUP)) { // :1808-1825
            soup_apply(i, g.amt[i], g.dif[i], bit);
        } else if(GASFIRE && MF_IS(mf, FSS_PHYS_GAS)) {
            gas2(i, x, y);
        }
    }
    // GAS in scan 2, world.cpp:1879-1899 (il / ir as locate() left them)
    __device__ __forceinline__ void gas2(cidx_t i, int x, int y) {
        const Grid& g = p.g;
        const cidx_t ial = il - g.pitch, iar = ir - g.pitch;
        const bool al = MF_IS(g.mf[ial], FSS_PHYS_AIR), ar = MF_IS(g.mf[iar], FSS_PHYS_AIR);
        if(al || ar) busy = true;
        if(al && !(ar && fss_rand(fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y), 1884, 0) % 2 == 0)) {
            swap_cells(i, ial);
            D(i); // :1886, :1889
            D(ial);
            mark(x - 1, y - 1);
        } else if(ar) {
            swap_cells(i, iar);
            D(i); // :1893, :1896
            D(iar);
            mark(x + 1, y - 1);
        }
    }


    // Scan 2's SOUP step (`fluidAmount += fluidAmountDiff`, :1808-1825) for chunk row r, run right after scan 1 has finished
    // row r-1, while the row's amt / dif lines are still in L1 (by scan 2 proper they have been evicted and come back from
    // L2 / HBM: 26 % of the liquid kernel's stall samples).  Exact because
    //  * after scan 1 of row r-1 nothing in scan 1 reads or writes the amt / dif of a SOUP cell in row r any more (the
    //    rules reach one row up and down; fire reaches further but only turns SOLID cells into fire);
    //  * an unvisited SOUP cell is not touched by any other cell's scan-2 visit before its own (SAND displaces cells of the
    //    row BELOW the moving cell, which scan 2 has already passed; GAS and SAND only move into AIR cells);
    //  * no scan-2 rule reads another cell's amt / dif, only types, and a cell that survives its update keeps its type.
    // A cell whose update would delete it (amount < MinValue -> NOTHING) changes type, which SAND / GAS neighbours to its
    // left test (`== AIR`) before scan 2 reaches it: those are left for scan 2, in the reference's order.
    __device__ __forceinline__ void early_apply_row(int r) {
        const unsigned rowvis = (unsigned)(vis >> (r * FSS_CHUNK_W)) & 0xFu;
        if(rowvis == 0xFu) return;
        const Grid& g = p.g;
        const cidx_t rb = row0 + (cidx_t)r * g.pitch + pb;
        uint32_t m[FSS_CHUNK_W];
        float a[FSS_CHUNK_W], d[FSS_CHUNK_W];
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            m[k] = g.mf[rb + 32u * k];
            a[k] = g.amt[rb + 32u * k];
            d[k] = g.dif[rb + 32u * k];
        }
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            if(((rowvis >> k) & 1u) || !MF_IS(m[k], FSS_PHYS_SOUP)) continue;
            const float amt = a[k] + d[k];
            if(amt < FLUID_MinValue) continue; // deleted by its update: scan 2 does it in order
            if(__float_as_uint(d[k]) != 0u) {
                busy = true;
                g.amt[rb + 32u * k] = amt;
                g.dif[rb + 32u * k] = 0.0f;
            }
            D(rb + 32u * k); // :1823: every liquid cell that scan 2 reaches, changed or not
            vis |= 1ull << (r * FSS_CHUNK_W + k);
        }
    }

    // SOUP in scan 2, world.cpp:1808-1825: `fluidAmount += fluidAmountDiff; fluidAmountDiff = 0`, below MinValue -> NOTHING
    __device__ __forceinline__ void soup_apply(cidx_t i, float amt0, float dif, unsigned long long bit) {
        const Grid& g = p.g;
        const float amt = amt0 + dif;
        if(__float_as_uint(dif) != 0u || amt < FLUID_MinValue) busy = true; // else the stores below rewrite the same bits
        if(amt < FLUID_MinValue) {
            st_cell(g, i, cell_nothing(p.nothing_mf));
        } else {
            g.amt[i] = amt;
            g.dif[i] = 0.0f;
        }
        D(i); // :1814, :1823
        vis |= bit;
    }

    // ---- scan 3, world.cpp:1905-1974 (the SOUP branch is empty, :1918-1943) -----------------------------------
    __device__ __forceinline__ void scan3_cell(int dx, int dy) {
        const unsigned long long bit = 1ull << (dy * FSS_CHUNK_W + dx);
        if(vis & bit) return;
        const Grid& g = p.g;
        const int x = cx + dx, y = cy + dy;
        const cidx_t i = locate(dx, dy);
        const uint32_t mf = g.mf[i];
        if(!MF_IS(mf, FSS_PHYS_GAS)) return;
        const bool l = MF_IS(g.mf[il], FSS_PHYS_AIR), r = MF_IS(g.mf[ir], FSS_PHYS_AIR);
        const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
        if(l || r || ((sinfo[mf_mat(mf)] & HI_STEAM) && p.condense >= 0)) busy = true;
        if(l && !(r && fss_rand(key, 1950, 0) % 2 == 0)) { // :1944-1964
            swap_cells(i, il);
            D(i); // :1952, :1955
            D(il);
            mark(x - 1, y);
        } else if(r) {
            swap_cells(i, ir);
            D(i); // :1959, :1962
            D(ir);
            mark(x + 1, y);
        } else if((sinfo[mf_mat(mf)] & HI_STEAM) && p.condense >= 0) { // :1965-1969
            if(fss_rand(key, 1966, 0) % 10 == 0) {
                st_cell(g, i, create_cell(p, p.condense, x, y, key, 0)); // Tiles::createWater()
                D(i);                                                    // :1968
            }
        }
    }

    static __device__ __forceinline__ void prefetch_l1(const void* q) {
#ifdef __CUDA_ARCH__
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
#else
        (void)q; // tests/hostsim compiles these rules for the host
#endif
    }

    // first touch of a chunk row (dy may be -1 .. 15), issued ahead of the scan: own four columns of mf / amt / dif
    // plus the mf of the column either side
    __device__ __forceinline__ void prefetch_row(int dy) const {
#ifndef K1_PF_LEVEL
#define K1_PF_LEVEL 1
#endif
#ifndef K1_PF_PAYLOAD
#define K1_PF_PAYLOAD 1
#endif
        if(K1_PF_LEVEL == 0) return;
        const cidx_t rb = row0 + (cidx_t)((long long)dy * (long long)p.g.pitch); // dy may be -1
        prefetch_l1(p.g.mf + rb + pl);
        prefetch_l1(p.g.mf + rb + pr);
        if(K1_PF_LEVEL >= 2 || LIQ) { // liquids flow sideways out of the chunk: the side columns' amounts (measured: helps K1_LIQUID only)
            prefetch_l1(p.g.amt + rb + pl);
            prefetch_l1(p.g.amt + rb + pr);
            prefetch_l1(p.g.dif + rb + pl);
            prefetch_l1(p.g.dif + rb + pr);
        }
#pragma unroll
        for(int k = 0; k < FSS_CHUNK_W; k++) {
            const cidx_t j = rb + pb + 32u * k;
            prefetch_l1(p.g.mf + j);
            prefetch_l1(p.g.amt + j);
            prefetch_l1(p.g.dif + j);
            if(K1_PF_LEVEL >= 3 || (K1_PF_PAYLOAD && !LIQ)) { // cells only move where SAND / GAS act: their colour and temperature
                prefetch_l1(p.g.col + j);
                prefetch_l1(p.g.temp + j);
            }
        }
    }

    __device__ __forceinline__ void scans() {
#ifndef K1_NO_PF_BELOW
        prefetch_row(FSS_CHUNK_H);
        prefetch_row(FSS_CHUNK_H - 1);
#endif
#pragma unroll 1
        for(int dy = FSS_CHUNK_H - 1; dy >= 0; dy--) {
            prefetch_row(dy - 1);
#pragma unroll 1
            for(int dx = 0; dx < FSS_CHUNK_W; dx++) scan1_cell(dx, dy);
            // (not in the general kernel: measured on c3 it costs the fire / gas iter more than it saves)
            if(!GASFIRE && dy < FSS_CHUNK_H - 1) early_apply_row(dy + 1);
        }
        if(!GASFIRE) early_apply_row(0);
        scan2_rest();
    }

    // SAND in scan 2 for scan2_sparse(): the same rule as sand2(), with everything it can need requested in ONE batch of
    // independent loads behind the cell's own mf word (the two diagonals with all their planes, the two side cells and the
    // cell below: none of them is stored to by this visit before the rule looks at it) instead of four dependent round
    // trips.  In the lock-step scan those chains hit lines the row prefetch has brought into L1; here every lane is at its
    // own position, most of them miss L1, and a lane's round is as long as its longest chain.
    struct Sand2Batch { // what one scan-2 position can need, requested together (visit2_sparse)
        Cell tile, dl, dr;
        uint32_t bmf, lmf, rmf;
    };
    __device__ __forceinline__ void sand2_sparse(cidx_t i, int x, int y, const Sand2Batch& b, unsigned long long bit) {
        const Grid& g = p.g;
        const cidx_t pitch = g.pitch;
        const uint32_t mf = b.tile.mf;
        const uint32_t mat = mf_mat(mf);
        const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)x, (uint32_t)y);
        if(!(mf & MF_MOVED) && busy) { // see sand2(): the draw cannot un-stick the cell whatever the probe counts
            if(!stopped_draw_can_hit(fss_rand(key, 1715, 0))) return;
        }
        const cidx_t ibl = il + pitch, ibr = ir + pitch, ib = i + pitch;
        const uint32_t blmf = b.dl.mf, brmf = b.dr.mf, bmf = b.bmf, lmf = b.lmf, rmf = b.rmf;
        Cell tile = b.tile;
        const Cell &dl = b.dl, &dr = b.dr;
        const bool can_l = can_move_into(blmf, mf), can_r = can_move_into(brmf, mf);
        if(!(can_l || can_r)) { // :1697-1734 with no open diagonal: `moved` ends up false
            if(mf & MF_MOVED) {
                busy = true;
                g.mf[i] = mf & ~MF_MOVED;
            }
            return;
        }
        const uint32_t info = sinfo[mat];
        const uint32_t slip = HI_SLIP(info);
        bool stopped = !(mf & MF_MOVED);
        if(stopped) { // :1697-1725
            int drop = (MF_IS(blmf, FSS_PHYS_AIR) || MF_IS(brmf, FSS_PHYS_AIR)) ? 1 : 0;
            uint32_t pl[9], pr[9];
#pragma unroll
            for(int pil = 1; pil < 10; pil++) {
                pl[pil - 1] = g.mf[ibl + (cidx_t)pil * pitch];
                pr[pil - 1] = g.mf[ibr + (cidx_t)pil * pitch];
            }
#pragma unroll
            for(int pil = 0; pil < 9; pil++) drop += (MF_IS(pl[pil], FSS_PHYS_AIR) || MF_IS(pr[pil], FSS_PHYS_AIR)) ? 1 : 0;
            const int unstable = drop + 1 - (int)HI_STAB(info);
            if(unstable > 0) {
                const int chance = 1000 / unstable;
                if(chance < 1000) busy = true; // a draw decides
                if(chance < 1000 && fss_rand(key, 1715, 0) % (uint32_t)chance == 0) stopped = false;
            }
            if(stopped) return; // :1727-1734, moved is already false
        }
        busy = true;
        const bool should_move = fss_rand(key, 1736, 0) % (2u * slip) != 0;
        if(!should_move) { // :1802 `tiles[index].moved = false`
            if(mf & MF_MOVED) g.mf[i] = mf & ~MF_MOVED;
            return;
        }
        if(fss_rand(key, 1741, 0) % 2 == 0) { // :1738-1753
            if(MF_IS(bmf, FSS_PHYS_SAND) && fss_rand(key, 1744, 0) % 2 == 0 && !(bmf & MF_MOVED)) g.mf[ib] = bmf | MF_MOVED;
        }
        // :1755-1800 (the left / right asymmetry of the visited marks as in sand2())
        const bool go_left = can_l && (!can_r || fss_rand(key, 1755, 0) % 2 == 0);
        const int sx = go_left ? -1 : 1;
        const cidx_t idiag = go_left ? ibl : ibr, iside = go_left ? il : ir;
        const Cell diag = go_left ? dl : dr;
        const uint32_t smf = go_left ? lmf : rmf;
        if(GASFIRE && MF_IS(diag.mf, FSS_PHYS_GAS)) maybe_gas = true; // (a GAS cell lands on a position that may be unvisited)
        if(MF_IS(smf, FSS_PHYS_AIR)) {
            st_cell(g, iside, diag);
            if(go_left) mark(x - 1, y);
            st_cell(g, i, cell_nothing(p.nothing_mf));
        } else {
            st_cell(g, i, diag);
            vis |= bit;
        }
        if(fss_rand(key, go_left ? 1768u : 1791u, 0) % (20u * slip) == 0) tile.mf &= ~MF_MOVED;
        st_cell(g, idiag, tile);
        mark(x + sx, y + 1);
    }

    // Scan 2 driven by the tickVisited bits instead of by position (K1_NOGAS only, entered from WinWalk<true>::run with the
    // bits its register-window scan 1 and fused liquid update left): scan 2 only acts on unvisited cells, and after scan 1 a
    // chunk has about five of them (resting SAND; liquid cells whose update deletes them), so every lane steps through ITS OWN
    // unvisited positions in the reference's order (rows bottom-up, columns left to right) and the rule bodies run with all
    // lanes that still have one -- instead of all 64 positions in lock step with one or two lanes inside the SAND rule.
    // The order inside a chunk is per lane anyway; a visit never clears a bit, and the bits it sets lie at positions the
    // order has already passed (or its own), so the set of positions to visit only shrinks as the reference's does.
    // `stopm`: positions that held an unvisited SAND cell with moved == false when scan 1 was done with them (WinWalk<true>::retire
    // sees every word of the chunk on its way out of the window).  Such a word cannot change before its scan-2 visit: a visit
    // at (x, y) writes its own cell, the cell below (row y+1: already passed, the order is bottom-up), a side cell only if that
    // is AIR, and a diagonal in row y+1.  So the pillar draw of those positions -- which ends 19 visits in 20 before anything is
    // loaded -- is evaluated here, lane by lane without touching memory, and only the positions that survive it take part in
    // the warp's steps (2.5 of the 5.4 unvisited positions of a c2 chunk are stopped cells).  Only when `busy` is already set:
    // see sand2().
    // A step is ONE round trip: the position's own words are requested together with everything the SAND rule can need.
    static __device__ __forceinline__ bool stopped_draw_can_hit(uint32_t r) {
        return !(r % 500u && r % 333u && r % 250u && r % 200u && r % 166u && r % 142u && r % 125u && r % 111u && r % 100u);
    }
    __device__ __forceinline__ void visit2_sparse(int dx, int dy, unsigned long long bit) {
        const Grid& g = p.g;
        const cidx_t i = locate(dx, dy);
        const cidx_t pitch = g.pitch, ibl = il + pitch, ibr = ir + pitch;
        Sand2Batch b;
        b.tile = ld_cell(g, i);
        b.dl = ld_cell(g, ibl);
        b.dr = ld_cell(g, ibr);
        b.bmf = g.mf[i + pitch];
        b.lmf = g.mf[il];
        b.rmf = g.mf[ir];
        if(MF_IS(b.tile.mf, FSS_PHYS_SAND)) sand2_sparse(i, cx + dx, cy + dy, b, bit);
        else if(MF_IS(b.tile.mf, FSS_PHYS_SOUP)) soup_apply(i, b.tile.amt, b.tile.dif, bit); // :1808-1825
        else if(GASFIRE && MF_IS(b.tile.mf, FSS_PHYS_GAS)) gas2(i, cx + dx, cy + dy);
    }
    // Scan 3 (world.cpp:1905-1974: only GAS acts) by the visited bits, like scan2_sparse.  `gasm`: the positions that held a GAS
    // cell when scan 1 was done with them; a GAS cell can only be somewhere else if scan 2 displaced one (maybe_gas, set by
    // sand2_sparse) -- scan 2's own GAS moves end on positions they mark visited.
    __device__ __forceinline__ void scan3_sparse(unsigned long long gasm) {
        unsigned long long todo = ~vis & (maybe_gas ? ~0ull : gasm);
        while(__any_sync(wmask, todo != 0ull)) {
            if(todo != 0ull) {
                const int top = 63 - __clzll((long long)todo);
                const int dy = top >> 2;
                const unsigned nib = (unsigned)(todo >> (dy * FSS_CHUNK_W)) & 0xFu;
                const int dx = __ffs((int)nib) - 1;
                todo &= ~(1ull << (dy * FSS_CHUNK_W + dx));
                scan3_cell(dx, dy); // (tests the visited bit itself)
            }
        }
    }
    __device__ __forceinline__ void scan2_sparse(unsigned long long stopm = 0ull) {
        unsigned long long todo = ~vis;
        if(busy) {
            unsigned long long m = todo & stopm;
            while(m != 0ull) {
                const int pos = 63 - __clzll((long long)m);
                const unsigned long long bit = 1ull << pos;
                m &= ~bit;
                const uint32_t key = fss_rng_cell_key(p.tkey, (uint32_t)(cx + (pos & (FSS_CHUNK_W - 1))), (uint32_t)(cy + (pos >> 2)));
                if(!stopped_draw_can_hit(fss_rand(key, 1715, 0))) todo &= ~bit;
            }
        }
        while(__any_sync(wmask, todo != 0ull)) {
            if(todo != 0ull) {
                const int top = 63 - __clzll((long long)todo);          // highest row with an unvisited position
                const int dy = top >> 2;
                const unsigned nib = (unsigned)(todo >> (dy * FSS_CHUNK_W)) & 0xFu;
                const int dx = __ffs((int)nib) - 1;                     // its leftmost unvisited column
                const unsigned long long bit = 1ull << (dy * FSS_CHUNK_W + dx);
                todo &= ~bit;
                if(!(vis & bit)) visit2_sparse(dx, dy, bit); // :1676 (a visit earlier in this scan may have marked it)
            }
        }
    }

    // scans 2 and 3 (also entered from WinWalk<true>::run, fss_rules_liq.cuh, after its register-window scan 1)
    __device__ __forceinline__ void scan2_rest() {
#ifdef EXP_SKIP_SCAN2 /* timing experiment only: wrong results */
        return;
#endif
        // scan 2 and scan 3 only act on cells that are not marked visited: when every lane's chunk is fully
        // visited (iter >= iterations everywhere: AIR, SOLID, spent SAND) the warp skips them
        K1_STAT(8);                                   // chunk visits entering scan 2
        g_k1_add(9, 64 - __popcll_host(vis));          // unvisited positions
        if(__all_sync(__activemask(), vis == ~0ull)) return;
#ifndef K1_PF2
#define K1_PF2 1 /* scan 2 is short per cell: its prefetch runs this many rows ahead */
#endif
#pragma unroll
        for(int k = 1; k <= K1_PF2; k++) prefetch_row(FSS_CHUNK_H - k);
        if(LIQ) {
            // only SOUP cells can be unvisited, and a SOUP cell of scan 2 touches nothing but itself: the row's twelve words
            // are loaded as one batch of independent loads (one exposed latency per row instead of up to eight)
#pragma unroll 1
            for(int dy = FSS_CHUNK_H - 1; dy >= 0; dy--) {
                if(dy >= K1_PF2) prefetch_row(dy - K1_PF2);
                const unsigned rowvis = (unsigned)(vis >> (dy * FSS_CHUNK_W)) & 0xFu;
                if(rowvis == 0xFu) continue;
                const cidx_t rb = row0 + (cidx_t)dy * p.g.pitch + pb;
                uint32_t m[FSS_CHUNK_W];
                float a[FSS_CHUNK_W], d[FSS_CHUNK_W];
#pragma unroll
                for(int k = 0; k < FSS_CHUNK_W; k++) {
                    m[k] = p.g.mf[rb + 32u * k];
                    a[k] = p.g.amt[rb + 32u * k];
                    d[k] = p.g.dif[rb + 32u * k];
                }
#pragma unroll
                for(int k = 0; k < FSS_CHUNK_W; k++)
                    if(!((rowvis >> k) & 1u) && MF_IS(m[k], FSS_PHYS_SOUP)) soup_apply(rb + 32u * k, a[k], d[k], 1ull << (dy * FSS_CHUNK_W + k));
            }
            // no GAS can be active: scan 3 has nothing to do
        } else {
#pragma unroll 1
            for(int dy = FSS_CHUNK_H - 1; dy >= 0; dy--) {
                if(dy >= K1_PF2) prefetch_row(dy - K1_PF2);
#pragma unroll 1
                for(int dx = 0; dx < FSS_CHUNK_W; dx++) scan2_cell(dx, dy);
            }
            if(GASFIRE) scan3();
        }
    }

    __device__ __forceinline__ void scan3() {
        // a GAS cell inside the chunk at scan 3 was GAS at its scan-1 visit or was swapped in by a SAND move
        if(!__any_sync(__activemask(), maybe_gas)) return;
#pragma unroll 1
        for(int dy = FSS_CHUNK_H - 1; dy >= 0; dy--) {
#pragma unroll 1
            for(int dx = 0; dx < FSS_CHUNK_W; dx++) scan3_cell(dx, dy);
        }
    }

    // early-out bookkeeping after the three scans (see SubstepParams::q)
    __device__ __forceinline__ void run(uint8_t* qown) {
        scans();
        if(!qown) return;
        if(!busy) {
            *qown = (uint8_t)p.iter;
        } else {
            // whatever this visit wrote lies inside its write window (+-2 columns, -2..+17 rows), which only the read
            // windows of the 3x3 chunks around it can see: those (and this chunk) must be visited again
#pragma unroll
            for(int a = -1; a <= 1; a++)
#pragma unroll
                for(int b = -1; b <= 1; b++) qown[a * p.q_w + b] = 0xFF;
        }
    }
};
