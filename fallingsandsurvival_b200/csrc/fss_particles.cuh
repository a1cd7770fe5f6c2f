// fss_particles.cuh -- World::particles on the device: World::tickParticles (world.cpp:2171-2311) and the
// `particles.insert(...)` half of World::tick (world.cpp:1180-1183, :1291, :1366, :1989-1994).
//
// The reference walks the list serially; a particle that lands writes one cell, and every later particle sees it.  The
// kernels keep exactly that order (list index = rank) but decide many particles per round:
//
//   k_particles_speculate  every undecided particle runs the reference's loop body against the grid as it is, read-only,
//                          and notes what it would do: nothing / leave the list / put its tile into cell T / pour its
//                          amount into cell T.  It reserves T (hashed table, atomicMin of the rank) and records the
//                          bounding box of the cells it looked at.  A particle that never looks at the grid (expired,
//                          outside the world or the tick zone, `phase`) is decided on the spot.
//   k_particles_check_cells  the same walk again (the grid has not changed); a particle that looked at a cell some
//                          EARLIER particle reserved is `unsure`: its real outcome is only known once that one has acted.
//                          An unsure particle may end up anywhere its path and the 32 x 32 landing search can reach, so it
//                          reserves that whole box in a table of 16 x 16-cell tiles.
//   k_particles_check_tiles  (twice) a particle whose looked-at box touches a tile reserved by an earlier unsure particle
//                          is unsure too: in the first pass it reserves its own box as well; one found only in the second
//                          pass has not told anybody, so it lowers the round's `cutoff` rank instead.
//   k_particles_commit     every sure particle below the cutoff acts.  All particles before it are either sure (their
//                          target is not a cell it looked at, and they act in this same round on the same grid) or
//                          unsure with a box that is disjoint from everything it looked at or writes; so its outcome is the
//                          serial one.  The lowest undecided particle is always sure, so every round retires at least one.
//
// The rounds stop when nobody is left; the list is then compacted in order (cub::DeviceSelect).
#pragma once
#include "fss_device.cuh"

#define FSS_PTILE 16        /* tile edge of the box reservations */
#define FSS_PBOX_MARGIN 18  /* landing search reach (16) + rounding */

enum { PA_KEEP = 0, PA_REMOVE = 1, PA_DEPOSIT = 2, PA_MERGE = 3 };

struct ParticleParams {
    Grid g;
    const DMat* mats;
    int n_mat;
    int width, height, zx, zy, zw, zh;
    fss_particle* in;   // the list at the start of the tick
    fss_particle* out;  // every particle after its step
    uint8_t* decided;   // 0 = still to act
    uint8_t* keep;      // 1 = stays in the list
    uint8_t* sure;      // this round
    int4* act;          // action, target x, target y, -
    int4* seen;         // box of the cells looked at this round (x0, y0, x1, y1 inclusive)
    int4* reach;        // box of everything the particle could look at or write this tick
    uint32_t* res_cell; // hashed by cell: lowest rank that wants to write it
    uint32_t res_mask;
    uint32_t* res_tile; // per 16 x 16 tile: lowest unsure rank that can reach it
    int tiles_w, tiles_h;
    uint32_t* ctl; // [0] particles left undecided after this round, [1] cutoff rank
    uint32_t n;
    uint8_t* q; // settled-chunk flags of the movement kernel (nullptr = off)
    int q_w, q_h, q_cj0;
};

__device__ __forceinline__ int pfloor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ uint32_t pcell_hash(int x, int y, uint32_t mask) {
    return fss_fmix32((uint32_t)y * 0x9E3779B1u ^ (uint32_t)x) & mask;
}

// the cells a walk looks at: the box around them, and how many
struct SeenBox {
    int x0, y0, x1, y1, count;
    __device__ SeenBox() : x0(0x7fffffff), y0(0x7fffffff), x1(-1), y1(-1), count(0) {}
    __device__ __forceinline__ void look(int x, int y) {
        x0 = min(x0, x);
        y0 = min(y0, y);
        x1 = max(x1, x);
        y1 = max(y1, y);
        count++;
    }
};
// did the walk look at a cell an earlier particle wants to write?
struct SeenReserved {
    const uint32_t* res;
    uint32_t mask, rank;
    bool hit;
    __device__ __forceinline__ void look(int x, int y) { hit = hit || res[pcell_hash(x, y, mask)] < rank; }
};

// The body of the remove_if lambda, world.cpp:2179-2310, without its writes to tiles[]: `cur` is advanced like the
// reference advances it, the write (if any) is returned.  Every tiles[] read goes through v.look().
template <class V>
__device__ int particle_walk(const ParticleParams& P, fss_particle& cur, V& v, int& tx, int& ty) {
    tx = ty = 0;
    if(cur.temporary && cur.lifetime <= 0) return PA_REMOVE; // :2182-2186
    if(cur.target_force != 0) {                              // :2188-2201
        const float tdx = cur.target_x - cur.x;
        const float tdy = cur.target_y - cur.y;
        const float normFac = sqrtf(tdx * tdx + tdy * tdy);
        cur.vx += tdx / normFac * cur.target_force;
        cur.vy += tdy / normFac * cur.target_force;
        if(normFac < 100) {
            cur.vx *= 0.95f;
            cur.vy *= 0.95f;
        }
    }
    const int lx = (int)cur.x, ly = (int)cur.y;
    if(cur.x < 0 || (int)(cur.x) >= P.width || cur.y < 0 || (int)(cur.y) >= P.height) return PA_REMOVE;   // :2206-2209
    if(!(lx >= P.zx && ly >= P.zy && lx < P.zx + P.zw && ly < P.zy + P.zh)) return PA_KEEP;              // :2211
    cur.vx += cur.ax;
    cur.vy += cur.ay;
    const int div = (int)((fabsf(cur.vx) + fabsf(cur.vy)) + 1); // :2216
    const float dvx = cur.vx / div;
    const float dvy = cur.vy / div;
    for(int i = 0; i < div; i++) {
        cur.x += dvx;
        cur.y += dvy;
        if(cur.x < 0 || (int)(cur.x) >= P.width || cur.y < 0 || (int)(cur.y) >= P.height) return PA_REMOVE; // :2225-2228
        if(cur.phase) continue;
        const int hx = (int)cur.x, hy = (int)cur.y;
        v.look(hx, hy);
        const uint32_t hit = P.g.mf[P.g.idx(hx, hy)];
        if(!MF_IS(hit, FSS_PHYS_AIR)) { // :2230
            const bool isObject = MF_IS(hit, FSS_PHYS_PASSABLE); // PhysicsType::OBJECT
            if(cur.in_object_state == 0) cur.in_object_state = isObject ? 1 : 2; // :2234-2246
            else if(cur.in_object_state == 1 && !isObject) cur.in_object_state = 2;
            if(!isObject || cur.in_object_state == 2) {
                if(cur.temporary) return PA_REMOVE; // :2249-2253
                v.look(lx, ly);
                if(!MF_IS(P.g.mf[P.g.idx(lx, ly)], FSS_PHYS_AIR)) {
                    // square spiral around the particle: the first AIR cell takes the tile, or (for a liquid) the first cell
                    // of the same material takes its amount, :2263-2290
                    const bool soup = mf_type(P.mats[cur.tile.material].bits) == FSS_PHYS_SOUP;
                    int x = 0, y = 0, dx = 0, dy = -1;
                    for(int j = 0; j < 32 * 32; j++) {
                        if(-16 <= x && x <= 16 && -16 <= y && y <= 16) {
                            const int px = (int)(cur.x + x), py = (int)(cur.y + y);
                            if(px >= 0 && px < P.width && py >= 0 && py < P.height) { // unchecked in the reference, see fss.h
                                v.look(px, py);
                                const uint32_t c = P.g.mf[P.g.idx(px, py)];
                                if(MF_IS(c, FSS_PHYS_AIR)) {
                                    tx = px;
                                    ty = py;
                                    return PA_DEPOSIT;
                                } else if(soup && mf_mat(c) == cur.tile.material) {
                                    tx = px;
                                    ty = py;
                                    return PA_MERGE;
                                }
                            }
                        }
                        if((x == y) || ((x < 0) && (x == -y)) || ((x > 0) && (x == 1 - y))) {
                            const int t = dx;
                            dx = -dy;
                            dy = t;
                        }
                        x += dx;
                        y += dy;
                    }
                    cur.vy = -4; // :2296-2299
                    cur.y -= 16;
                    return PA_KEEP;
                }
                tx = lx; // :2301-2306
                ty = ly;
                return PA_DEPOSIT;
            }
        }
    }
    if(cur.lifetime > 0) cur.lifetime--; // :2311-2313
    return PA_KEEP;
}

__device__ __forceinline__ void reserve_tiles(const ParticleParams& P, const int4 b, uint32_t rank) {
    for(int ty = b.y / FSS_PTILE; ty <= b.w / FSS_PTILE; ty++)
        for(int tx = b.x / FSS_PTILE; tx <= b.z / FSS_PTILE; tx++) atomicMin(&P.res_tile[ty * P.tiles_w + tx], rank);
}
__device__ __forceinline__ bool earlier_in_tiles(const ParticleParams& P, const int4 b, uint32_t rank) {
    bool hit = false;
    for(int ty = b.y / FSS_PTILE; ty <= b.w / FSS_PTILE; ty++)
        for(int tx = b.x / FSS_PTILE; tx <= b.z / FSS_PTILE; tx++) hit = hit || P.res_tile[ty * P.tiles_w + tx] < rank;
    return hit;
}

__global__ void __launch_bounds__(128) k_particles_speculate(const __grid_constant__ ParticleParams P) {
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    if(i >= P.n || P.decided[i]) return;
    fss_particle cur = P.in[i];
    // everything this particle could look at or write this tick, whatever the others do: its path (at most |v| + 1 cells
    // from where it is, after this tick's acceleration) and the landing search around any point of it
    float vx = cur.vx + cur.ax, vy = cur.vy + cur.ay;
    if(cur.target_force != 0) { // bound only: the exact update is in particle_walk
        vx = fabsf(vx) + fabsf(cur.target_force);
        vy = fabsf(vy) + fabsf(cur.target_force);
    }
    const float rx = fminf(fabsf(vx), 1e6f) + FSS_PBOX_MARGIN, ry = fminf(fabsf(vy), 1e6f) + FSS_PBOX_MARGIN;
    const float sx = fminf(fmaxf(cur.x, -1e6f), 1e6f), sy = fminf(fmaxf(cur.y, -1e6f), 1e6f);
    int4 reach;
    reach.x = max(0, min(P.width - 1, (int)floorf(sx - rx)));
    reach.y = max(0, min(P.height - 1, (int)floorf(sy - ry)));
    reach.z = max(0, min(P.width - 1, (int)ceilf(sx + rx)));
    reach.w = max(0, min(P.height - 1, (int)ceilf(sy + ry)));
    SeenBox v;
    int tx, ty;
    const int action = particle_walk(P, cur, v, tx, ty);
    P.out[i] = cur;
    P.act[i] = make_int4(action, tx, ty, 0);
    if(v.count == 0) { // never looked at the grid: nothing anybody does changes this outcome, and it writes nothing
        P.decided[i] = 1;
        P.keep[i] = action == PA_KEEP;
        return;
    }
    P.seen[i] = make_int4(v.x0, v.y0, v.x1, v.y1);
    P.reach[i] = reach;
    if(action >= PA_DEPOSIT) atomicMin(&P.res_cell[pcell_hash(tx, ty, P.res_mask)], i);
}

__global__ void __launch_bounds__(128) k_particles_check_cells(const __grid_constant__ ParticleParams P) {
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    if(i >= P.n || P.decided[i]) return;
    fss_particle cur = P.in[i];
    SeenReserved v{P.res_cell, P.res_mask, i, false};
    int tx, ty;
    particle_walk(P, cur, v, tx, ty);
    P.sure[i] = !v.hit;
    if(v.hit) reserve_tiles(P, P.reach[i], i);
}

// LAST = false: a particle found unsure here reserves its reach too.  LAST = true: it lowers the cutoff instead.
template <bool LAST>
__global__ void __launch_bounds__(128) k_particles_check_tiles(const __grid_constant__ ParticleParams P) {
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    if(i >= P.n || P.decided[i] || !P.sure[i]) return;
    if(!earlier_in_tiles(P, P.seen[i], i)) return;
    P.sure[i] = 0;
    if(LAST) atomicMin(&P.ctl[1], i);
    else reserve_tiles(P, P.reach[i], i);
}

__global__ void __launch_bounds__(128) k_particles_commit(const __grid_constant__ ParticleParams P) {
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    if(i >= P.n || P.decided[i]) return;
    if(!P.sure[i] || i >= P.ctl[1]) {
        atomicAdd(&P.ctl[0], 1u);
        return;
    }
    const int4 a = P.act[i];
    P.decided[i] = 1;
    P.keep[i] = a.x == PA_KEEP;
    if(a.x < PA_DEPOSIT) return;
    const cidx_t c = P.g.idx(a.y, a.z);
    const fss_cell t = P.out[i].tile;
    if(a.x == PA_DEPOSIT) { // tiles[...] = cur->tile
        Cell n;
        n.mf = mf_pack(t.material, t.moved, t.settle_count, P.mats[t.material].bits);
        n.col = t.color;
        n.temp = t.temperature;
        n.amt = t.fluid_amount;
        n.dif = t.fluid_amount_diff;
        st_cell(P.g, c, n);
    } else { // tiles[...].fluidAmount += cur->tile.fluidAmount
        P.g.amt[c] = P.g.amt[c] + t.fluid_amount;
    }
    if(P.q) { // the chunks whose read window sees this cell are no longer known to be settled (dirty_chunks in fss_abi.cu)
        const int c0 = max(0, pfloor_div(a.y - 5 - P.zx, FSS_CHUNK_W) + 1);
        const int c1 = min(P.q_w, pfloor_div(a.y + 2 - P.zx, FSS_CHUNK_W) + 2);
        const int r0 = max(0, pfloor_div(a.z - 25 - P.zy, FSS_CHUNK_H) - P.q_cj0 + 1);
        const int r1 = min(P.q_h, pfloor_div(a.z + 2 - P.zy, FSS_CHUNK_H) - P.q_cj0 + 2);
        for(int r = r0; r < r1; r++)
            for(int cc = c0; cc < c1; cc++) P.q[(size_t)r * P.q_w + cc] = 0xFF;
    }
}

// ---- spawn records -> particles -------------------------------------------------------------------------
struct AdoptParams {
    const fss_particle_spawn* spawn;
    uint64_t* key;
    uint32_t* idx;
    const uint64_t* key_sorted;
    const uint32_t* idx_sorted;
    fss_particle* dst; // first free slot of the list
    uint32_t n;
    int zx, zy;
};

__global__ void __launch_bounds__(256) k_spawn_keys(const __grid_constant__ AdoptParams A) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if(i >= A.n) return;
    const fss_particle_spawn& s = A.spawn[i];
    A.key[i] = fss_spawn_order_key(s.tick_iter, s.x, s.y, A.zx, A.zy);
    A.idx[i] = i;
}

// `new Particle(...)` at world.cpp:1180-1183, :1291, :1366 (the DO_MULTITHREADING branches); draw order: fss.h
__global__ void __launch_bounds__(256) k_spawn_to_particles(const __grid_constant__ AdoptParams A) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if(i >= A.n) return;
    const fss_particle_spawn s = A.spawn[A.idx_sorted[i]];
    uint32_t j = 0; // which of the records of this cell visit
    while(j < i && A.key_sorted[i - j - 1] == A.key_sorted[i]) j++;
    const uint32_t ry = fss_rand(s.cell_key, s.site, 2u * j), rx = fss_rand(s.cell_key, s.site, 2u * j + 1u);
    fss_particle p;
    p.x = (float)(int)s.x;
    p.ax = 0.f;
    p.target_x = p.target_y = p.target_force = 0.f;
    p.lifetime = 0;
    p.fade_time = 60; // Particle.hpp:27
    p.in_object_state = 0;
    p.phase = 0;
    p.temporary = 0;
    p.tile = s.cell;
    if(s.site == 1180) {
        p.y = (float)((int)s.y - 1);
        p.vx = (float)((int)(rx % 10) - 5) / 20.0f;
        p.vy = -((float)(int)(ry % 10) / 10.0f) / 3.0f + -0.5f;
        p.ay = 0.01f;
        p.temporary = 1;
        p.lifetime = 30;
        p.fade_time = 10;
    } else if(s.site == 1291) {
        p.y = (float)((int)s.y + 1);
        p.vx = (float)((int)(rx % 10) - 5) / 20.0f;
        p.vy = (float)(-((int)(ry % 2) + 3)) / 10.0f + 1.5f;
        p.ay = 0.1f;
    } else {
        p.y = (float)((int)s.y + 1);
        p.vx = (float)((int)(rx % 10) - 5) / 30.0f;
        p.vy = (float)(-((int)(ry % 2) + 3)) / 10.0f + 1.0f;
        p.ay = 0.1f;
    }
    A.dst[i] = p;
}
