// tests/hostsim/hostsim_particles.cpp -- the per-particle functions of the product's particle kernels
// (fallingsandsurvival_b200/csrc/fss_particles.cuh: speculate / cascade / commit, the spawn-record conversion) compiled for
// the HOST and driven like fss_tick_particles drives the multi-kernel path (csrc/fss_abi.cu), one "thread" after the other
// inside every phase, so that the round scheme as written can be compared with the oracle without a GPU
// (tests/test_hostsim_particles.py).  Restated here and therefore not covered: the host loop of fss_tick_particles, the
// sort and the compaction (cub on the device), the cooperative-launch kernel.  Test infrastructure only.
#include <stdio.h>
#include <stdlib.h>

#include <numeric>
#include <vector>

#include "fss_particles.cuh"

struct HostParticles {
    int width, height, zx, zy, zw, zh;
    Grid g;
    std::vector<uint32_t> mf, col;
    std::vector<int32_t> temp;
    std::vector<float> amt, dif;
    std::vector<uint8_t> dirty;
    DMat mats[FSS_MAX_MATERIALS];
    int n_mat = 0;
    std::vector<fss_particle> list;
    uint64_t rounds = 0, passes = 0;
};

static void spiral_init() {
    int x = 0, y = 0, dx = 0, dy = -1, n = 0;
    for(int j = 0; j < 32 * 32; j++) {
        if(-16 <= x && x <= 16 && -16 <= y && y <= 16) {
            c_spiral[n].x = (signed char)x;
            c_spiral[n].y = (signed char)y;
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
}

extern "C" {

void* hp_create(int width, int height, int zx, int zy, int zw, int zh, int track_dirty) {
    HostParticles* w = new HostParticles();
    w->width = width;
    w->height = height;
    w->zx = zx;
    w->zy = zy;
    w->zw = zw;
    w->zh = zh;
    const uint32_t pitch = (uint32_t)(width + FSS_BLOCK_COLS - 1) / FSS_BLOCK_COLS * FSS_BLOCK_COLS;
    const size_t n = (size_t)pitch * height;
    w->mf.assign(n, 0);
    w->col.assign(n, 0);
    w->temp.assign(n, 0);
    w->amt.assign(n, 0.f);
    w->dif.assign(n, 0.f);
    w->g.mf = w->mf.data();
    w->g.col = w->col.data();
    w->g.temp = w->temp.data();
    w->g.amt = w->amt.data();
    w->g.dif = w->dif.data();
    w->g.fx = w->g.fy = nullptr;
    w->g.dirty = nullptr;
    if(track_dirty) {
        w->dirty.assign(n, 0);
        w->g.dirty = w->dirty.data();
    }
    w->g.pitch = pitch;
    w->g.sy0 = 0;
    w->g.rows = height;
    spiral_init();
    return w;
}
void hp_destroy(void* h) { delete(HostParticles*)h; }

// only what the particle kernels read of a material: its cached mf bits (physics type ...)
void hp_set_materials(void* h, const fss_material* t, int n) {
    HostParticles* w = (HostParticles*)h;
    memset(w->mats, 0, sizeof w->mats);
    for(int i = 0; i < n; i++) {
        const int iters = t[i].iterations < 0 ? 0 : (t[i].iterations > 7 ? 7 : t[i].iterations);
        w->mats[i].bits = ((uint32_t)t[i].physics_type << MF_TYPE_SHIFT) | ((uint32_t)iters << MF_ITERS_SHIFT) | ((t[i].flags & FSS_MAT_IS_FIRE) ? MF_FIRE : 0u);
    }
    w->n_mat = n;
}

void hp_upload(void* h, const fss_cell* cells) {
    HostParticles* w = (HostParticles*)h;
    for(int y = 0; y < w->height; y++)
        for(int x = 0; x < w->width; x++) {
            const fss_cell& c = cells[(size_t)y * w->width + x];
            const size_t i = w->g.idx(x, y);
            w->mf[i] = mf_pack(c.material, c.moved, c.settle_count, w->mats[c.material].bits);
            w->col[i] = c.color;
            w->temp[i] = c.temperature;
            w->amt[i] = c.fluid_amount;
            w->dif[i] = c.fluid_amount_diff;
        }
}

void hp_download(void* h, fss_cell* cells, uint8_t* dirty) {
    HostParticles* w = (HostParticles*)h;
    for(int y = 0; y < w->height; y++)
        for(int x = 0; x < w->width; x++) {
            fss_cell& c = cells[(size_t)y * w->width + x];
            const size_t i = w->g.idx(x, y);
            c.material = (uint16_t)mf_mat(w->mf[i]);
            c.moved = (w->mf[i] & MF_MOVED) ? 1 : 0;
            c.settle_count = (uint8_t)mf_settle(w->mf[i]);
            c.color = w->col[i];
            c.temperature = w->temp[i];
            c.fluid_amount = w->amt[i];
            c.fluid_amount_diff = w->dif[i];
            if(dirty) dirty[(size_t)y * w->width + x] = w->g.dirty ? w->dirty[i] : 0;
        }
}

void hp_add(void* h, const fss_particle* p, size_t n) {
    HostParticles* w = (HostParticles*)h;
    w->list.insert(w->list.end(), p, p + n);
}

// spawn records -> particles at the end of the list: k_spawn_keys, a stable sort, k_spawn_to_particles
void hp_adopt(void* h, const fss_particle_spawn* spawn, size_t n) {
    HostParticles* w = (HostParticles*)h;
    if(!n) return;
    std::vector<uint64_t> key(n), key_sorted(n);
    std::vector<uint32_t> idx(n), idx_sorted(n);
    AdoptParams A;
    A.spawn = spawn;
    A.key = key.data();
    A.idx = idx.data();
    A.key_sorted = key_sorted.data();
    A.idx_sorted = idx_sorted.data();
    A.n = (uint32_t)n;
    A.zx = w->zx;
    A.zy = w->zy;
    for(uint32_t i = 0; i < n; i++) {
        blockIdx.x = i / 256;
        threadIdx.x = i % 256;
        k_spawn_keys(A);
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    for(size_t s = 0; s < n; s++) {
        key_sorted[s] = key[order[s]];
        idx_sorted[s] = idx[order[s]];
    }
    const size_t n_old = w->list.size();
    w->list.resize(n_old + n);
    A.dst = w->list.data() + n_old;
    for(uint32_t i = 0; i < n; i++) {
        blockIdx.x = i / 256;
        threadIdx.x = i % 256;
        k_spawn_to_particles(A);
    }
    blockIdx.x = threadIdx.x = 0;
}

// World::tickParticles: the rounds of fss_tick_particles (multi-kernel path).  slots_per_particle: size of the hashed tables
// (small values force collisions); order: 0 = threads of a phase in rank order, 1 = in reverse, 2 = a fixed shuffle
int hp_tick(void* h, int slots_per_particle, int order) {
    HostParticles* w = (HostParticles*)h;
    const uint32_t n = (uint32_t)w->list.size();
    w->rounds = w->passes = 0;
    if(!n) return 0;
    std::vector<fss_particle> out(n);
    std::vector<uint8_t> decided(n, 0), keep(n, 0), sure(n, 0), big(n, 0);
    std::vector<int4> act(n), seen(n), reach(n);
    std::vector<uint32_t> ran(n, 0), listed(n, 0);
    size_t slots = 16;
    while(slots < (size_t)n * (size_t)slots_per_particle) slots *= 2;
    std::vector<uint32_t> cell_tables(4 * slots);
    const int tiles_w = (w->width + FSS_PTILE - 1) / FSS_PTILE, tiles_h = (w->height + FSS_PTILE - 1) / FSS_PTILE;
    std::vector<uint32_t> tile_tables(3 * (size_t)tiles_w * tiles_h);
    uint32_t ctl[8];
    ParticleParams P;
    memset(&P, 0, sizeof P);
    P.g = w->g;
    P.mats = w->mats;
    P.n_mat = w->n_mat;
    P.width = w->width;
    P.height = w->height;
    P.zx = w->zx;
    P.zy = w->zy;
    P.zw = w->zw;
    P.zh = w->zh;
    P.in = w->list.data();
    P.out = out.data();
    P.decided = decided.data();
    P.keep = keep.data();
    P.sure = sure.data();
    P.big = big.data();
    P.act = act.data();
    P.seen = seen.data();
    P.reach = reach.data();
    P.ran = ran.data();
    P.listed = listed.data();
    P.w_act = cell_tables.data();
    P.r_act = cell_tables.data() + slots;
    P.w_pot = cell_tables.data() + 2 * slots;
    P.r_pot = cell_tables.data() + 3 * slots;
    P.res_mask = (uint32_t)(slots - 1);
    P.res_tile = tile_tables.data();
    P.w_tile = tile_tables.data() + (size_t)tiles_w * tiles_h;
    P.w_when = tile_tables.data() + 2 * (size_t)tiles_w * tiles_h;
    P.tiles_w = tiles_w;
    P.tiles_h = tiles_h;
    P.tile_mask = 0;
    P.ctl = ctl;
    P.n = n;
    P.q = nullptr;
    std::vector<uint32_t> seq(n);
    std::iota(seq.begin(), seq.end(), 0u);
    if(order == 1) std::reverse(seq.begin(), seq.end());
    if(order == 2) {
        uint32_t s = 12345u;
        for(uint32_t i = n; i > 1; i--) {
            s = s * 1664525u + 1013904223u;
            std::swap(seq[i - 1], seq[s % i]);
        }
    }
    for(uint32_t left = n; left;) {
        ctl[0] = 0u;
        ctl[1] = 0xFFFFFFFFu;
        ctl[3] = 0u;
        std::fill(cell_tables.begin(), cell_tables.end(), 0xFFFFFFFFu);
        std::fill(tile_tables.begin(), tile_tables.begin() + 2 * (size_t)tiles_w * tiles_h, 0xFFFFFFFFu);
        std::fill(tile_tables.begin() + 2 * (size_t)tiles_w * tiles_h, tile_tables.end(), 0u);
        for(uint32_t i : seq) particles_speculate_one(P, i);
        for(P.pass = 1; P.pass < FSS_PCASCADE_MAX; P.pass++) {
            for(uint32_t i : seq) particles_cascade_one<false>(P, i);
            w->passes++;
            if(ctl[3] < P.pass) break;
        }
        P.pass = 0;
        for(uint32_t i : seq) particles_cascade_one<true>(P, i);
        uint32_t now = 0;
        std::vector<std::pair<uint32_t, int4>> go;
        uint32_t why_unsure = 0, why_boxed = 0, why_reader_act = 0, why_reader_pot = 0;
        for(uint32_t i : seq) {
            if(decided[i]) continue;
            int4 a;
            if(particles_may_commit(P, i, a)) go.push_back({i, a});
            else {
                now++;
                if(getenv("HP_STATS")) { // why does it wait?
                    if(!(sure[i] & PS_SURE)) ((sure[i] & PS_BOXED) ? why_boxed : why_unsure)++;
                    else if(P.r_act[pcell_hash(act[i].y, act[i].z, P.res_mask)] < i) why_reader_act++;
                    else why_reader_pot++;
                }
            }
        }
        if(getenv("HP_STATS"))
            fprintf(stderr, "round %llu: %u undecided before, %zu act, wait: %u unsure, %u boxed, %u a lower particle looked at the target, %u may look; passes so far %llu\n",
                    (unsigned long long)w->rounds + 1, left, go.size(), why_unsure, why_boxed, why_reader_act, why_reader_pot, (unsigned long long)w->passes);
        if(getenv("HP_TWO_PHASE")) {
            // EXPERIMENT (host only, to count rounds): a sure particle whose target was merely LOOKED AT by a lower particle need not
            // wait if that particle acts in this very round (it looked before anybody wrote, which is its place in the serial
            // order).  So: everybody who does not act publishes the cells it looked at (R_stay); a particle that would act but
            // whose target is in R_stay of a lower one stays as well and publishes; repeat until nobody changes.
            std::vector<uint32_t> r_stay(slots, 0xFFFFFFFFu);
            std::vector<uint8_t> acts(n, 0);
            auto publish_reads = [&](uint32_t i) {
                fss_particle cur = P.in[i];
                SeenPublish v{r_stay.data(), P.res_mask, i, 0x7fffffff, 0x7fffffff, -1, -1, 0};
                int tx, ty, k0;
                particle_walk(P, cur, v, tx, ty, k0);
            };
            go.clear();
            now = 0;
            for(uint32_t i : seq) {
                if(decided[i]) continue;
                const int4 a = act[i];
                bool ok = (sure[i] & PS_SURE) && i < ctl[1];
                if(ok && a.x >= PA_DEPOSIT) ok = !(P.r_pot[pcell_hash(a.y, a.z, P.res_mask)] < i || P.res_tile[ptile_of(P, a.y, a.z)] < i);
                acts[i] = ok;
                if(!ok) publish_reads(i);
            }
            for(bool changed = true; changed;) {
                changed = false;
                for(uint32_t i : seq) {
                    if(decided[i] || !acts[i]) continue;
                    const int4 a = act[i];
                    if(a.x >= PA_DEPOSIT && r_stay[pcell_hash(a.y, a.z, P.res_mask)] < i) {
                        acts[i] = 0;
                        publish_reads(i);
                        changed = true;
                    }
                }
            }
            for(uint32_t i : seq) {
                if(decided[i]) continue;
                if(acts[i]) go.push_back({i, act[i]});
                else now++;
            }
        }
        for(auto& it : go) particles_commit_one(P, it.first, it.second); // after every decision: like the kernel, decisions never see this round's writes
        w->rounds++;
        if(now >= left) return -1; // a round that retired nobody
        left = now;
    }
    std::vector<fss_particle> next;
    for(uint32_t i = 0; i < n; i++)
        if(keep[i]) next.push_back(out[i]);
    w->list.swap(next);
    return 0;
}

size_t hp_count(void* h) { return ((HostParticles*)h)->list.size(); }
void hp_list(void* h, fss_particle* out) {
    HostParticles* w = (HostParticles*)h;
    if(!w->list.empty()) memcpy(out, w->list.data(), w->list.size() * sizeof(fss_particle));
}
uint64_t hp_rounds(void* h) { return ((HostParticles*)h)->rounds; }

} // extern "C"
