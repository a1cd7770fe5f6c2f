// tests/hostsim/hostsim.cpp -- the product's movement rules (fallingsandsurvival_b200/csrc/fss_rules.cuh, the code of kernel K1)
// compiled for the HOST and driven one micro-chunk "thread" at a time, so that the rule code itself can be compared with
// the oracle without a GPU (tests/test_hostsim_rules.py).  A launch of k_substep is race-free by construction (DESIGN.md
// section 2), so running its threads one after the other gives the launch's result.  What is restated here, and therefore
// NOT covered: the kernel wrapper (csrc/fss_kernels.cuh k_substep: the shared-memory table and the settled-chunk test) and
// the material-table conversion of fss_set_materials (csrc/fss_abi.cu).  Test infrastructure only.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "fss_rules.cuh"
#include "fss_rules_liq.cuh"

#ifdef K1_STATS
unsigned long long g_k1_stats[32];
extern "C" void hs_stats(unsigned long long* out, int reset) {
    memcpy(out, g_k1_stats, sizeof g_k1_stats);
    if(reset) memset(g_k1_stats, 0, sizeof g_k1_stats);
}
#endif

struct HostWorld {
    int width, height;
    Grid g;
    std::vector<uint32_t> mf, col;
    std::vector<int32_t> temp;
    std::vector<float> amt, dif, fx, fy;
    std::vector<uint8_t> dirty;
    int n_mat = 0, fire_id = -1;
    DMat mats[FSS_MAX_MATERIALS];
    DTex tex[FSS_MAX_TEXTURES];
    std::vector<uint32_t> tex_px[FSS_MAX_TEXTURES];
    uint32_t sinfo[FSS_MAX_MATERIALS];
    uint32_t seed, flags;
    int condense;
    int zx, zy, zw, zh;
    std::vector<fss_particle_spawn> parts;
    unsigned long long part_count[2] = {0, 0};
    std::vector<uint8_t> q;
    int q_w = 0, q_h = 0;
    std::vector<uint8_t> fire_verdict; // k_fire_prepass -> window kernel (csrc/fss_rules_liq.cuh fire_prepass_chunk)
};

// recompute_liq of csrc/fss_abi.cu restated from the grid's content: can fire be split off into the pre-pass?
static bool fire_static(const HostWorld* w) {
    std::vector<bool> pres((size_t)w->n_mat, false);
    for(size_t i = 0; i < w->mf.size(); i++) pres[mf_mat(w->mf[i])] = true;
    if(w->fire_id >= 0) pres[(size_t)w->fire_id] = true;
    for(bool grew = true; grew;) {
        grew = false;
        for(int m = 0; m < w->n_mat; m++) {
            if(!pres[m]) continue;
            const DMat& d = w->mats[m];
            if(mf_type(d.bits) == FSS_PHYS_SAND)
                for(int r = 0; r < d.n_react; r++)
                    if(!pres[d.r_target[r]]) pres[d.r_target[r]] = grew = true;
            if((d.flags & FSS_MAT_IS_STEAM) && w->condense >= 0 && !pres[w->condense]) pres[w->condense] = grew = true;
        }
    }
    uint32_t sand_rank = 0, fire_rank = 0xFFFFFFFFu;
    bool solid_product = false;
    for(int m = 0; m < w->n_mat; m++) {
        if(!pres[m]) continue;
        const DMat& d = w->mats[m];
        if(mf_type(d.bits) == FSS_PHYS_SAND) {
            sand_rank = std::max(sand_rank, d.bits & MF_RANK_MASK);
            for(int r = 0; r < d.n_react; r++)
                if(mf_type(w->mats[d.r_target[r]].bits) == FSS_PHYS_SOLID) solid_product = true;
        }
        if(d.bits & MF_FIRE) fire_rank = std::min(fire_rank, d.bits & MF_RANK_MASK);
    }
    return sand_rank <= fire_rank && !solid_product;
}

template <int MODE, bool FLOW, bool DIRTY>
static void run_launch(HostWorld* w, const SubstepParams& p0) {
    SubstepParams p = p0;
    if(w->fire_verdict.size() != w->mf.size()) w->fire_verdict.assign(w->mf.size(), 0);
    p.fire_verdict = w->fire_verdict.data();
    // fss_abi.cu keeps track of which materials can be present; here the grid itself is asked: is there a SAND cell whose
    // material has reactions (world.cpp:1251-1274)?  Without one no reaction can create one during the launch.
    bool sand_reacts = false;
    for(size_t i = 0; i < w->mf.size() && !sand_reacts; i++)
        sand_reacts = MF_IS(w->mf[i], FSS_PHYS_SAND) && w->mats[mf_mat(w->mf[i])].n_react > 0;
    int fire_iters = 0; // (fire_iters of fss_abi.cu restated)
    for(int m = 0; m < w->n_mat; m++)
        if(w->mats[m].bits & MF_FIRE) fire_iters = std::max(fire_iters, std::min(mf_iters(w->mats[m].bits), 6));
    const bool fstatic = fire_static(w);
    const bool window = liq_window_applies(MODE, p.iter, p.flags, FLOW, DIRTY, sand_reacts, fire_iters, fstatic); // launch_substep's choice (csrc/fss_abi.cu)
    if(window && MODE == K1_GENERAL && p.iter < fire_iters) { // k_fire_prepass
        for(int by = 0; by < p.ny; by++)
            for(int gi = 0; gi < p.nx; gi++) {
                if(p.q && p.q[(size_t)(2 * by + p.ofs_y + 1) * p.q_w + (2 * gi + p.ofs_x + 1)] <= (uint8_t)p.iter) continue;
                fire_prepass_chunk(p, p.x_begin + 2 * FSS_CHUNK_W * gi, p.y_begin + 2 * FSS_CHUNK_H * by, true, 1u);
            }
    }
    for(int by = 0; by < p.ny; by++) {
        for(int gi = 0; gi < p.nx; gi++) {
            uint8_t* qown = nullptr;
            if(p.q) { // the settled-chunk test of k_substep
                qown = p.q + (size_t)(2 * by + p.ofs_y + 1) * p.q_w + (2 * gi + p.ofs_x + 1);
                if(*qown <= (uint8_t)p.iter) continue;
            }
            const int cx = p.x_begin + 2 * FSS_CHUNK_W * gi, cy = p.y_begin + 2 * FSS_CHUNK_H * by;
            if(window) {
                uint32_t pay[WIN_PAY_WORDS > WIN_NB_WORDS ? WIN_PAY_WORDS : WIN_NB_WORDS]; // the kernel's shared-memory slice of this thread
                if(liq_window_general(MODE, sand_reacts)) {
                    int gas_iters = 0; // (gas_iters of fss_abi.cu restated)
                    for(int m = 0; m < w->n_mat; m++)
                        if(mf_type(w->mats[m].bits) == FSS_PHYS_GAS) gas_iters = std::max(gas_iters, std::min(mf_iters(w->mats[m].bits), 6));
                    if(p.iter == 0 || (MODE == K1_GENERAL && p.iter < fire_iters)) {
                        WinWalk<true, true, 2> walk(p, w->sinfo, cx, cy, 1u, pay, 1);
                        walk.run(qown);
                    } else if(p.iter < gas_iters) {
                        WinWalk<true, false, 2> walk(p, w->sinfo, cx, cy, 1u, pay, 1);
                        walk.run(qown);
                    } else {
                        WinWalk<true, false, 1> walk(p, w->sinfo, cx, cy, 1u, pay, 1);
                        walk.run(qown);
                    }
                } else if(p.iter == 0) {
                    WinWalk<MODE == K1_NOGAS, true> walk(p, w->sinfo, cx, cy, 1u, pay, 1);
                    walk.run(qown);
                } else {
                    WinWalk<MODE == K1_NOGAS, false> walk(p, w->sinfo, cx, cy, 1u, pay, 1);
                    walk.run(qown);
                }
            } else {
                ChunkWalk<MODE, FLOW, DIRTY> walk(p, w->sinfo, cx, cy, 1u);
                walk.run(qown);
            }
        }
    }
}

extern "C" {

void* hs_create(int width, int height, uint32_t seed, uint32_t flags, int condense) {
    HostWorld* w = new HostWorld();
    w->width = width;
    w->height = height;
    w->seed = seed;
    w->flags = flags;
    w->condense = condense;
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
    if(flags & FSS_TRACK_FLOW) {
        w->fx.assign(n, 0.f);
        w->fy.assign(n, 0.f);
        w->g.fx = w->fx.data();
        w->g.fy = w->fy.data();
    }
    if(flags & FSS_TRACK_DIRTY) {
        w->dirty.assign(n, 0);
        w->g.dirty = w->dirty.data();
    }
    w->g.pitch = pitch;
    w->g.sy0 = 0;
    w->g.rows = height;
    memset(w->tex, 0, sizeof w->tex);
    w->parts.resize(1 << 20);
    return w;
}

void hs_destroy(void* h) { delete(HostWorld*)h; }

// fss_set_materials (csrc/fss_abi.cu), restated: the derived mf bits, the density ranks, maxStability
int hs_set_materials(void* h, const fss_material* t, int n) {
    HostWorld* w = (HostWorld*)h;
    if(n < 1 || n > FSS_MAX_MATERIALS) return -1;
    memset(w->mats, 0, sizeof w->mats);
    std::vector<float> ds;
    for(int i = 0; i < n; i++) ds.push_back(t[i].density);
    std::sort(ds.begin(), ds.end());
    ds.erase(std::unique(ds.begin(), ds.end()), ds.end());
    w->fire_id = -1;
    for(int i = 0; i < n; i++) {
        const fss_material& m = t[i];
        DMat& d = w->mats[i];
        d.density = m.density;
        const int iters = m.iterations < 0 ? 0 : (m.iterations > 7 ? 7 : m.iterations);
        d.bits = ((uint32_t)m.physics_type << MF_TYPE_SHIFT) | ((uint32_t)iters << MF_ITERS_SHIFT) | ((m.flags & FSS_MAT_IS_FIRE) ? MF_FIRE : 0u) |
                 ((uint32_t)(std::lower_bound(ds.begin(), ds.end(), m.density) - ds.begin()) << MF_RANK_SHIFT);
        d.flags = m.flags;
        d.slip = m.slipperyness;
        d.max_stability = m.slipperyness > 0 ? (int)(8 / sqrt((double)m.slipperyness) + 1) : 255;
        d.n_react = m.n_reactions;
        for(int r = 0; r < m.n_reactions; r++) {
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
        if((m.flags & FSS_MAT_IS_FIRE) && w->fire_id < 0) w->fire_id = i;
        // the word k_substep keeps in shared memory per material (csrc/fss_kernels.cuh)
        const uint32_t stab = d.max_stability > 255 ? 255u : (uint32_t)d.max_stability;
        w->sinfo[i] = ((uint32_t)d.slip & 0xFFFFu) | (stab << 16) | (((uint32_t)d.n_react & 3u) << 24) | ((d.flags & FSS_MAT_IS_STEAM) ? HI_STEAM : 0u);
    }
    w->n_mat = n;
    return 0;
}

int hs_set_texture(void* h, int index, const uint32_t* argb, int tw, int th) {
    HostWorld* w = (HostWorld*)h;
    w->tex_px[index].assign(argb, argb + (size_t)tw * th);
    w->tex[index].px = w->tex_px[index].data();
    w->tex[index].w = tw;
    w->tex[index].h = th;
    return 0;
}

void hs_set_zone(void* h, int x, int y, int zw, int zh, int early_out) {
    HostWorld* w = (HostWorld*)h;
    w->zx = x;
    w->zy = y;
    w->zw = zw;
    w->zh = zh;
    w->q.clear();
    if(early_out) {
        w->q_w = zw / FSS_CHUNK_W + 2;
        w->q_h = zh / FSS_CHUNK_H + 2;
        w->q.assign((size_t)w->q_w * w->q_h, 0xFF);
    }
}

void hs_upload(void* h, const fss_cell* cells) { // k_upload
    HostWorld* w = (HostWorld*)h;
    for(int y = 0; y < w->height; y++)
        for(int x = 0; x < w->width; x++) {
            const fss_cell& c = cells[(size_t)y * w->width + x];
            const size_t i = w->g.idx(x, y);
            uint32_t mat = c.material < w->n_mat ? c.material : 0;
            w->mf[i] = mf_pack(mat, c.moved, c.settle_count, w->mats[mat].bits);
            w->col[i] = c.color;
            w->temp[i] = c.temperature;
            w->amt[i] = c.fluid_amount;
            w->dif[i] = c.fluid_amount_diff;
        }
    if(!w->q.empty()) std::fill(w->q.begin(), w->q.end(), 0xFF);
}

void hs_download(void* h, fss_cell* cells, float* fx, float* fy, uint8_t* dirty) { // k_download
    HostWorld* w = (HostWorld*)h;
    for(int y = 0; y < w->height; y++)
        for(int x = 0; x < w->width; x++) {
            fss_cell& c = cells[(size_t)y * w->width + x];
            const size_t i = w->g.idx(x, y);
            const uint32_t mf = w->mf[i];
            c.material = (uint16_t)mf_mat(mf);
            c.moved = (mf & MF_MOVED) ? 1 : 0;
            c.settle_count = (uint8_t)mf_settle(mf);
            c.color = w->col[i];
            c.temperature = w->temp[i];
            c.fluid_amount = w->amt[i];
            c.fluid_amount_diff = w->dif[i];
            if(fx && w->g.fx) fx[(size_t)y * w->width + x] = w->fx[i], fy[(size_t)y * w->width + x] = w->fy[i];
            if(dirty && w->g.dirty) dirty[(size_t)y * w->width + x] = w->dirty[i];
        }
}

// one (iter, tk) sub-step with the kernel variant `mode` (K1_GENERAL 0, K1_NOGAS 1, K1_LIQUID 2): launch_substep of csrc/fss_abi.cu
int hs_substep(void* h, uint32_t tick_ct, int iter, int tk, int mode) {
    HostWorld* w = (HostWorld*)h;
    SubstepParams p;
    memset(&p, 0, sizeof p);
    const int ofs_x = tk % 2, ofs_y = 1 - ((tk % 4) / 2);
    p.g = w->g;
    p.mats = w->mats;
    p.tex = w->tex;
    p.n_mat = w->n_mat;
    p.x_begin = w->zx + ofs_x * FSS_CHUNK_W;
    p.y_begin = w->zy + ofs_y * FSS_CHUNK_H;
    const int x_end = w->zx + w->zw, y_end = w->zy + w->zh;
    p.nx = p.x_begin < x_end ? (x_end - p.x_begin + 2 * FSS_CHUNK_W - 1) / (2 * FSS_CHUNK_W) : 0;
    p.ny = p.y_begin < y_end ? (y_end - p.y_begin + 2 * FSS_CHUNK_H - 1) / (2 * FSS_CHUNK_H) : 0;
    p.tkey = fss_rng_tick_key(w->seed, tick_ct, (uint32_t)iter);
    p.iter = iter;
    p.tick_iter = tick_ct * 8u + (uint32_t)iter;
    p.flags = w->flags;
    p.condense = w->condense;
    p.fire_id = w->fire_id;
    p.nothing_mf = 0u | w->mats[0].bits;
    p.parts = w->parts.data();
    p.part_count = w->part_count;
    p.part_cap = (uint32_t)w->parts.size();
    p.q = w->q.empty() ? nullptr : w->q.data();
    p.qskip = nullptr;
    p.q_w = w->q_w;
    p.ofs_x = ofs_x;
    p.ofs_y = ofs_y;
    const bool flow = w->g.fx != nullptr, dirty = w->g.dirty != nullptr;
#define HS_RUN(M)                                                          \
    do {                                                                   \
        if(dirty) {                                                        \
            if(flow) run_launch<M, true, true>(w, p);                      \
            else run_launch<M, false, true>(w, p);                         \
        } else {                                                           \
            if(flow) run_launch<M, true, false>(w, p);                     \
            else run_launch<M, false, false>(w, p);                        \
        }                                                                  \
    } while(0)
    if(mode == K1_LIQUID) HS_RUN(K1_LIQUID);
    else if(mode == K1_NOGAS) HS_RUN(K1_NOGAS);
    else if(mode == K1_GENERAL) HS_RUN(K1_GENERAL);
    else return -1;
    return 0;
}

size_t hs_drain(void* h, fss_particle_spawn* out, size_t cap) {
    HostWorld* w = (HostWorld*)h;
    size_t n = (size_t)std::min<unsigned long long>(w->part_count[0], cap);
    memcpy(out, w->parts.data(), n * sizeof(fss_particle_spawn));
    w->part_count[0] = 0;
    return n;
}

} // extern "C"
