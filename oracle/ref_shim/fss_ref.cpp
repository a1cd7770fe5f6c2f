/*
 * fss_ref.cpp -- C API around the reference's OWN World::tick / World::tickTemperature lines.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load the library built from this file (oracle/_ref/libfss_ref_*.so).
 *
 * The rule lines are not in this repository: oracle/build_ref.py extracts
 * FallingSandSurvival/world.cpp:1053-1062, :1075-1088, :1091-2018, :2027-2041, :2043-2143, :2171-2338
 * into oracle/_ref/gen/tick_lines.inc (with #line directives so __LINE__ keeps the reference's
 * numbering) and this file #includes it.  Instrumentation, all outside the rule lines:
 *   - CHUNK_W / CHUNK_H come from -D (the reference hard-codes 128, Chunk.hpp:18-19);
 *   - rand() -> fss_ref_rand(__LINE__)  (fss_ref_pre.h), keyed per (seed, tick, iter, x, y, site, k);
 *   - `FSS_REF_CTX(x, y);` appended to the three `int index = x + y * width;` lines
 *     (world.cpp:1159, :1674, :1910) so the RNG knows the cell being visited;
 *   - `&& FSS_REF_PARTICLES` appended to the two free-fall tests (world.cpp:1288, :1352);
 *   - the trailing physicsCheck block (world.cpp:2019-2026, Box2D hand-off) is cut.
 */
#include "fss_ref_world.hpp"
#include "Textures.hpp"
#include "../../include/fss.h"
#include "../../include/fss_rng.h"

#include <chrono>

/* ---- statics the extracted lines and Tiles.cpp expect --------------------------------------- */
ctpl::thread_pool* World::tickPool = nullptr;
ctpl::thread_pool* World::tickVisitedPool = nullptr;

SDL_Surface* Textures::testTexture = nullptr;
SDL_Surface* Textures::dirt1Texture = nullptr;
SDL_Surface* Textures::stone1Texture = nullptr;
SDL_Surface* Textures::smoothStone = nullptr;
SDL_Surface* Textures::cobbleStone = nullptr;
SDL_Surface* Textures::flatCobbleStone = nullptr;
SDL_Surface* Textures::smoothDirt = nullptr;
SDL_Surface* Textures::cobbleDirt = nullptr;
SDL_Surface* Textures::flatCobbleDirt = nullptr;
SDL_Surface* Textures::softDirt = nullptr;
SDL_Surface* Textures::cloud = nullptr;
SDL_Surface* Textures::gold = nullptr;
SDL_Surface* Textures::goldMolten = nullptr;
SDL_Surface* Textures::goldSolid = nullptr;
SDL_Surface* Textures::iron = nullptr;
SDL_Surface* Textures::obsidian = nullptr;
SDL_Surface* Textures::caveBG = nullptr;

/* ---- keyed RNG context ----------------------------------------------------------------------- */
static uint32_t g_seed = 0;
static int g_particles = 1;
#define FSS_REF_PARTICLES (g_particles != 0)
#define FSS_REF_CTX(x, y) fss_ref_ctx(tickCt, iter, (x), (y))

struct RngCtx {
    uint32_t cellkey = 0;
    int n = 0;
    int site[16];
    int count[16];
};
static thread_local RngCtx t_ctx;

void fss_ref_ctx(int tick_ct, int iter, int x, int y) {
    uint32_t tkey = fss_rng_tick_key(g_seed, (uint32_t)tick_ct, (uint32_t)iter);
    t_ctx.cellkey = fss_rng_cell_key(tkey, (uint32_t)x, (uint32_t)y);
    t_ctx.n = 0;
}

int fss_ref_rand(int site) {
    RngCtx& c = t_ctx;
    int k = 0;
    int i = 0;
    for(; i < c.n; i++) {
        if(c.site[i] == site) {
            k = ++c.count[i];
            break;
        }
    }
    if(i == c.n && c.n < 16) {
        c.site[c.n] = site;
        c.count[c.n] = 0;
        c.n++;
    }
    return (int)fss_rand(c.cellkey, (uint32_t)site, (uint32_t)k);
}

/* ---- the reference's lines ------------------------------------------------------------------- */
using std::max; /* world.cpp:2258 `max(X, Y)` */
#include "tick_lines.inc"

/* ---- Game.cpp:2584-2650, the loop of the `dirty` job of Game::tick, with the names it uses bound to arguments ---- */
#ifndef SDL_ALPHA_TRANSPARENT
#define SDL_ALPHA_TRANSPARENT 0
#endif
static void ref_dirty_job(World* world, unsigned char* dpixels_ar, unsigned char* dpixelsFire_ar, unsigned char* dpixelsFlow_ar,
                          unsigned char* dpixelsEmission_ar, uint16_t* movingTiles, bool& hadDirty, bool& hadFire, bool& hadFlow) {
#include "dirty_lines.inc"
}

/* ---- C API ----------------------------------------------------------------------------------- */
#define FSS_REF_N_TEX 16
static SDL_Surface g_tex[FSS_REF_N_TEX];
static std::vector<uint32_t> g_tex_px[FSS_REF_N_TEX];
static bool g_inited = false;

struct RefParticle {
    float x, y, vx, vy;
    fss_cell cell;
    int32_t temporary;
};

struct RefWorld {
    World w;
    std::vector<RefParticle> parts;
    bool keep = false; /* leave World::particles alone after a tick (for World::tickParticles) */
};

static void to_cell(const MaterialInstance& m, fss_cell* c) {
    c->material = (uint16_t)m.mat->id;
    c->moved = m.moved ? 1 : 0;
    c->settle_count = m.settleCount;
    c->color = m.color;
    c->temperature = m.temperature;
    c->fluid_amount = m.fluidAmount;
    c->fluid_amount_diff = m.fluidAmountDiff;
}

extern "C" {

/* texture t is a deterministic 128x128 ARGB pattern; the package's synthetic_texture() makes the same */
uint32_t fssref_texture_pixel(int t, int x, int y) {
    return 0xFF000000u | (fss_fmix32((uint32_t)t * 0x9E3779B1u + (uint32_t)(y * 128 + x)) & 0x00FFFFFFu);
}

int fssref_init(unsigned material_seed) {
    if(g_inited) return 0;
    for(int t = 0; t < FSS_REF_N_TEX; t++) {
        g_tex_px[t].resize(128 * 128);
        for(int y = 0; y < 128; y++)
            for(int x = 0; x < 128; x++) g_tex_px[t][y * 128 + x] = fssref_texture_pixel(t, x, y);
        g_tex[t].w = 128;
        g_tex[t].h = 128;
        g_tex[t].pitch = 128 * 4;
        g_tex[t].pixels = g_tex_px[t].data();
    }
    /* texture indices are part of the test contract (see materials.py TEXTURE_INDEX) */
    Textures::testTexture = &g_tex[0];
    Textures::dirt1Texture = &g_tex[1];
    Textures::stone1Texture = &g_tex[2];
    Textures::smoothStone = &g_tex[3];
    Textures::cobbleStone = &g_tex[4];
    Textures::flatCobbleStone = &g_tex[5];
    Textures::smoothDirt = &g_tex[6];
    Textures::cobbleDirt = &g_tex[7];
    Textures::flatCobbleDirt = &g_tex[8];
    Textures::softDirt = &g_tex[9];
    Textures::cloud = &g_tex[10];
    Textures::gold = &g_tex[11];
    Textures::goldMolten = &g_tex[12];
    Textures::goldSolid = &g_tex[13];
    Textures::iron = &g_tex[14];
    Textures::obsidian = &g_tex[15];
    Textures::caveBG = &g_tex[0];
    srand(material_seed);
    Materials::init();
    g_inited = true;
    return 0;
}

int fssref_chunk_w(void) { return CHUNK_W; }
int fssref_chunk_h(void) { return CHUNK_H; }
int fssref_keyed_rng(void) {
#ifdef FSS_REF_LIBC_RAND
    return 0;
#else
    return 1;
#endif
}

int fssref_num_materials(void) { return (int)Materials::MATERIALS.size(); }

/* exports Material fields; the create_* recipe is not data in the reference (it is code in
 * Tiles.cpp:190-268), so those fields are left zero here and checked through fssref_create(). */
int fssref_get_material(int id, fss_material* out, char* name, int name_cap) {
    if(id < 0 || id >= (int)Materials::MATERIALS.size()) return -1;
    Material* m = Materials::MATERIALS[id];
    memset(out, 0, sizeof(*out));
    out->physics_type = m->physicsType;
    out->iterations = m->iterations;
    out->slipperyness = m->slipperyness;
    out->density = m->density;
    out->add_temp = m->addTemp;
    out->conduction_self = m->conductionSelf;
    out->conduction_other = m->conductionOther;
    out->color = m->color;
    out->flags = (m->id == Materials::FIRE.id ? FSS_MAT_IS_FIRE : 0) | (m->id == Materials::STEAM.id ? FSS_MAT_IS_STEAM : 0);
    out->n_reactions = m->react ? m->nReactions : 0;
    for(int i = 0; i < out->n_reactions && i < FSS_MAX_REACTIONS; i++) {
        out->reactions[i].type = m->reactions[i].type;
        out->reactions[i].data1 = m->reactions[i].data1;
        out->reactions[i].data2 = m->reactions[i].data2;
    }
    if(name && name_cap > 0) snprintf(name, name_cap, "%s", m->name.c_str());
    return m->interact ? 1 : 0;
}

/* Tiles::create(mat, x, y) evaluated under an explicit RNG context */
int fssref_create(int mat, int x, int y, uint32_t seed, int tick_ct, int iter, int cx, int cy, fss_cell* out) {
    if(mat < 0 || mat >= (int)Materials::MATERIALS.size()) return -1;
    g_seed = seed;
    fss_ref_ctx(tick_ct, iter, cx, cy);
    MaterialInstance m = Tiles::create(Materials::MATERIALS[mat], x, y);
    to_cell(m, out);
    return 0;
}

void* fssref_world_create(int w, int h, int zx, int zy, int zw, int zh, int threads) {
    if(!g_inited) fssref_init(1234);
    if(w <= 0 || h <= 0 || w > 65535 || h > 65535) return nullptr;
    RefWorld* r = new RefWorld();
    World& W = r->w;
    W.width = (uint16_t)w;
    W.height = (uint16_t)h;
    size_t n = (size_t)w * h;
    W.tiles = new MaterialInstance[n];
    W.flowX = new float[n]();
    W.flowY = new float[n]();
    W.prevFlowX = new float[n]();
    W.prevFlowY = new float[n]();
    W.tickVisited1 = new bool[n]();
    W.tickVisited2 = new bool[n]();
    W.dirty = new bool[n]();
    W.newTemps = new int32_t[n]();
    W.tickZone = SDL_Rect{zx, zy, zw, zh};
    for(size_t i = 0; i < n; i++) W.tiles[i] = Tiles::NOTHING;
    if(threads < 1) threads = 1;
    if(World::tickPool == nullptr) World::tickPool = new ctpl::thread_pool(threads); /* reference: 6, world.cpp:64 */
    else if(World::tickPool->size() != threads) World::tickPool->resize(threads);
    if(World::tickVisitedPool == nullptr) World::tickVisitedPool = new ctpl::thread_pool(1);
    return r;
}

void fssref_world_destroy(void* h) {
    RefWorld* r = (RefWorld*)h;
    if(!r) return;
    World& W = r->w;
    delete[] W.tiles;
    delete[] W.flowX;
    delete[] W.flowY;
    delete[] W.prevFlowX;
    delete[] W.prevFlowY;
    delete[] W.tickVisited1;
    delete[] W.tickVisited2;
    delete[] W.dirty;
    delete[] W.newTemps;
    for(Particle* p : W.particles) delete p;
    delete r;
}

int fssref_set_cells(void* h, const fss_cell* cells) {
    World& W = ((RefWorld*)h)->w;
    size_t n = (size_t)W.width * W.height;
    int nm = (int)Materials::MATERIALS.size();
    for(size_t i = 0; i < n; i++) {
        const fss_cell& c = cells[i];
        if(c.material >= nm) return -1;
        MaterialInstance& m = W.tiles[i];
        m.mat = Materials::MATERIALS[c.material];
        m.color = c.color;
        m.temperature = c.temperature;
        m.moved = c.moved != 0;
        m.settleCount = c.settle_count;
        m.fluidAmount = c.fluid_amount;
        m.fluidAmountDiff = c.fluid_amount_diff;
    }
    return 0;
}

/* World::flowX / flowY (world.hpp:99-100), accumulated by the SOUP rule (world.cpp:1405, :1447, :1477, :1509) */
int fssref_get_flow(void* h, float* fx, float* fy) {
    World& W = ((RefWorld*)h)->w;
    size_t n = (size_t)W.width * W.height;
    memcpy(fx, W.flowX, n * sizeof(float));
    memcpy(fy, W.flowY, n * sizeof(float));
    return 0;
}

int fssref_get_cells(void* h, fss_cell* cells) {
    World& W = ((RefWorld*)h)->w;
    size_t n = (size_t)W.width * W.height;
    for(size_t i = 0; i < n; i++) to_cell(W.tiles[i], &cells[i]);
    return 0;
}

static void collect_particles(RefWorld* r) {
    World& W = r->w;
    for(Particle* p : W.particles) {
        RefParticle q;
        q.x = p->x;
        q.y = p->y;
        q.vx = p->vx;
        q.vy = p->vy;
        q.temporary = p->temporary ? 1 : 0;
        to_cell(p->tile, &q.cell);
        r->parts.push_back(q);
        delete p;
    }
    W.particles.clear();
}

/* one World::tick() with RNG seed `seed`, World::tickCt = tick_ct; returns seconds spent inside tick() */
double fssref_tick(void* h, uint32_t seed, int tick_ct, int particles) {
    RefWorld* r = (RefWorld*)h;
    g_seed = seed;
    g_particles = particles;
    r->w.tickCt = tick_ct;
    auto t0 = std::chrono::steady_clock::now();
    r->w.tick();
    auto t1 = std::chrono::steady_clock::now();
    if(!r->keep) collect_particles(r);
    return std::chrono::duration<double>(t1 - t0).count();
}

double fssref_tick_temperature(void* h) {
    RefWorld* r = (RefWorld*)h;
    auto t0 = std::chrono::steady_clock::now();
    r->w.tickTemperature();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

size_t fssref_num_particles(void* h) { return ((RefWorld*)h)->parts.size(); }

/* copies and clears; record = {x, y, vx, vy} floats, cell, temporary */
size_t fssref_drain_particles(void* h, float* xyv, fss_cell* cells, int32_t* temporary, size_t cap) {
    RefWorld* r = (RefWorld*)h;
    size_t n = r->parts.size() < cap ? r->parts.size() : cap;
    for(size_t i = 0; i < n; i++) {
        xyv[4 * i + 0] = r->parts[i].x;
        xyv[4 * i + 1] = r->parts[i].y;
        xyv[4 * i + 2] = r->parts[i].vx;
        xyv[4 * i + 3] = r->parts[i].vy;
        cells[i] = r->parts[i].cell;
        temporary[i] = r->parts[i].temporary;
    }
    r->parts.clear();
    return n;
}

/* ---- World::dirty and the dirty -> pixels job (Game.cpp:2579-2650, :2751) -------------------------------- */
int fssref_get_dirty(void* h, uint8_t* out) {
    World& W = ((RefWorld*)h)->w;
    size_t n = (size_t)W.width * W.height;
    for(size_t i = 0; i < n; i++) out[i] = W.dirty[i] ? 1 : 0;
    return 0;
}

int fssref_mark_dirty(void* h, int x, int y, int w, int hh) {
    World& W = ((RefWorld*)h)->w;
    for(int yy = y; yy < y + hh; yy++)
        for(int xx = x; xx < x + w; xx++)
            if(xx >= 0 && xx < W.width && yy >= 0 && yy < W.height) W.dirty[xx + (size_t)yy * W.width] = true;
    return 0;
}

int fssref_get_material_looks(uint8_t* alpha, uint32_t* emit, int cap) {
    int n = (int)Materials::MATERIALS.size() < cap ? (int)Materials::MATERIALS.size() : cap;
    for(int i = 0; i < n; i++) {
        alpha[i] = Materials::MATERIALS[i]->alpha;
        emit[i] = Materials::MATERIALS[i]->emitColor;
    }
    return n;
}

/* pixels: 4 layers (main, fire, emission, flow) of width*height*4 bytes each, kept by the caller between calls;
 * moving: one uint16_t per material like Game::movingTiles (Game.hpp:204); info = hadDirty, hadFire, hadFlow */
double fssref_render_dirty(void* h, uint8_t* main_px, uint8_t* fire_px, uint8_t* emission_px, uint8_t* flow_px, uint16_t* moving, uint32_t* info) {
    World& W = ((RefWorld*)h)->w;
    bool hadDirty = false, hadFire = false, hadFlow = false;
    auto t0 = std::chrono::steady_clock::now();
    for(int i = 0; i < Materials::nMaterials; i++) moving[i] = 0;                                /* Game.cpp:2579 */
    ref_dirty_job(&W, main_px, fire_px, flow_px, emission_px, moving, hadDirty, hadFire, hadFlow);
    if(hadDirty) memset(W.dirty, false, (size_t)W.width * W.height);                             /* Game.cpp:2751 */
    auto t1 = std::chrono::steady_clock::now();
    info[0] = hadDirty;
    info[1] = hadFire;
    info[2] = hadFlow;
    return std::chrono::duration<double>(t1 - t0).count();
}

/* ---- World::particles kept alive for World::tickParticles (world.cpp:2171-2338) ------------------------ */
void fssref_keep_particles(void* h, int on) { ((RefWorld*)h)->keep = on != 0; }

double fssref_tick_particles(void* h) {
    RefWorld* r = (RefWorld*)h;
    auto t0 = std::chrono::steady_clock::now();
    r->w.tickParticles();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

size_t fssref_live_particles(void* h) { return ((RefWorld*)h)->w.particles.size(); }

size_t fssref_get_particles(void* h, fss_particle* out, size_t cap) {
    World& W = ((RefWorld*)h)->w;
    size_t n = W.particles.size() < cap ? W.particles.size() : cap;
    for(size_t i = 0; i < n; i++) {
        const Particle* p = W.particles[i];
        fss_particle& q = out[i];
        memset(&q, 0, sizeof q);
        q.x = p->x; q.y = p->y; q.vx = p->vx; q.vy = p->vy; q.ax = p->ax; q.ay = p->ay;
        q.target_x = p->targetX; q.target_y = p->targetY; q.target_force = p->targetForce;
        q.lifetime = p->lifetime; q.fade_time = p->fadeTime;
        q.in_object_state = p->inObjectState; q.phase = p->phase ? 1 : 0; q.temporary = p->temporary ? 1 : 0;
        to_cell(p->tile, &q.tile);
    }
    return n;
}

int fssref_add_particles(void* h, const fss_particle* in, size_t n) {
    World& W = ((RefWorld*)h)->w;
    int nm = (int)Materials::MATERIALS.size();
    for(size_t i = 0; i < n; i++) {
        const fss_particle& q = in[i];
        if(q.tile.material >= nm) return -1;
        MaterialInstance m(Materials::MATERIALS[q.tile.material], q.tile.color, q.tile.temperature);
        m.moved = q.tile.moved != 0;
        m.settleCount = q.tile.settle_count;
        m.fluidAmount = q.tile.fluid_amount;
        m.fluidAmountDiff = q.tile.fluid_amount_diff;
        Particle* p = new Particle(m, q.x, q.y, q.vx, q.vy, q.ax, q.ay);
        p->targetX = q.target_x; p->targetY = q.target_y; p->targetForce = q.target_force;
        p->lifetime = q.lifetime; p->fadeTime = q.fade_time;
        p->inObjectState = q.in_object_state; p->phase = q.phase != 0; p->temporary = q.temporary != 0;
        W.particles.push_back(p);
    }
    return 0;
}

} /* extern "C" */
