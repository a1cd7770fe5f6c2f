// fss_world.hpp -- C++ host-side mirror of the reference's World / Material / MaterialInstance surface for the
// tick path, implemented on the C ABI of fss.h (header-only, C++14 like the reference).
//
// The reference's game code talks to the simulation through `World`'s public members
// (FallingSandSurvival/world.hpp:98-146): the raw `tiles[]` / `dirty[]` arrays, `width`, `height`, `tickZone`,
// `tickCt`, `getTile` / `setTile` (world.cpp:1053-1062), `tick()` (world.cpp:1091) and `tickTemperature()` (world.cpp:2043).
// fss::World offers the same members with the same meaning; `tiles[]` is the host mirror, the grid itself lives on the GPU:
//
//     fss::World world;
//     world.init(width, height, materials);          // World::init: allocates tiles[], dirty[] (+ the device planes)
//     world.setTile(x, y, fss::MaterialInstance(&sand, color));   // or write world.tiles[] directly and call markDirty
//     world.tickZone = {128, 128, width - 256, height - 256};     // Game.cpp:2216
//     world.tick();                                   // uploads what the host changed, runs the 24 sub-steps, brings the
//                                                     // zone back into tiles[] and sets dirty[] like the CPU tick does
//     if(tickTime % 4 == 2) world.tickTemperature();  // Game.cpp:2759
//
// Differences a maintainer has to know: rand() inside the tick is the keyed RNG of fss_rng.h;
// MaterialInstance::id is re-issued on the host; particles arrive as spawn records (world.spawned) instead of
// `new Particle`; errors throw fss::Error (the reference's tick cannot fail).
#ifndef FSS_WORLD_HPP
#define FSS_WORLD_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "fss.h"

namespace fss {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what + ": " + fss_last_error()), code(c) {}
};
inline void check(int rc, const char* what) {
    if(rc != FSS_OK) throw Error(rc, what);
}

namespace PhysicsType { // PhysicsType.hpp:6-12
enum { AIR = FSS_PHYS_AIR, SOLID = FSS_PHYS_SOLID, SAND = FSS_PHYS_SAND, SOUP = FSS_PHYS_SOUP, GAS = FSS_PHYS_GAS, PASSABLE = FSS_PHYS_PASSABLE, OBJECT = FSS_PHYS_PASSABLE };
}

struct MaterialInteraction { // Material.hpp:21-27 (only the temperature reactions are live on the tick path)
    int type = 0;
    int data1 = 0;
    int data2 = 0;
};

// One row of Materials::MATERIALS (Material.hpp:29-59) plus how Tiles::create(mat, x, y) colours a new cell of it
// (Tiles.cpp:190-268), which the tick needs when fire spreads, steam condenses or a reaction fires.
struct Material {
    std::string name;
    int id = 0;
    int physicsType = PhysicsType::AIR;
    std::uint8_t alpha = 255;
    std::uint32_t emitColor = 0; // Material.hpp:38
    float density = 0;
    int iterations = 0;
    std::uint32_t color = 0;
    std::uint32_t addTemp = 0;
    float conductionSelf = 1.0f;
    float conductionOther = 1.0f;
    bool react = false;
    std::vector<MaterialInteraction> reactions;
    int slipperyness = 1;
    bool isFire = false, isSteam = false; // `mat->id == Materials::FIRE.id` / `STEAM.id` (world.cpp:1171, :1965)
    // Tiles::create recipe
    int createKind = FSS_CREATE_FLAT;
    std::uint32_t createBase = 0, createRandMod = 0, createRandShift = 0, createRandSite = 0;
    int createTexture = -1;
    std::int32_t createTemperature = 0;

    Material() {}
    Material(int id_, std::string name_, int type, int slip, float dens, int iters, std::uint32_t col = 0)
        : name(std::move(name_)), id(id_), physicsType(type), density(dens), iterations(iters), color(col), slipperyness(slip), createBase(col) {}
};

// MaterialInstance.hpp:12-29
struct MaterialInstance {
    const Material* mat = nullptr;
    std::uint32_t color = 0;
    std::int32_t temperature = 0;
    std::uint32_t id = 0;
    bool moved = false;
    float fluidAmount = 2.0f;
    float fluidAmountDiff = 0.0f;
    std::uint8_t settleCount = 0;
    MaterialInstance() {}
    MaterialInstance(const Material* m, std::uint32_t col, std::int32_t temp = 0) : mat(m), color(col), temperature(temp) {}
};

struct Rect { // SDL_Rect
    int x = 0, y = 0, w = 0, h = 0;
};

class World {
public:
    MaterialInstance* tiles = nullptr; // host mirror, row-major width*height (world.hpp:98)
    bool* dirty = nullptr;             // set where tick() changed material or colour (world.hpp:139, consumed by Game.cpp:2584-2650)
    int width = 0, height = 0;
    int tickCt = 0;                    // world.cpp:2017
    Rect tickZone;                     // world.hpp:132
    std::vector<fss_particle_spawn> spawned; // cells that left the grid during the last tick (world.cpp:1180, :1291, :1366)
    bool emitParticles = true;
    // World::particles (world.hpp) kept on the device: tick() then leaves the spawn records where they are and
    // tickParticles() turns them into particles and runs World::tickParticles (world.cpp:2171-2311) on the GPU;
    // `particles` is the host copy of the list for World::renderParticles (world.cpp:2145-2169).
    bool deviceParticles = false;
    std::vector<fss_particle> particles;
    // World::dirty kept on the device exactly as the reference sets it, and the dirty -> pixels job of Game::tick
    // (Game.cpp:2579-2650) run there: set before init(); then renderDirty() replaces the job, and dirty[] stays unused.
    bool deviceRender = false;
    std::vector<unsigned char> pixels_ar, pixelsFire_ar, pixelsEmission_ar, pixelsFlow_ar; // Game.cpp:2452-2455, width * height * 4
    std::vector<std::uint64_t> movingTiles;                                                // Game.hpp:204
    bool hadDirty = false, hadFire = false, hadFlow = false;                               // Game.cpp:2443-2447

    World() {}
    World(const World&) = delete;
    World& operator=(const World&) = delete;
    ~World() {
        if(core_) fss_destroy(core_);
        delete[] tiles;
        delete[] dirty;
    }

    // World::init (world.cpp:50): the array part.  `materials` must be dense in id; materials[0] is what Tiles::NOTHING is made of.
    void init(int w, int h, const std::vector<Material>& materials, std::uint32_t seed = 0xF55, int device = 0, int condenseMaterial = -1) {
        width = w;
        height = h;
        mats_ = materials;
        tiles = new MaterialInstance[(size_t)w * h];
        dirty = new bool[(size_t)w * h]();
        for(size_t i = 0; i < (size_t)w * h; i++) tiles[i].mat = &mats_[0];
        fss_config cfg;
        std::memset(&cfg, 0, sizeof cfg);
        cfg.struct_size = sizeof cfg;
        cfg.width = (std::uint32_t)w;
        cfg.height = (std::uint32_t)h;
        cfg.seed = seed;
        cfg.device = device;
        cfg.flags = (emitParticles ? FSS_PARTICLES_EMIT : 0u) | (deviceRender ? (FSS_TRACK_DIRTY | FSS_TRACK_FLOW) : 0u);
        cfg.condense_material = condenseMaterial;
        check(fss_create(&cfg, &core_), "fss_create");
        std::vector<fss_material> tbl(mats_.size());
        for(size_t i = 0; i < mats_.size(); i++) {
            const Material& m = mats_[i];
            fss_material& d = tbl[i];
            std::memset(&d, 0, sizeof d);
            d.physics_type = m.physicsType;
            d.iterations = m.iterations;
            d.slipperyness = m.slipperyness;
            d.density = m.density;
            d.add_temp = m.addTemp;
            d.conduction_self = m.conductionSelf;
            d.conduction_other = m.conductionOther;
            d.color = m.color;
            d.flags = (m.isFire ? FSS_MAT_IS_FIRE : 0u) | (m.isSteam ? FSS_MAT_IS_STEAM : 0u);
            if(m.react)
                for(const MaterialInteraction& r : m.reactions)
                    if(d.n_reactions < FSS_MAX_REACTIONS) d.reactions[d.n_reactions++] = fss_reaction{r.type, r.data1, r.data2};
            d.create_kind = m.createKind;
            d.create_base = m.createBase;
            d.create_rand_mod = m.createRandMod;
            d.create_rand_shift = m.createRandShift;
            d.create_rand_site = m.createRandSite;
            d.create_texture = m.createTexture;
            d.create_temperature = m.createTemperature;
        }
        check(fss_set_materials(core_, tbl.data(), (int)tbl.size()), "fss_set_materials");
        if(deviceRender) {
            std::vector<std::uint8_t> alpha(mats_.size());
            std::vector<std::uint32_t> emit(mats_.size());
            for(size_t i = 0; i < mats_.size(); i++) alpha[i] = mats_[i].alpha, emit[i] = mats_[i].emitColor;
            check(fss_set_material_looks(core_, alpha.data(), emit.data(), (int)mats_.size()), "fss_set_material_looks");
        }
        markDirty(0, 0, w, h);
    }

    void setTexture(int index, const std::uint32_t* argb, int w, int h) { check(fss_set_texture(core_, index, argb, w, h), "fss_set_texture"); }

    // World::getTile / setTile (world.cpp:1053-1062): out-of-range reads return a SOLID tile, out-of-range writes are ignored
    MaterialInstance getTile(int x, int y) const {
        if(x < 0 || x >= width || y < 0 || y >= height) {
            for(const Material& m : mats_)
                if(m.physicsType == PhysicsType::SOLID) return MaterialInstance(&m, m.color);
            return MaterialInstance(&mats_[0], 0);
        }
        return tiles[x + (size_t)y * width];
    }
    void setTile(int x, int y, const MaterialInstance& t) {
        if(x < 0 || x >= width || y < 0 || y >= height) return;
        tiles[x + (size_t)y * width] = t;
        dirty[x + (size_t)y * width] = true;
        markDirty(x, y, 1, 1);
    }
    // after writing tiles[] directly (brush, rigid-body stamping, chunk merge ...): the rectangle goes to the GPU before the next tick
    void markDirty(int x, int y, int w, int h) {
        if(deviceRender) addMark(x, y, w, h); // World::setTile and the game's direct writes set dirty[] themselves
        if(!pending_) {
            px0_ = x, py0_ = y, px1_ = x + w, py1_ = y + h;
            pending_ = true;
        } else {
            px0_ = x < px0_ ? x : px0_;
            py0_ = y < py0_ ? y : py0_;
            px1_ = x + w > px1_ ? x + w : px1_;
            py1_ = y + h > py1_ ? y + h : py1_;
        }
    }

    // World::tick (world.cpp:1091-2041) minus the Box2D physicsCheck epilogue, which stays with the caller
    void tick() {
        flush();
        applyZone();
        check(fss_tick(core_, (std::uint32_t)tickCt), "fss_tick");
        tickCt++;
        // the tick writes at most 2 cells outside the zone (fire ignition, world.cpp:1198-1206)
        pull(tickZone.x - 2, tickZone.y - 2, tickZone.w + 4, tickZone.h + 4, true);
        if(deviceParticles) {
            spawned.clear();
            return;
        }
        // every record the tick produced: the cells they stand for have left the grid, so none may be dropped.  A query
        // (cap 0) tells how many wait; a spawn site that found the device buffer full left its cell in the grid instead
        // (fss_stats.spawns_refused), so nothing is lost there either.
        size_t n = 0, total = 0;
        check(fss_drain_particles(core_, nullptr, 0, &n, &total), "fss_drain_particles");
        spawned.resize(total);
        if(total) check(fss_drain_particles(core_, spawned.data(), spawned.size(), &n, &total), "fss_drain_particles");
        spawned.resize(n);
    }

    // World::tickParticles (world.cpp:2171-2311); needs deviceParticles.  Landed particles change tiles[] (dirty is set).
    void tickParticles() {
        flush();
        applyZone();
        check(fss_tick_particles(core_), "fss_tick_particles");
        size_t n = 0;
        check(fss_particles_count(core_, &n), "fss_particles_count");
        particles.resize(n);
        check(fss_particles_download(core_, particles.data(), n, &n), "fss_particles_download");
        // a particle that starts inside the zone lands within its path and the 32 x 32 landing search around it
        pull(tickZone.x - 48, tickZone.y - 48, tickZone.w + 96, tickZone.h + 96, true);
    }
    // particles.push_back(new Particle(...)) by game code
    void addParticle(const fss_particle& p) { check(fss_particles_add(core_, &p, 1), "fss_particles_add"); }

    // The `dirty` job of Game::tick (Game.cpp:2579-2650) and the memset after it (:2751) on the device; the window
    // (x, y, w, h) of the four layers is then copied into pixels*_ar (the rest of those arrays keeps its last content).
    void renderDirty(int x, int y, int w, int h) {
        flush();
        fss_render_info info;
        movingTiles.assign(mats_.size(), 0);
        check(fss_render_dirty(core_, movingTiles.data(), (int)mats_.size(), &info), "fss_render_dirty");
        hadDirty = info.had_dirty != 0;
        hadFire = info.had_fire != 0;
        hadFlow = info.had_flow != 0;
        if(x < 0) w += x, x = 0;
        if(y < 0) h += y, y = 0;
        if(x + w > width) w = width - x;
        if(y + h > height) h = height - y;
        if(w <= 0 || h <= 0) return;
        std::vector<unsigned char>* layers[4] = {&pixels_ar, &pixelsFire_ar, &pixelsEmission_ar, &pixelsFlow_ar};
        for(int k = 0; k < 4; k++) {
            layers[k]->resize((size_t)width * height * 4);
            check(fss_download_pixels(core_, k, x, y, w, h, layers[k]->data() + ((size_t)y * width + x) * 4, (size_t)width * 4), "fss_download_pixels");
        }
    }

    // World::tickTemperature (world.cpp:2043-2143)
    void tickTemperature() {
        flush();
        applyZone();
        check(fss_tick_temperature(core_), "fss_tick_temperature");
        pull(tickZone.x, tickZone.y, tickZone.w, tickZone.h, false);
    }

    // World::tickChunks' grid shift (world.cpp:2617-2630) on both copies
    void shift(int changeX, int changeY) {
        flush();
        check(fss_shift_grid(core_, changeX, changeY), "fss_shift_grid");
        pull(0, 0, width, height, false);
    }

    std::uint64_t deviceDigest() {
        flush();
        std::uint64_t d = 0;
        check(fss_hash(core_, &d), "fss_hash");
        return d;
    }
    fss_world* core() { return core_; }
    static fss_cell toCell(const MaterialInstance& t) {
        fss_cell c;
        c.material = (std::uint16_t)t.mat->id;
        c.moved = t.moved ? 1 : 0;
        c.settle_count = t.settleCount;
        c.color = t.color;
        c.temperature = t.temperature;
        c.fluid_amount = t.fluidAmount;
        c.fluid_amount_diff = t.fluidAmountDiff;
        return c;
    }

private:
    fss_world* core_ = nullptr;
    std::vector<Material> mats_;
    std::vector<fss_cell> stage_;
    std::vector<Rect> marks_;
    bool pending_ = false;
    int px0_ = 0, py0_ = 0, px1_ = 0, py1_ = 0;
    Rect zoneSet_;
    std::uint32_t nextId_ = 1;

    // World::dirty marks of host edits, kept exact (no cell is marked that the game did not mark) and short: a new rectangle
    // that an existing one contains is dropped, one that extends an existing one to a larger rectangle (same rows and adjacent
    // or overlapping columns, or the other way round: a brush stroke, a stamped scan line) is merged into it
    void addMark(int x, int y, int w, int h) {
        if(w <= 0 || h <= 0) return;
        for(size_t k = marks_.size(); k-- > 0 && marks_.size() - k <= 64;) { // the recent ones: edits come in runs
            Rect& m = marks_[k];
            if(x >= m.x && y >= m.y && x + w <= m.x + m.w && y + h <= m.y + m.h) return;
            if(y == m.y && h == m.h && x <= m.x + m.w && x + w >= m.x) {
                const int x1 = std::max(m.x + m.w, x + w);
                m.x = std::min(m.x, x);
                m.w = x1 - m.x;
                return;
            }
            if(x == m.x && w == m.w && y <= m.y + m.h && y + h >= m.y) {
                const int y1 = std::max(m.y + m.h, y + h);
                m.y = std::min(m.y, y);
                m.h = y1 - m.y;
                return;
            }
        }
        marks_.push_back(Rect{x, y, w, h});
    }

    void applyZone() {
        if(zoneSet_.x != tickZone.x || zoneSet_.y != tickZone.y || zoneSet_.w != tickZone.w || zoneSet_.h != tickZone.h) {
            check(fss_set_tick_zone(core_, tickZone.x, tickZone.y, tickZone.w, tickZone.h), "fss_set_tick_zone");
            zoneSet_ = tickZone;
        }
    }
    void flush() { // host edits -> device
        if(!pending_) return;
        const int x = px0_ < 0 ? 0 : px0_, y = py0_ < 0 ? 0 : py0_;
        const int w = (px1_ > width ? width : px1_) - x, h = (py1_ > height ? height : py1_) - y;
        pending_ = false;
        if(w <= 0 || h <= 0) return;
        stage_.resize((size_t)w * h);
        for(int r = 0; r < h; r++)
            for(int c = 0; c < w; c++) stage_[(size_t)r * w + c] = toCell(tiles[(x + c) + (size_t)(y + r) * width]);
        check(fss_upload_rect(core_, x, y, w, h, stage_.data(), (size_t)w), "fss_upload_rect");
        if(!marks_.empty()) { // all marks of the frame in one kernel launch
            static_assert(sizeof(Rect) == sizeof(fss_rect), "Rect and fss_rect are both four ints");
            check(fss_mark_dirty_list(core_, reinterpret_cast<const fss_rect*>(marks_.data()), marks_.size()), "fss_mark_dirty_list");
            marks_.clear();
        }
    }
    void pull(int x, int y, int w, int h, bool setDirty) { // device -> tiles[]
        if(x < 0) w += x, x = 0;
        if(y < 0) h += y, y = 0;
        if(x + w > width) w = width - x;
        if(y + h > height) h = height - y;
        if(w <= 0 || h <= 0) return;
        stage_.resize((size_t)w * h);
        check(fss_download_rect(core_, x, y, w, h, stage_.data(), (size_t)w), "fss_download_rect");
        for(int r = 0; r < h; r++) {
            for(int c = 0; c < w; c++) {
                const fss_cell& s = stage_[(size_t)r * w + c];
                const size_t i = (x + c) + (size_t)(y + r) * width;
                MaterialInstance& t = tiles[i];
                if(t.mat->id != (int)s.material || t.color != s.color) {
                    if(setDirty) dirty[i] = true;
                    if(t.mat->id != (int)s.material) t.id = nextId_++; // MaterialInstance::_curID++ (MaterialInstance.cpp:7)
                }
                t.mat = &mats_[s.material];
                t.color = s.color;
                t.temperature = s.temperature;
                t.moved = s.moved != 0;
                t.settleCount = s.settle_count;
                t.fluidAmount = s.fluid_amount;
                t.fluidAmountDiff = s.fluid_amount_diff;
            }
        }
    }
};

} // namespace fss
#endif
