// Game-side code written against fss::World (include/fss_world.hpp), the C++ mirror of the reference's World for the tick
// path: materials as Materials.cpp defines them, cells painted through setTile / tiles[], a brush stroke between ticks,
// World::tick() every tick and World::tickTemperature() every 4th (Game.cpp:2428, :2759).
// Writes the initial grid, the edits and the final tiles[] as fss_cell arrays so that tests/test_gpu_parity.py can replay
// the same script on the oracle and compare bit for bit.
#include <cstdio>
#include <string>
#include <vector>

#include "fss_rng.h"
#include "fss_world.hpp"

using namespace fss;

static void dump(const char* path, const World& w) {
    std::vector<fss_cell> out((size_t)w.width * w.height);
    for(size_t i = 0; i < out.size(); i++) out[i] = World::toCell(w.tiles[i]);
    FILE* f = std::fopen(path, "wb");
    std::fwrite(out.data(), sizeof(fss_cell), out.size(), f);
    std::fclose(f);
}

int main(int argc, char** argv) {
    if(argc < 2) return 2;
    const std::string dir = argv[1];
    const std::string mode = argc > 2 ? argv[2] : "";
    const bool deviceRender = mode == "render";                 // World::dirty + the dirty -> pixels job of Game::tick on the GPU as well
    const bool deviceParticles = mode == "particles" || deviceRender; // World::tickParticles on the GPU as well
    const int W = 328, H = 240, ticks = 12;
    try {
        // Materials.cpp:5-46 in miniature; ids are dense
        std::vector<Material> mats;
        mats.emplace_back(0, "_AIR", PhysicsType::AIR, 0, 0.0f, 0);
        mats.back().conductionSelf = mats.back().conductionOther = 0.8f;
        mats.emplace_back(1, "Cobblestone", PhysicsType::SOLID, 0, 1.0f, 0, 0x808080);
        mats.back().conductionSelf = 0.01f;
        mats.back().conductionOther = 0.4f;
        mats.emplace_back(2, "Test Sand", PhysicsType::SAND, 20, 10.0f, 2, 0xDC9B64);
        mats.emplace_back(3, "Water", PhysicsType::SOUP, 0, 1.5f, 6, 0x00B69F);
        mats.back().createTemperature = -1023;
        mats.emplace_back(4, "Steam", PhysicsType::GAS, 0, -1.0f, 1, 0x666666);
        mats.back().isSteam = true;
        mats.emplace_back(5, "Lava", PhysicsType::SOUP, 0, 2.0f, 1, 0xFF7C00);
        mats.back().addTemp = 2;
        mats.back().conductionSelf = 0.5f;
        mats.back().conductionOther = 0.7f;
        mats.back().createTemperature = 1024;

        if(deviceRender) { // Materials.cpp: water is translucent, lava glows
            mats[3].alpha = 0x80;
            mats[5].alpha = 0xC0;
            mats[5].emitColor = 0xC0FF7C00u;
        }
        World world;
        world.deviceRender = deviceRender;
        world.init(W, H, mats, 0xF55, 0, /*condense into*/ 3);
        for(int y = 0; y < H; y++) {
            for(int x = 0; x < W; x++) {
                const bool border = x < 16 || y < 16 || x >= W - 16 || y >= H - 16;
                const uint32_t r = fss_rng_cell_key(99u, (uint32_t)x, (uint32_t)y) % 100u;
                const Material* m = border ? &mats[1] : r < 22 ? &mats[2] : r < 40 ? &mats[3] : r < 50 ? &mats[1] : r < 55 ? &mats[4] : r < 58 ? &mats[5] : &mats[0];
                world.setTile(x, y, MaterialInstance(m, m->color + (r << 8), m->createTemperature));
            }
        }
        dump((dir + "/initial.bin").c_str(), world);
        world.tickZone = Rect{16, 16, W - 32, H - 48};
        world.deviceParticles = deviceParticles;
        for(int t = 0; t < ticks; t++) {
            if(t == 5) { // a brush stroke of sand written straight into tiles[], like Game.cpp:1147
                for(int y = 40; y < 52; y++)
                    for(int x = 100; x < 140; x++) world.tiles[x + (size_t)y * W] = MaterialInstance(&mats[2], 0xABCDEF);
                world.markDirty(100, 40, 40, 12);
            }
            world.tick();
            if(deviceParticles) {
                if(t == 3) { // game code throwing a grain of sand, like the tools in Game.cpp do with particles.push_back
                    fss_particle p = {};
                    p.x = 150.5f, p.y = 60.25f, p.vx = 1.5f, p.vy = -2.0f, p.ay = 0.1f, p.fade_time = 60;
                    p.tile = World::toCell(MaterialInstance(&mats[2], 0x112233));
                    world.addParticle(p);
                }
                world.tickParticles(); // Game.cpp:2459-2483
            }
            if(t % 4 == 2) world.tickTemperature();
            if(deviceRender && t % 3 == 2) world.renderDirty(0, 0, W, H); // Game.cpp:2579-2650
        }
        if(deviceRender) {
            FILE* f = std::fopen((dir + "/layers.bin").c_str(), "wb");
            std::fwrite(world.pixels_ar.data(), 1, world.pixels_ar.size(), f);
            std::fwrite(world.pixelsFire_ar.data(), 1, world.pixelsFire_ar.size(), f);
            std::fwrite(world.pixelsEmission_ar.data(), 1, world.pixelsEmission_ar.size(), f);
            std::fwrite(world.pixelsFlow_ar.data(), 1, world.pixelsFlow_ar.size(), f);
            std::fclose(f);
            std::printf("moving");
            for(std::uint64_t v : world.movingTiles) std::printf(" %llu", (unsigned long long)v);
            std::printf(" had %d %d %d\n", (int)world.hadDirty, (int)world.hadFire, (int)world.hadFlow);
        }
        if(deviceParticles) {
            FILE* f = std::fopen((dir + "/particles.bin").c_str(), "wb");
            std::fwrite(world.particles.data(), sizeof(fss_particle), world.particles.size(), f);
            std::fclose(f);
        }
        dump((dir + "/final.bin").c_str(), world);
        size_t ndirty = 0;
        for(size_t i = 0; i < (size_t)W * H; i++) ndirty += world.dirty[i];
        std::printf("digest %016llx dirty %zu spawned_last_tick %zu tickCt %d\n", (unsigned long long)world.deviceDigest(), ndirty, world.spawned.size(), world.tickCt);
    } catch(const Error& e) {
        std::fprintf(stderr, "fss::Error %d: %s\n", e.code, e.what());
        return e.code == FSS_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
