/* Cut-down declaration of the reference's `World` (FallingSandSurvival/world.hpp:89-211):
 * only the members that world.cpp:1053-1062, :1075-1088, :1091-2041, :2043-2143 and :2171-2338 use.
 * The method BODIES are never restated here: oracle/build_ref.py extracts them from
 * /root/reference at build time.  Test infrastructure; not part of the product. */
#pragma once
#include <SDL2/SDL.h>
#include "Tiles.hpp"
#include "Particle.hpp"
#include "ctpl_stl.h"

/* world.hpp:76-87 */
#define FLUID_MaxValue 0.5f
#define FLUID_MinValue 0.0005f
#define FLUID_MaxCompression 0.1f
#define FLUID_MinFlow 0.05f
#define FLUID_MaxFlow 8.0f
#define FLUID_FlowSpeed 1.0f

#ifndef CHUNK_W
#error "build with -DCHUNK_W=.. -DCHUNK_H=.."
#endif

/* easy_profiler (ProfilerConfig.hpp:4-17) compiled out */
#define EASY_FUNCTION(...)
#define EASY_BLOCK(...)
#define EASY_END_BLOCK
#define EASY_THREAD(...)

class World {
public:
    MaterialInstance* tiles = nullptr;
    float* flowX = nullptr;
    float* flowY = nullptr;
    float* prevFlowX = nullptr;
    float* prevFlowY = nullptr;
    std::vector<Particle*> particles;
    uint16_t width = 0;
    uint16_t height = 0;
    int tickCt = 0;
    static ctpl::thread_pool* tickPool;
    static ctpl::thread_pool* tickVisitedPool;
    bool* tickVisited1 = nullptr;
    bool* tickVisited2 = nullptr;
    int32_t* newTemps = nullptr;
    bool* dirty = nullptr;
    SDL_Rect tickZone{};

    MaterialInstance getTile(int x, int y);
    void setTile(int x, int y, MaterialInstance type);
    void tick();
    void tickTemperature();
    void tickParticles();
};
