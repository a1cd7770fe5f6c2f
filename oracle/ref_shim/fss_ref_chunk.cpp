/* oracle/_ref/libfss_ref_chunk.so -- the reference's OWN Chunk::read / Chunk::write (FallingSandSurvival/Chunk.cpp:47-187,
 * :189-281), extracted at build time by oracle/build_ref.py into chunk_lines.inc, behind a cut-down `Chunk` declaration with
 * the reference's member names, linked against the system's liblz4 (the library the game links; declared here because the
 * image ships its .so without headers).  TEST INFRASTRUCTURE ONLY: tests/test_chunkio.py reads files written by these lines
 * with the product's codec and the other way round. */
#include <stdint.h>
#include <stdlib.h>
#include <sys/stat.h>

#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "SDL2/SDL.h"
#include "Materials.hpp"
#include "MaterialInstance.hpp"

extern "C" {
int LZ4_compressBound(int inputSize);
int LZ4_compress_fast(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);
int LZ4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);
}

#define EASY_FUNCTION(...)
#define EASY_BLOCK(...)
#define EASY_END_BLOCK
#define logCritical(...) (fss_ref_chunk_errors++)
static int fss_ref_chunk_errors = 0;

#define CHUNK_W 128
#define CHUNK_H 128
using namespace std;

typedef struct { /* Chunk.hpp:27-31 */
    Uint16 index;
    Uint32 color;
    int32_t temperature;
} MaterialInstanceData;

class Chunk { /* the members Chunk::read / Chunk::write use (Chunk.hpp:33-62) */
public:
    std::string fname;
    int x = 0, y = 0;
    bool hasMeta = false;
    int8_t generationPhase = 0;
    bool hasTileCache = false;
    MaterialInstance* tiles = nullptr;
    MaterialInstance* layer2 = nullptr;
    Uint32* background = nullptr;
    void read();
    void write(MaterialInstance* tiles, MaterialInstance* layer2, Uint32* background);
};

#include "chunk_lines.inc"

struct Rec { /* fss_tile_record */
    uint16_t index;
    uint32_t color;
    int32_t temperature;
};

extern "C" {

void fssrefchunk_init(unsigned seed) {
    srand(seed);
    Materials::init();
}

/* Chunk::write on MaterialInstance arrays built from the records; returns the number of logCritical calls */
int fssrefchunk_write(const char* path, int phase, const Rec* tiles, const Rec* layer2, const uint32_t* background) {
    const int n = CHUNK_W * CHUNK_H;
    std::vector<MaterialInstance> t(n), l(n);
    std::vector<Uint32> bg(background, background + n);
    for(int i = 0; i < n; i++) {
        t[i] = MaterialInstance(Materials::MATERIALS_ARRAY[tiles[i].index], tiles[i].color, tiles[i].temperature);
        l[i] = MaterialInstance(Materials::MATERIALS_ARRAY[layer2[i].index], layer2[i].color, layer2[i].temperature);
    }
    Chunk c;
    c.fname = path;
    c.generationPhase = (int8_t)phase;
    fss_ref_chunk_errors = 0;
    c.write(t.data(), l.data(), bg.data());
    return fss_ref_chunk_errors;
}

/* Chunk::read; returns the number of logCritical calls, -1 if it threw */
int fssrefchunk_read(const char* path, int* phase, Rec* tiles, Rec* layer2, uint32_t* background) {
    const int n = CHUNK_W * CHUNK_H;
    Chunk c;
    c.fname = path;
    fss_ref_chunk_errors = 0;
    try {
        c.read();
    } catch(const std::exception&) {
        return -1;
    }
    *phase = c.generationPhase;
    for(int i = 0; i < n; i++) {
        tiles[i] = Rec{(uint16_t)c.tiles[i].mat->id, c.tiles[i].color, c.tiles[i].temperature};
        layer2[i] = Rec{(uint16_t)c.layer2[i].mat->id, c.layer2[i].color, c.layer2[i].temperature};
        background[i] = c.background[i];
    }
    free(c.tiles);
    free(c.layer2);
    delete[] c.background;
    return fss_ref_chunk_errors;
}

} /* extern "C" */
