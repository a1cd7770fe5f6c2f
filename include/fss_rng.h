/*
 * fss_rng.h -- the counter-based RNG that replaces libc rand() on the World::tick path.
 *
 * The reference calls rand() a data-dependent number of times per cell
 * (FallingSandSurvival/world.cpp:1172-1966), from six pool
 * threads that share libc's hidden state, so its output depends on thread timing.
 * Every implementation of the path in this repository (the CUDA kernels, the plain-C
 * oracle and the instrumented build of the reference's own lines) replaces each
 * `rand()` by fss_rand(cellkey, site, k):
 *
 *   tkey    = fss_rng_tick_key(seed, tickCt, iter)        one value per sub-step
 *   cellkey = fss_rng_cell_key(tkey, x, y)                one value per cell visit,
 *                                                         x,y = GLOBAL world coordinates
 *   site    = the line number of the rand() call in the reference's world.cpp
 *             (1172, 1174, ... 1966), or FSS_SITE_TILES + line in Tiles.cpp for the
 *             colour draws of Tiles::create*  (Tiles.cpp:16, :54, :60, :184)
 *   k       = how many times this site has already drawn during this cell visit
 *             (0 for almost every site; the 5x5 ignition loop at world.cpp:1202 and the
 *             particle velocity expressions draw more than once)
 *
 * The value is 31-bit non-negative like glibc's rand(); all call sites use `rand() % n`.
 * This header is the specification: C, C++ and CUDA all compile this one definition.
 */
#ifndef FSS_RNG_H
#define FSS_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define FSS_HD __host__ __device__ __forceinline__
#else
#define FSS_HD static inline
#endif

#define FSS_SITE_TILES 10000u /* site id = FSS_SITE_TILES + line number in Tiles.cpp */

FSS_HD uint32_t fss_fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

FSS_HD uint32_t fss_rng_tick_key(uint32_t seed, uint32_t tick_ct, uint32_t iter) {
    return fss_fmix32(seed ^ fss_fmix32(tick_ct * 8u + iter + 0x9E3779B9u));
}

FSS_HD uint32_t fss_rng_cell_key(uint32_t tkey, uint32_t x, uint32_t y) {
    return fss_fmix32(tkey ^ (x * 0x9E3779B1u) ^ (y * 0x85EBCA77u));
}

FSS_HD uint32_t fss_rand(uint32_t cellkey, uint32_t site, uint32_t k) {
    return fss_fmix32(cellkey + site * 0x27D4EB2Fu + k * 0x165667B1u) >> 1;
}

/* per-cell 64-bit digest used by fss_hash(): order-independent sum over cells */
FSS_HD uint64_t fss_cell_digest(uint32_t x, uint32_t y, uint32_t mat_moved_settle, uint32_t color,
                                uint32_t temperature, uint32_t amount_bits, uint32_t diff_bits) {
    uint64_t h = ((uint64_t)y << 32) | x;
    h = (h ^ mat_moved_settle) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h = (h ^ color) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    h = (h ^ temperature) * 0x94D049BB133111EBull;
    h ^= h >> 29;
    h = (h ^ amount_bits) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
    h = (h ^ diff_bits) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}

#endif /* FSS_RNG_H */
