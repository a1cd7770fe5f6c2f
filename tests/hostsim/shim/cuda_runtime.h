/* tests/hostsim: a stand-in for <cuda_runtime.h> so that the movement RULES of the product (csrc/fss_rules.cuh, device
 * code) compile for the host, one "thread" at a time.  Every chunk thread runs as a warp of one lane: the warp-wide
 * primitives degenerate to their single-lane meaning, which is exactly what the rules compute when the other lanes of a
 * warp have nothing to do.  Test infrastructure only. */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__

struct uint3 {
    unsigned x, y, z;
};
struct int4 {
    int x, y, z, w;
};
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, gridDim = {1, 1, 1};

using std::max;
using std::min;

static inline unsigned __float_as_uint(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(unsigned u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline unsigned __activemask() { return 1u; }
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
static inline int __any_sync(unsigned, int pred) { return pred != 0; }
static inline int __all_sync(unsigned, int pred) { return pred != 0; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
template <class T>
static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return v; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    unsigned long long old = *p;
    *p = old + v;
    return old;
}

/* ---- what csrc/fss_particles.cuh needs on top ---- */
#define __constant__
#define __shared__ static
struct char2 {
    signed char x, y;
};
static inline void __syncthreads() {}
static inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = *p;
    if(v < old) *p = v;
    return old;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = *p;
    if(v > old) *p = v;
    return old;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
    unsigned old = *p;
    *p = old + v;
    return old;
}
