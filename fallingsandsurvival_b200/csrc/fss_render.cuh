// fss_render.cuh -- the `dirty` job of Game::tick (Game.cpp:2579-2650) on the device: dirty cells -> four RGBA layers,
// the flow map's smoothing state, the per-material count of dirty cells; and World::dirty housekeeping.
//
// K4 k_render_dirty: one thread per stored cell, in storage order (the dirty byte plane is read as full lines; everything
// else is only touched where a cell is dirty).  The pixel layers are row-major like the reference's byte arrays (they are
// copied out as they are); the smoothing state prevFlowX / prevFlowY is private to this kernel and kept in storage order.
#pragma once
#include "fss_device.cuh"

struct DLook { // what the renderer reads of a Material (Material.hpp:34-38) beyond what the movement kernels keep
    uint32_t emit_color;
    int32_t iterations;
    uint32_t alpha;
};

struct RenderParams {
    Grid g;
    const DLook* looks;
    int n_mat;
    int width;
    uint32_t* px[4]; // main, fire, emission, flow: row-major width * stored rows, bytes r g b a
    float *prev_fx, *prev_fy;
    unsigned long long* moving; // n_mat counters
    uint32_t* info;             // hadDirty, hadFire, hadFlow
    size_t cells;               // stored cells (rows * pitch)
    int own_lo, own_hi;         // global rows this world renders: a strip's ghost rows belong to its neighbours
};

__device__ __forceinline__ uint32_t rgba_bytes(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return r | (g << 8) | (b << 16) | (a << 24); }

__global__ void __launch_bounds__(256) k_render_dirty(const __grid_constant__ RenderParams p) {
    __shared__ unsigned int hist[FSS_MAX_MATERIALS];
    for(int m = threadIdx.x; m < FSS_MAX_MATERIALS; m += 256) hist[m] = 0u;
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    bool fire = false, flow = false, any = false;
    uint32_t my_mat = 0xFFFFFFFFu; // material of this thread's dirty cell
    if(i < p.cells && p.g.dirty[i]) {
        const uint32_t row = (uint32_t)(i / p.g.pitch), x = fss_unperm((uint32_t)(i % p.g.pitch));
        if((int)x < p.width && (int)row + p.g.sy0 >= p.own_lo && (int)row + p.g.sy0 < p.own_hi) {
            any = true;
            p.g.dirty[i] = 0; // Game.cpp:2751: the memset after the job, for the cells that were set
            const uint32_t mf = p.g.mf[i];
            const uint32_t mat = mf_mat(mf);
            my_mat = mat;
            const size_t o = (size_t)row * p.width + x; // the layers hold the stored rows
            if(MF_IS(mf, FSS_PHYS_AIR)) { // :2590-2607
                p.px[0][o] = 0u;
                p.px[1][o] = 0u;
                p.px[2][o] = 0u;
            } else {
                const uint32_t color = p.g.col[i];
                const DLook lk = p.looks[mat];
                const uint32_t rgba = rgba_bytes((color >> 16) & 0xffu, (color >> 8) & 0xffu, color & 0xffu, lk.alpha & 0xffu);
                p.px[0][o] = rgba; // :2611-2614
                p.px[2][o] = rgba_bytes((lk.emit_color >> 16) & 0xffu, (lk.emit_color >> 8) & 0xffu, lk.emit_color & 0xffu, lk.emit_color >> 24); // :2616-2619
                if(mf & MF_FIRE) { // :2621-2627
                    p.px[1][o] = rgba;
                    fire = true;
                }
                if(MF_IS(mf, FSS_PHYS_SOUP)) { // :2628-2643: float state, double arithmetic, like the reference
                    const float fx = p.g.fx ? p.g.fx[i] : 0.f, fy = p.g.fy ? p.g.fy[i] : 0.f;
                    const float pfx = p.prev_fx[i], pfy = p.prev_fy[i];
                    float nfx = (float)((double)pfx + (double)(fx - pfx) * 0.25);
                    float nfy = (float)((double)pfy + (double)(fy - pfy) * 0.25);
                    if(nfy < 0) nfy = (float)((double)nfy * 0.5);
                    const double k = 3.0 / (double)lk.iterations + 0.5;
                    double gy = (double)nfy * k / 4.0 + 0.5, gx = (double)nfx * k / 4.0 + 0.5;
                    gy = gy < 0.0 ? 0.0 : gy; // std::max(v, 0.0)
                    gx = gx < 0.0 ? 0.0 : gx;
                    gy = 1.0 < gy ? 1.0 : gy; // std::min(v, 1.0)
                    gx = 1.0 < gx ? 1.0 : gx;
                    p.px[3][o] = rgba_bytes((uint32_t)(int)(gx * 255) & 0xffu, (uint32_t)(int)(gy * 255) & 0xffu, 0u, 0xffu);
                    p.prev_fx[i] = nfx;
                    p.prev_fy[i] = nfy;
                    flow = true;
                }
            }
            if(p.g.fx) { // :2606-2607, :2642-2647: consumed
                p.g.fx[i] = 0.f;
                p.g.fy[i] = 0.f;
            }
        }
    }
    { // :2589 movingTiles[mat->id]++: the lanes of a warp that hold the same material add once
        const unsigned peers = __match_any_sync(0xffffffffu, my_mat);
        if(my_mat != 0xFFFFFFFFu && (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[my_mat], (unsigned)__popc(peers));
    }
    if(__syncthreads_or(any)) {
        if(threadIdx.x == 0) p.info[0] = 1u;
        for(int m = threadIdx.x; m < p.n_mat; m += 256)
            if(hist[m]) atomicAdd(&p.moving[m], (unsigned long long)hist[m]);
    }
    if(fire) p.info[1] = 1u;
    if(flow) p.info[2] = 1u;
}

// The movement kernel skips the visits of chunks it has verified settled.  Such a visit is the identity on the cells but not
// on World::dirty: scan 2 marks every liquid cell it reaches (world.cpp:1823), i.e. every liquid cell of the chunk that is
// still active at that iter (the gate of world.cpp:1163-1166); nothing else in a settled chunk gets a mark.  The kernel notes
// per chunk the lowest iter it skipped (SubstepParams::qskip); this turns the notes into the marks they stand for.
struct SkippedMarks {
    Grid g;
    const uint8_t* qskip;
    int q_w, q_cj0, zx, zy, zw, zh;
    size_t cells;
};
__global__ void __launch_bounds__(256) k_skipped_marks(const __grid_constant__ SkippedMarks p) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if(i >= p.cells) return;
    const int gx = (int)fss_unperm((uint32_t)(i % p.g.pitch)) - p.zx, gy = (int)(i / p.g.pitch) + p.g.sy0 - p.zy;
    if(gx < 0 || gx >= p.zw || gy < 0 || gy >= p.zh) return;
    const uint8_t k = p.qskip[(size_t)(gy / FSS_CHUNK_H - p.q_cj0 + 1) * p.q_w + gx / FSS_CHUNK_W + 1];
    if(k == 0xFF) return;
    const uint32_t mf = p.g.mf[i];
    if(MF_IS(mf, FSS_PHYS_SOUP) && mf_iters(mf) > (int)k) p.g.dirty[i] = 1;
}

// dirty[...] = true over a rectangle (host code that edits tiles[] sets it itself: World::setTile world.cpp:1061, Game.cpp:1148 ...)
__global__ void __launch_bounds__(256) k_mark_dirty(Grid g, int rx, int ry, int rw, int rh) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if(k >= rw * rh) return;
    g.dirty[g.idx(rx + k % rw, ry + k / rw)] = 1;
}

// the same for a list of rectangles (already clipped to the world) in one launch: blockIdx.y = rectangle
__global__ void __launch_bounds__(256) k_mark_dirty_list(Grid g, const int4* __restrict__ rects) {
    const int4 r = rects[blockIdx.y]; // x, y, w, h
    const long long n = (long long)r.z * r.w;
    for(long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256)
        g.dirty[g.idx(r.x + (int)(k % r.z), r.y + (int)(k / r.z))] = 1;
}

// one byte plane of a rectangle, permuted storage -> row-major staging
__global__ void __launch_bounds__(256) k_download_bytes(const uint8_t* __restrict__ plane, uint8_t* __restrict__ out, uint32_t pitch, int sy0,
                                                        int rx, int ry, int rw, int rh) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if(k >= rw * rh) return;
    const int x = rx + k % rw, y = ry + k / rw;
    out[k] = plane[(size_t)(y - sy0) * pitch + fss_perm((uint32_t)x)];
}
