// rptr_reorder.cuh -- ray reordering between the shade and the trace stage of the wavefront.
//
// The reference's megakernel traces every ray where its pixel is; a wavefront is free to choose the order in which a queue
// is traced, and the trace kernel is bound by L1 data-pipe wavefronts: lanes of a warp fetch unrelated BVH nodes, so every
// 32-byte sector costs a wavefront of its own (profiles/r02_full_trace_*).  Bounce and shadow rays leave the shade stage in
// screen order, i.e. with unrelated origins; this pass bins them by a 12-bit key of (origin, direction) so that the rays a
// warp fetches together start in the same cell of the scene (bounce rays) or run down the same beam (sun shadow rays).
// Results do not depend on the order of a queue (every ray writes its own path slot; one shadow ray per path and bounce), so
// images stay bit-identical with the option on or off -- tests/test_gpu_parity.py holds both against the oracle.
//
// One counting pass: k_bin_count (histogram of the keys, per-CTA in shared memory) -> k_bin_scan (exclusive scan, one CTA)
// -> k_bin_scatter (tiles of RPTR_BIN_TILE entries: local ranks from shared-memory atomics, one global atomic per tile and
// non-empty bin claims the output range, entries go to cursor + rank).  Not stable -- nothing needs it to be.
#pragma once
#include "rptr_shading.cuh"

namespace rp {

#define RPTR_BIN_BITS 12
#define RPTR_BINS (1 << RPTR_BIN_BITS)
#define RPTR_BIN_THREADS 512
#define RPTR_BIN_PER_THREAD 8
#define RPTR_BIN_TILE (RPTR_BIN_THREADS * RPTR_BIN_PER_THREAD)

// key modes (option "reorder_bounce" / "reorder_shadow")
#define RPTR_KEY_NONE 0
#define RPTR_KEY_ORIGIN 1        // 4 + 4 + 4 bit Morton code of the origin
#define RPTR_KEY_OCTANT_ORIGIN 2 // direction octant (major), 3 + 3 + 3 bit Morton code of the origin
#define RPTR_KEY_BEAM 3          // 6 + 6 bit Morton code of the origin projected along a fixed direction (the sun): parallel rays of one beam
#define RPTR_KEY_ORIGIN_OCTANT 4 // 3 + 3 + 3 bit Morton code of the origin (major), direction octant

struct ReorderKeys {
    uint16_t *keys_b; // key of entry i of the next bounce queue (null: off)
    uint16_t *keys_s; // key of shadow ray i (null: off)
    int32_t mode_b, mode_s;
    float center[3], inv_half; // scene box -> [-1, 1]^3
    float bu[3], bv[3];        // orthonormal pair perpendicular to the beam direction
};

RPTR_HD uint32_t spread3(uint32_t x) { // 4 bits -> every third bit
    x &= 0xfu;
    x = (x | (x << 4)) & 0x0c3u;
    x = (x | (x << 2)) & 0x249u;
    return x;
}
RPTR_HD uint32_t spread2(uint32_t x) { // 6 bits -> every second bit
    x &= 0x3fu;
    x = (x | (x << 4)) & 0x30fu;
    x = (x | (x << 2)) & 0x333u;
    x = (x | (x << 1)) & 0x555u;
    return x;
}
RPTR_HD uint32_t quant_unit(float s, uint32_t cells) { // s in [-1, 1] -> [0, cells)
    const float f = (s * 0.5f + 0.5f) * (float)cells;
    const int32_t q = (int32_t)f;
    return (uint32_t)(q < 0 ? 0 : (q >= (int32_t)cells ? (int32_t)cells - 1 : q));
}
RPTR_HD uint32_t ray_key(const ReorderKeys &rk, int32_t mode, float3 o, float3 d) {
    const float x = (o.x - rk.center[0]) * rk.inv_half, y = (o.y - rk.center[1]) * rk.inv_half, z = (o.z - rk.center[2]) * rk.inv_half;
    if (mode == RPTR_KEY_ORIGIN) return spread3(quant_unit(x, 16u)) | (spread3(quant_unit(y, 16u)) << 1) | (spread3(quant_unit(z, 16u)) << 2);
    if (mode == RPTR_KEY_BEAM) {
        const float u = (x * rk.bu[0] + y * rk.bu[1] + z * rk.bu[2]) * 0.57735027f, v = (x * rk.bv[0] + y * rk.bv[1] + z * rk.bv[2]) * 0.57735027f;
        return spread2(quant_unit(u, 64u)) | (spread2(quant_unit(v, 64u)) << 1);
    }
    const uint32_t oct = (d.x >= 0.0f ? 1u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 4u : 0u);
    const uint32_t m = spread3(quant_unit(x, 8u)) | (spread3(quant_unit(y, 8u)) << 1) | (spread3(quant_unit(z, 8u)) << 2);
    return mode == RPTR_KEY_OCTANT_ORIGIN ? ((oct << 9) | m) : ((m << 3) | oct);
}

#if defined(__CUDACC__)

__global__ void __launch_bounds__(RPTR_BIN_THREADS) k_bin_count(const uint16_t *keys, const uint32_t *count, uint32_t *hist) {
    __shared__ uint32_t s_hist[RPTR_BINS];
    const uint32_t n = *count;
    if ((size_t)blockIdx.x * RPTR_BIN_TILE >= n) return;
    for (uint32_t b = threadIdx.x; b < RPTR_BINS; b += RPTR_BIN_THREADS) s_hist[b] = 0u;
    __syncthreads();
    for (size_t base = (size_t)blockIdx.x * RPTR_BIN_TILE; base < n; base += (size_t)gridDim.x * RPTR_BIN_TILE)
#pragma unroll
        for (int k = 0; k < RPTR_BIN_PER_THREAD; ++k) {
            const size_t i = base + (size_t)k * RPTR_BIN_THREADS + threadIdx.x;
            if (i < n) atomicAdd(&s_hist[keys[i] & (RPTR_BINS - 1)], 1u);
        }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < RPTR_BINS; b += RPTR_BIN_THREADS)
        if (s_hist[b]) atomicAdd(&hist[b], s_hist[b]);
}

// exclusive scan of the histogram in place: hist[b] becomes the output cursor of bin b
__global__ void __launch_bounds__(1024) k_bin_scan(uint32_t *hist) {
    __shared__ uint32_t s_warp[32];
    constexpr int PER = RPTR_BINS / 1024;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = hist[threadIdx.x * PER + k]; sum += v[k]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t run = inc - sum + (warp > 0 ? s_warp[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < PER; ++k) { hist[threadIdx.x * PER + k] = run; run += v[k]; }
}

// values == nullptr: the entries are their own indices (shadow rays -> a permutation of 0 .. n - 1)
__global__ void __launch_bounds__(RPTR_BIN_THREADS) k_bin_scatter(const uint16_t *keys, const uint32_t *values, const uint32_t *count,
                                                                  uint32_t *cursor, uint32_t *out) {
    __shared__ uint32_t s_hist[RPTR_BINS];
    const uint32_t n = *count;
    for (size_t base = (size_t)blockIdx.x * RPTR_BIN_TILE; base < n; base += (size_t)gridDim.x * RPTR_BIN_TILE) {
        for (uint32_t b = threadIdx.x; b < RPTR_BINS; b += RPTR_BIN_THREADS) s_hist[b] = 0u;
        __syncthreads();
        uint32_t key[RPTR_BIN_PER_THREAD], rank[RPTR_BIN_PER_THREAD];
#pragma unroll
        for (int k = 0; k < RPTR_BIN_PER_THREAD; ++k) {
            const size_t i = base + (size_t)k * RPTR_BIN_THREADS + threadIdx.x;
            key[k] = 0xffffffffu;
            if (i < n) {
                key[k] = keys[i] & (RPTR_BINS - 1);
                rank[k] = atomicAdd(&s_hist[key[k]], 1u);
            }
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < RPTR_BINS; b += RPTR_BIN_THREADS) {
            const uint32_t c = s_hist[b];
            if (c) s_hist[b] = atomicAdd(&cursor[b], c);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RPTR_BIN_PER_THREAD; ++k) {
            const size_t i = base + (size_t)k * RPTR_BIN_THREADS + threadIdx.x;
            if (key[k] != 0xffffffffu) out[s_hist[key[k]] + rank[k]] = values ? values[i] : (uint32_t)i;
        }
        __syncthreads();
    }
}

#endif // __CUDACC__

} // namespace rp
