// rptr_cuda.cu -- librptr_cuda.so: the wavefront path tracer and its C ABI (include/rptr_cuda.h).
//
// One frame = for each wave of sample layers:  raygen -> [ trace -> shade -> shadow ] x max_path_depth -> resolve.
// Path state lives in HBM as 16-byte SoA records indexed by path slot (slot = layer * local_pixels + local_pixel, so
// neighbouring threads hold neighbouring pixels); queues of live path slots and of shadow rays are compacted with
// warp-aggregated atomics; every kernel is a persistent grid (a multiple of the SM count) that strides over a
// device-resident count, so a whole frame is enqueued without a host round trip.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (RPTR-FP contract, see rptr_math.cuh).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rptr_cuda.h"
#include "rptr_host.hpp"
#include "rptr_trace_kernels.cuh"
#include "rptr_reorder.cuh"
#include "rptr_trace_tail.cuh"
#include "rptr_bvh_build.hpp"
#include "rptr_post.cuh"

using namespace rp;

#ifndef RPTR_IMPLIED_INITIAL_STATE
#define RPTR_IMPLIED_INITIAL_STATE 0 // see k_raygen
#endif

// ---------------------------------------------------------------------------------------------------------------------
// device-side buffers
// ---------------------------------------------------------------------------------------------------------------------
struct Wave {
    float4 *ray_o;  // origin.xyz, tmin
    float4 *ray_d;  // dir.xyz, tmax
    float4 *hit;    // t, u, v, bits(leaf triangle index | -1)
    float4 *thr;    // throughput.rgb, prev_bounce_pdf
    float4 *illum;  // illum.rgb, total_t
    uint2 *rngb;    // sampler word that advances with the draws (LCG state), bounce
    uint32_t *rng2; // second sampler word (Sobol index / BN pixelID); allocated for rng_variant != UNIFORM only
    uint32_t *rng3; // LCG of the stochastic alpha test when the pointset is not the LCG (pt_megakernel.glsl:354-358); with rng2
    float4 *foot;   // texture_footprint (a GLSL mat2: m00, m01, m10, m11); allocated for scenes with image textures only
    float4 *sh_o;   // shadow queue: origin.xyz, tmin
    float4 *sh_d;   //               dir.xyz, tmax
    float4 *sh_c;   //               contribution.rgb, bits(path slot)
    uint32_t *queue[2];
    uint32_t *hitq;       // slots of the closest-hit rays of the current bounce that hit something (dense; input of the shade stage)
    uint32_t *hit_counts; // per bounce
    uint32_t *counts; // per bounce d: [4d] live paths entering it, [4d+1] its shadow rays, [4d+2], [4d+3] fetch cursors
    // ray reordering (rptr_reorder.cuh; allocated when option reorder_bounce / reorder_shadow is on)
    uint16_t *keys_b, *keys_s; // bin keys of the next bounce queue's entries / of the shadow rays, written by the shade stage
    uint32_t *queue_sorted;    // the bounce queue in bin order (what the trace stage reads; the shade stage keeps the screen order)
    uint32_t *sh_perm;         // shadow-ray indices in bin order
    uint32_t *bin_hist;        // per bounce d: RPTR_BINS counters / cursors of the bounce queue, then of the shadow queue
    // tail hand-over (rptr_trace_tail.cuh): records of the closest-hit launch, then of the shadow launch; per bounce d: [4d], [4d+1] record
    // counts of the two launches, [4d+2], [4d+3] the cursors of their tail kernels
    TailRec *tail;
    uint32_t *tail_counts;
};

// The fp16 AOV images of the reference (aov_albedo_roughness_buffer, aov_normal_depth_buffer: vulkan/accumulate.glsl:19-23).
// Every sample layer of a frame imageStore()s to them, last writer wins; with the sequential reading of a batch
// (batch_spp = k == k frames of 1) that is the LAST layer of the frame, whose slots are [slot_lo, slot_lo + local_pixels).
struct AovTarget {
    ushort4 *albedo_roughness; // null: AOV images off
    ushort4 *normal_depth;
    ushort4 *motion_jitter;
    uint32_t slot_lo;
};
__device__ __forceinline__ ushort4 to_half4(float x, float y, float z, float w) { // rgba16f store: round to nearest even
    return make_ushort4(__half_as_ushort(__float2half_rn(x)), __half_as_ushort(__float2half_rn(y)), __half_as_ushort(__float2half_rn(z)),
                        __half_as_ushort(__float2half_rn(w)));
}
__device__ __forceinline__ void store_aov(const AovTarget &t, const TileMap &tm, uint32_t slot, const AovSample &a) {
    const uint32_t lp = slot - t.slot_lo;
    const size_t px = (size_t)local_row_to_global(tm, (int32_t)(lp / (uint32_t)tm.width)) * tm.width + lp % (uint32_t)tm.width;
    t.albedo_roughness[px] = to_half4(a.albedo.x, a.albedo.y, a.albedo.z, a.roughness);
    t.normal_depth[px] = to_half4(a.normal.x, a.normal.y, a.normal.z, a.depth);
    t.motion_jitter[px] = to_half4(a.motion[0], a.motion[1], a.jitter[0], a.jitter[1]);
}

struct DevCounters {
    unsigned long long closest_rays, shadow_rays, shaded_vertices, closest_nodes, closest_tris, shadow_nodes, shadow_tris, samples;
};

// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_append(uint32_t *counter, bool pred) {
    // warp-aggregated atomic: one atomicAdd per warp, returns this lane's slot (valid when pred)
    const unsigned mask = __ballot_sync(__activemask(), pred);
    if (!pred) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1));
}

__device__ __forceinline__ void flush_counter(unsigned long long *dst, unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}

// raygen (vulkan/pt_megakernel.glsl:310-365): one thread per path slot of the wave.  Sample index of layer l of the wave =
// sample_base + first_layer + l (frames: sample_base = view_params.frame_id = fp.first_sample; ray queries: 0,
// accumulation_frame_offset of record_frame, vulkan/render_vulkan.cpp:2982).  With `queries` the camera ray is replaced by the
// caller's ray AFTER the sampler has been seeded and the pixel-filter draws consumed (:326-334).
__global__ void __launch_bounds__(256) k_raygen(FrameParams fp, TileMap tm, Wave w, uint32_t sample_base, int32_t first_layer, int32_t n_layers,
                                                const rptr_render_ray_query *queries) {
    const uint32_t n = (uint32_t)n_layers * (uint32_t)tm.local_pixels;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const int32_t layer = (int32_t)(slot / (uint32_t)tm.local_pixels);
        const uint32_t lp = slot % (uint32_t)tm.local_pixels;
        uint32_t upx, upy;
        tile_pixel(tm, lp, upx, upy);
        const int32_t px = (int32_t)upx, py = (int32_t)upy;
        PathState ps;
        generate_primary(fp, px, py, sample_base + (uint32_t)(first_layer + layer), ps);
        if (queries) {
            const rptr_render_ray_query q = queries[lp];
            ps.o = f3(q.origin[0], q.origin[1], q.origin[2]);
            ps.d = f3(q.dir[0], q.dir[1], q.dir[2]);
            ps.tmax = q.t_max;
        }
        // RPTR_IMPLIED_INITIAL_STATE: thr = (1, 1, 1, prev_pdf = 2e16) and illum = 0 are implied for a path that has not been
        // shaded yet (bounce counter 0): the first shade launch and the resolve kernel substitute them instead of reading 32
        // bytes raygen would have to write
        w.ray_o[slot] = f4(ps.o.x, ps.o.y, ps.o.z, ps.tmin);
        w.ray_d[slot] = f4(ps.d.x, ps.d.y, ps.d.z, ps.tmax);
#if !RPTR_IMPLIED_INITIAL_STATE
        w.thr[slot] = f4(1.0f, 1.0f, 1.0f, ps.prev_pdf);
        w.illum[slot] = f4(0.0f, 0.0f, 0.0f, 0.0f);
#endif
        w.rngb[slot] = make_uint2(ps.rng, 0u);
        if (w.foot) {
            if (queries) init_footprint(fp, ps); // the footprint block runs on the ray actually traced (pt_megakernel.glsl:326-351)
            w.foot[slot] = f4(ps.foot.m00, ps.foot.m01, ps.foot.m10, ps.foot.m11);
        }
        if (fp.rng_variant != 0) {
            w.rng2[slot] = ps.rng_b;
            if (w.rng3) w.rng3[slot] = alpha_lcg_seed(fp, px, py, sample_base + (uint32_t)(first_layer + layer));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) w.counts[0] = n;
}

// closest hit for the live paths of bounce d
__global__ void __launch_bounds__(128) k_trace(BvhDev bvh, SceneDev sc, Wave w, const uint32_t *queue, const uint32_t *count, DevCounters *dc, uint32_t *hit_count) {
    const uint32_t n = *count;
    TraceCounters cnt{0, 0};
    unsigned long long rays = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = queue ? queue[i] : i;
        const float4 o = w.ray_o[slot], d = w.ray_d[slot];
        HitRec h;
        uint32_t *ap = w.rng3 ? w.rng3 + slot : &w.rngb[slot].x;
        uint32_t st = *ap;
        const uint32_t before = st;
        closest_hit_filtered(bvh, sc, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), o.w, d.w, st, h, cnt);
        if (st != before) *ap = st;
        w.hit[slot] = f4(h.t, h.u, h.v, __int_as_float(h.tri));
        if (hit_count) {
            const uint32_t hi = warp_append(hit_count, h.tri >= 0);
            if (h.tri >= 0) w.hitq[hi] = slot;
        }
        rays++;
    }
    flush_counter(&dc->closest_rays, rays);
    flush_counter(&dc->closest_nodes, cnt.nodes);
    flush_counter(&dc->closest_tris, cnt.tris);
}

// Shade the vertices found by the trace stage; emits shadow rays and the next bounce's queue.
// Divergence control: a CTA takes tiles of RPTR_SHADE_TILE queue entries, counting-sorts them in shared memory by
// (miss | material id) and hands every warp a run of equal keys, so that sky evaluation, Lambert, GGX, transmissive and
// emissive vertices do not share warps (the reference's megakernel pays that divergence inside every workgroup).
#ifndef RPTR_SHADE_THREADS
#define RPTR_SHADE_THREADS 128
#endif
#ifndef RPTR_SHADE_PER_THREAD
#define RPTR_SHADE_PER_THREAD 4
#endif
#define RPTR_SHADE_TILE (RPTR_SHADE_THREADS * RPTR_SHADE_PER_THREAD)
#define RPTR_SHADE_KEYS 16
#ifndef RPTR_SHADE_MIN_BLOCKS
#define RPTR_SHADE_MIN_BLOCKS 4
#endif
template <int FEAT>
__global__ void __launch_bounds__(RPTR_SHADE_THREADS, RPTR_SHADE_MIN_BLOCKS) k_shade(FrameParams fp, SceneDev sc, BvhDev bvh, Wave w, const uint32_t *queue,
                                                              const uint32_t *count, uint32_t *next_queue, uint32_t *next_count,
                                                              uint32_t *shadow_count, DevCounters *dc, AovTarget aov, TileMap tm, int sort_tiles, int first_bounce,
                                                              ReorderKeys rk) {
    __shared__ uint32_t s_hist[RPTR_SHADE_KEYS], s_base[RPTR_SHADE_KEYS];
    __shared__ uint32_t s_sorted[RPTR_SHADE_TILE];
    const uint32_t n = *count;
    unsigned long long verts = 0;
    const uint32_t n_tiles = (n + RPTR_SHADE_TILE - 1) / RPTR_SHADE_TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t tile_base = tile * RPTR_SHADE_TILE;
        const uint32_t tile_count = min((uint32_t)RPTR_SHADE_TILE, n - tile_base);
        if (sort_tiles) { // (every material of the scene takes the same code path otherwise: nothing to sort, rptr_cuda_draw_frame)
        if (threadIdx.x < RPTR_SHADE_KEYS) s_hist[threadIdx.x] = 0;
        __syncthreads();
        uint32_t my_slot[RPTR_SHADE_PER_THREAD], my_key[RPTR_SHADE_PER_THREAD], my_pos[RPTR_SHADE_PER_THREAD];
        // the two dependent gathers of the key (queue word, then the hit record) are issued for all of the thread's entries before
        // any of them is consumed: two memory round trips per tile instead of eight (profiles/r02_shade_source_stalls.md)
        int my_tri[RPTR_SHADE_PER_THREAD];
#pragma unroll
        for (int k = 0; k < RPTR_SHADE_PER_THREAD; ++k) {
            const uint32_t j = k * RPTR_SHADE_THREADS + threadIdx.x;
            my_slot[k] = j < tile_count ? (queue ? queue[tile_base + j] : tile_base + j) : 0xffffffffu;
        }
#pragma unroll
        for (int k = 0; k < RPTR_SHADE_PER_THREAD; ++k) my_tri[k] = my_slot[k] != 0xffffffffu ? __float_as_int(w.hit[my_slot[k]].w) : -1;
#pragma unroll
        for (int k = 0; k < RPTR_SHADE_PER_THREAD; ++k) {
            my_key[k] = 0xffffffffu;
            if (my_slot[k] != 0xffffffffu) {
                const int tri = my_tri[k];
                uint32_t key = tri >= 0 ? 1u : 0u; // miss / hit

                if (tri >= 0 && sort_tiles > 1) { // several code paths in the scene: look the material up
                    const Tri *tr = bvh.tris + tri;
                    const GeomInst &g = sc.ginst[tri_geom_inst(*tr)];
                    // key = code path of shade_vertex, not the material itself (keeps neighbouring pixels together):
                    // Lambert / GGX / thin or thick transmission, plus the emitter-MIS variant of each
                    const rptr_base_material &m = sc.materials[calc_hit_material_id(g, (uint32_t)tr->prim)];
                    if (m.ior > 1.0f) key = 2u;
                    if ((FEAT & RPTR_FEAT_TRANSMISSION) && fp.transmission && m.ior > 1.0f && m.specular_transmission > 0.0f) key = (m.flags & RPTR_BASE_MATERIAL_ONESIDED) ? 4u : 3u;
                    if (m.emission_intensity != 0.0f) key += 5u;
                }
                my_key[k] = key;
                my_pos[k] = atomicAdd(&s_hist[key], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int b = 0; b < RPTR_SHADE_KEYS; ++b) { s_base[b] = acc; acc += s_hist[b]; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RPTR_SHADE_PER_THREAD; ++k)
            if (my_key[k] != 0xffffffffu) s_sorted[s_base[my_key[k]] + my_pos[k]] = my_slot[k];
        __syncthreads();
        }
        for (int k = 0; k < RPTR_SHADE_PER_THREAD; ++k) {
            const uint32_t j = k * RPTR_SHADE_THREADS + threadIdx.x;
            const bool active = j < tile_count;
            bool cont = false, shadow = false;
            uint32_t slot = 0, key_b = 0;
            ShadowRay sh;
            sh.tmax = -1.0f;
            if (active) {
                slot = sort_tiles ? s_sorted[j] : (queue ? queue[tile_base + j] : tile_base + j);
                const float4 hit = w.hit[slot];
                const int tri = __float_as_int(hit.w);
                if (tri >= 0) { // a miss ends the path; its sky term is added by k_resolve from the untouched path state
                    const float4 o = w.ray_o[slot], d = w.ray_d[slot];
                    const bool implied = RPTR_IMPLIED_INITIAL_STATE && first_bounce;
                    const float4 thr = implied ? f4(1.0f, 1.0f, 1.0f, 2.e16f) : w.thr[slot]; // generate_primary's initial state
                    const float4 il = implied ? f4(0.0f, 0.0f, 0.0f, 0.0f) : w.illum[slot];
                    const uint2 rb = w.rngb[slot];
                    PathState ps;
                    ps.o = f3(o.x, o.y, o.z); ps.tmin = o.w;
                    ps.d = f3(d.x, d.y, d.z); ps.tmax = d.w;
                    ps.thr = f3(thr.x, thr.y, thr.z); ps.prev_pdf = thr.w;
                    ps.illum = f3(il.x, il.y, il.z); ps.total_t = il.w;
                    ps.rng = rb.x; ps.bounce = (int)rb.y;
                    ps.rng_b = ((FEAT & RPTR_FEAT_QMC) && fp.rng_variant != 0) ? w.rng2[slot] : 0u;
                    ps.rng_dim = 0;
                    if ((FEAT & RPTR_FEAT_TEXTURES) && fp.image_textures) { const float4 ft = w.foot[slot]; ps.foot = Footprint{ft.x, ft.y, ft.z, ft.w}; }
                    verts++;
                    // first vertex of a path of the frame's last sample layer: its attributes go to the AOV images
                    const bool want_aov = aov.albedo_roughness && ps.bounce == 0 && slot - aov.slot_lo < (uint32_t)tm.local_pixels;
                    AovSample as;
                    ShadeResult r = shade_hit<FEAT>(fp, sc, ps, hit.x, hit.y, hit.z, &bvh.tris[tri], sh, want_aov ? &as : nullptr);
                    if (want_aov) store_aov(aov, tm, slot, as);
                    cont = r == SHADE_CONTINUE;
                    shadow = sh.tmax > 0.0f;
                    w.illum[slot] = f4(ps.illum.x, ps.illum.y, ps.illum.z, ps.total_t);
                    w.rngb[slot] = make_uint2(ps.rng, (uint32_t)ps.bounce);
                    if (cont) {
                        w.ray_o[slot] = f4(ps.o.x, ps.o.y, ps.o.z, ps.tmin);
                        w.ray_d[slot] = f4(ps.d.x, ps.d.y, ps.d.z, ps.tmax);
                        w.thr[slot] = f4(ps.thr.x, ps.thr.y, ps.thr.z, ps.prev_pdf);
                        if ((FEAT & RPTR_FEAT_TEXTURES) && fp.image_textures) w.foot[slot] = f4(ps.foot.m00, ps.foot.m01, ps.foot.m10, ps.foot.m11);
                        if (rk.keys_b) key_b = ray_key(rk, rk.mode_b, ps.o, ps.d);
                    }
                }
            }
            __syncwarp();
            const uint32_t qi = warp_append(next_count, cont);
            if (cont) {
                next_queue[qi] = slot;
                if (rk.keys_b) rk.keys_b[qi] = (uint16_t)key_b;
            }
            const uint32_t si = warp_append(shadow_count, shadow);
            if (shadow) {
                w.sh_o[si] = f4(sh.o.x, sh.o.y, sh.o.z, sh.tmin);
                w.sh_d[si] = f4(sh.d.x, sh.d.y, sh.d.z, sh.tmax);
                w.sh_c[si] = f4(sh.contrib.x, sh.contrib.y, sh.contrib.z, __uint_as_float(slot));
                if (rk.keys_s) rk.keys_s[si] = (uint16_t)ray_key(rk, rk.mode_s, sh.o, sh.d);
            }
        }
        if (sort_tiles) __syncthreads();
    }
    flush_counter(&dc->shaded_vertices, verts);
}

// any-hit visibility for the NEE samples of bounce d (vulkan/pt_megakernel.glsl:216-272); unoccluded -> illum += contrib
__global__ void __launch_bounds__(128) k_shadow(BvhDev bvh, Wave w, const uint32_t *count, DevCounters *dc, AlphaFilter af, TileMap tm) {
    const uint32_t n = *count;
    TraceCounters cnt{0, 0};
    unsigned long long rays = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o = w.sh_o[i], d = w.sh_d[i];
        HitRec h;
        const float4 c = w.sh_c[i];
        const uint32_t slot = __float_as_uint(c.w);
        af.pixel_linear = tile_pixel_linear(tm, slot % (uint32_t)tm.local_pixels);
        const bool occluded = trace_ray<true>(bvh, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), o.w, d.w, h, cnt, o.w, 0x7fffffff, &af);
        rays++;
        if (!occluded) {
            float4 il = w.illum[slot];
            il.x = il.x + c.x; il.y = il.y + c.y; il.z = il.z + c.z;
            w.illum[slot] = il;
        }
    }
    flush_counter(&dc->shadow_rays, rays);
    flush_counter(&dc->shadow_nodes, cnt.nodes);
    flush_counter(&dc->shadow_tris, cnt.tris);
}

// The vec4 main_spp returns for the path in `slot` (vulkan/pt_megakernel.glsl:736): paths that ended with a miss (hit record
// still says "no triangle") get their sky / sun-disc term here, from the ray direction, throughput and previous-bounce pdf
// they died with (shade_miss).  *primary_miss = the primary ray left the scene.
__device__ __forceinline__ float4 path_sample(const FrameParams &fp, const Wave &w, uint32_t slot, bool *primary_miss) {
    const float alpha = w.rngb[slot].y == 0u ? 0.0f : 1.0f;
    const bool miss = __float_as_int(w.hit[slot].w) < 0;
    const bool untouched = RPTR_IMPLIED_INITIAL_STATE && miss && alpha == 0.0f; // primary ray left the scene: never shaded
    float4 il = untouched ? f4(0.0f, 0.0f, 0.0f, 0.0f) : w.illum[slot];
    if (miss) {
        const float4 d = w.ray_d[slot];
        const float4 thr = untouched ? f4(1.0f, 1.0f, 1.0f, 2.e16f) : w.thr[slot];
        const float3 r = shade_miss(fp.sp, f3(il.x, il.y, il.z), f3(thr.x, thr.y, thr.z), f3(d.x, d.y, d.z), thr.w);
        il.x = r.x; il.y = r.y; il.z = r.z;
    }
    *primary_miss = miss && alpha == 0.0f;
    return f4(il.x, il.y, il.z, alpha);
}

// accumulate.glsl:68-73 + process_samples.comp:116-129 replayed in sample order for the layers of this wave:
// sample k (0-based since the last reset) is stored when k == 0 and folded as m += (x - m) / float(k + 1) otherwise.
__global__ void __launch_bounds__(256) k_resolve(FrameParams fp, TileMap tm, Wave w, float4 *accum, uint32_t first_sample, int32_t n_layers,
                                                 DevCounters *dc, AovTarget aov, int discard_history) {
    unsigned long long samples = 0;
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < (uint32_t)tm.local_pixels; lp += gridDim.x * blockDim.x) {
        const int32_t px = (int32_t)(lp % (uint32_t)tm.width);
        const int32_t py = local_row_to_global(tm, (int32_t)(lp / (uint32_t)tm.width));
        float4 *dst = accum + (size_t)py * tm.width + px;
        float4 m = *dst;
        for (int32_t l = 0; l < n_layers; ++l) {
            const uint32_t slot = (uint32_t)l * (uint32_t)tm.local_pixels + lp;
            bool primary_miss;
            const float4 x = path_sample(fp, w, slot, &primary_miss);
            if (primary_miss && aov.albedo_roughness && slot - aov.slot_lo < (uint32_t)tm.local_pixels) store_aov(aov, tm, slot, aov_of_miss(fp));
            const uint32_t k = first_sample + (uint32_t)l;
            if (k > 0 && !discard_history) { // process_samples.comp:116-127: REPROJECTION_MODE_DISCARD_HISTORY keeps the new sample only
                const float denom = (float)(k + 1u);
                m.x += (x.x - m.x) / denom;
                m.y += (x.y - m.y) / denom;
                m.z += (x.z - m.z) / denom;
                m.w += (x.w - m.w) / denom;
            } else
                m = x;
            samples++;
        }
        *dst = m;
    }
    flush_counter(&dc->samples, samples);
}

// accumulate_query (vulkan/accumulate.glsl:32-42) replayed in sample order for the layers of this wave (the reference runs the
// layers of a batch concurrently; one after the other is the race-free reading, as for frames).  Statement by statement:
//   accum = sample_index > 0 ? ray_results[q] : 0;  accum += (x - accum) / (sample_index + 1);
//   if (sample_index == 0) ray_results[q] = accum; else ray_results[q] += accum;
// -- for sample_index > 0 the stored value is the OLD result PLUS the updated mean, not the mean: kept as written.
__global__ void __launch_bounds__(256) k_resolve_queries(FrameParams fp, Wave w, float4 *results, uint32_t n_queries, uint32_t first_sample,
                                                         int32_t n_layers, DevCounters *dc) {
    unsigned long long samples = 0;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_queries; q += gridDim.x * blockDim.x) {
        float4 r = results[q];
        for (int32_t l = 0; l < n_layers; ++l) {
            bool primary_miss;
            const float4 x = path_sample(fp, w, (uint32_t)l * n_queries + q, &primary_miss);
            const uint32_t k = first_sample + (uint32_t)l;
            float4 a = k > 0 ? r : f4(0.0f, 0.0f, 0.0f, 0.0f);
            const float denom = (float)(k + 1u);
            a.x += (x.x - a.x) / denom; a.y += (x.y - a.y) / denom; a.z += (x.z - a.z) / denom; a.w += (x.w - a.w) / denom;
            if (k == 0) r = a;
            else { r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w; }
            samples++;
        }
        results[q] = r;
    }
    flush_counter(&dc->samples, samples);
}

// RQ_CLOSEST (vulkan/rt_intersect.comp:30-68): opaque closest hit (gl_RayFlagsOpaqueEXT: no alpha test) over
// (RAY_EPSILON * |origin|, t_max); mode < 0 leaves the result slot untouched; a miss stores (-1, -1, bits(-1), bits(-1)).
// The rays go through the persistent traversal kernel like the path tracer's own (k_trace_persistent<false, false>):
// k_rq_prepare turns the queries into its ray records, k_rq_pack its hit records into the result words.
__global__ void __launch_bounds__(256) k_rq_prepare(const rptr_render_ray_query *q, int32_t n, float4 *ray_o, float4 *ray_d, uint32_t *counts) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float3 o = f3(q[i].origin[0], q[i].origin[1], q[i].origin[2]);
        const bool skip = q[i].mode_or_data < 0;
        ray_o[i] = f4(o.x, o.y, o.z, RPTR_RAY_EPSILON * length(o));
        ray_d[i] = f4(q[i].dir[0], q[i].dir[1], q[i].dir[2], skip ? -1.0f : q[i].t_max); // empty range: a skipped query finds nothing
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { counts[0] = (uint32_t)n; counts[1] = 0u; }
}
__global__ void __launch_bounds__(256) k_rq_pack(BvhDev bvh, const rptr_render_ray_query *q, int32_t n, const float4 *hit, float4 *results, float *hit_t) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (q[i].mode_or_data < 0) continue;
        const float4 h = hit[i];
        const int32_t tri = __float_as_int(h.w);
        const bool ok = tri >= 0;
        int32_t gi = -1, prim = -1;
        if (ok) { gi = tri_geom_inst(bvh.tris[tri]); prim = bvh.tris[tri].prim; }
        results[i] = f4(ok ? h.y : -1.0f, ok ? h.z : -1.0f, __int_as_float(gi), __int_as_float(prim));
        if (hit_t) hit_t[i] = ok ? h.x : -1.0f;
    }
}
// the same service with one ray per thread (option "trace_kernel" = 1: A/B reference)
__global__ void __launch_bounds__(128) k_ray_queries(BvhDev bvh, const rptr_render_ray_query *q, int32_t n, float4 *results, float *hit_t) {
    TraceCounters cnt{0, 0};
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (q[i].mode_or_data < 0) continue;
        const float3 o = f3(q[i].origin[0], q[i].origin[1], q[i].origin[2]);
        const float tmin = RPTR_RAY_EPSILON * length(o);
        HitRec h;
        const bool ok = trace_ray<false>(bvh, o, f3(q[i].dir[0], q[i].dir[1], q[i].dir[2]), tmin, q[i].t_max, h, cnt);
        int32_t gi = -1, prim = -1;
        if (ok) { gi = tri_geom_inst(bvh.tris[h.tri]); prim = bvh.tris[h.tri].prim; }
        results[i] = f4(ok ? h.u : -1.0f, ok ? h.v : -1.0f, __int_as_float(gi), __int_as_float(prim));
        if (hit_t) hit_t[i] = ok ? h.t : -1.0f;
    }
}

// LDR framebuffer of the display path (process_samples.comp:138-200 with ENABLE_AOV_BUFFERS; not part of .pfm parity):
// exposure for the colour channel, the AOV images for output_channel 1 / 2 (output_moment selects roughness / depth),
// then linear_to_srgb (rendering/util.glsl:25-28) and the rgba8 store.  Tone-mapping operators (early_tone_mapping_mode >= 0)
// and the motion / jitter image are not produced; pixels whose alpha is negative are left as they are (:139-140).
__device__ __forceinline__ float4 half4_to_float4(ushort4 h) {
    return f4(__half2float(__ushort_as_half(h.x)), __half2float(__ushort_as_half(h.y)), __half2float(__ushort_as_half(h.z)),
              __half2float(__ushort_as_half(h.w)));
}
// tonemap() of rendering/postprocess/tonemapping_utils.glsl:9-32 (modes: postprocess/tonemapping.h:7-9)
__device__ __forceinline__ void tonemap(int mode, float4 &c) {
    if (mode == 2) { // FAST_TONE_MAPPING
        c.x = c.x / (1.0f + c.x); c.y = c.y / (1.0f + c.y); c.z = c.z / (1.0f + c.z);
    } else if (mode == 1) { // NEUTRAL_TONE_MAPPING
        const float level = fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, 1.0f));
        const float a = 0.1f * log2f(level);
        const float scale = (a * (1.0f - 0.8f) + 1.0f * 0.8f) / level; // mix(a, 1, 0.8) / level
        c.x *= scale; c.y *= scale; c.z *= scale;
    }
}
// display_alpha: with the temporal resolve the accumulator's alpha holds the history weight and the display chain goes on with the
// alpha of the frame's own sample (process_samples.comp:107-113: the return value of reproject_and_accumulate).
// upscale: render_upscale_factor; the LDR target is upscale x larger, 2 replicates every pixel 2 x 2, any other factor stores the
// pixel at its own coordinates (process_samples.comp:192-199, as written).
__global__ void __launch_bounds__(256) k_to_srgb8(const float4 *accum, const float *display_alpha, const ushort4 *aov_ar, const ushort4 *aov_nd,
                                                  const ushort4 *aov_mj, uchar4 *out, uint32_t n, float exposure_scale, int output_channel,
                                                  int output_moment, int tone_mapping_mode, float width, float height, int upscale) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 c = accum[i];
        if (display_alpha) c.w = display_alpha[i];
        c.w = fminf(c.w, 1.0f);
        if (!(c.w >= 0.0f)) continue;
        if (output_channel == 0) {
            c.x *= exposure_scale; c.y *= exposure_scale; c.z *= exposure_scale;
            if (tone_mapping_mode >= 0) tonemap(tone_mapping_mode, c);
        } else if (output_channel == 1 && aov_ar) {
            c = half4_to_float4(aov_ar[i]);
            if (output_moment != 0) c.x = c.y = c.z = c.w;
        } else if (output_channel == 2 && aov_nd) {
            c = half4_to_float4(aov_nd[i]);
            if (output_moment != 0) c.x = c.y = c.z = c.w * 0.05f;
            else { c.x = c.x * 0.5f + 0.5f; c.y = c.y * 0.5f + 0.5f; c.z = c.z * 0.5f + 0.5f; }
        } else if (output_channel == 3 && aov_mj) { // OUTPUT_CHANNEL_MOTION_JITTER (process_samples.comp:163-178)
            const float4 m = half4_to_float4(aov_mj[i]);
            if (output_moment == 0) c = f4(fabsf(10.0f * m.x), fabsf(10.0f * m.y), 0.0f, 1.0f);
            else { // undo the scaling of update_view_parameters: back to the Halton point in [0, 1)
                const float jx = (m.z + 1.0f / width) * (width / 2.0f), jy = (m.w + 1.0f / height) * (height / 2.0f);
                c = f4(jx * 0.5f + 0.5f, jy * 0.5f + 0.5f, 0.0f, 1.0f);
            }
        } else if (output_channel == 3) {
            c = f4(0.0f, 0.0f, 0.0f, 1.0f);
        }
        const float v[3] = {c.x, c.y, c.z};
        unsigned char o[3];
        for (int k = 0; k < 3; ++k) {
            const float x = v[k];
            const float s = (x <= 0.0031308f) ? 12.92f * x : 1.055f * powf(fmaxf(fabsf(x), 1.192092896e-07f), 1.0f / 2.4f) - 0.055f;
            o[k] = (unsigned char)(fminf(fmaxf(s, 0.0f), 1.0f) * 255.0f + 0.5f); // NaN (0 * inf of a far depth) stores 0
        }
        const uchar4 px = make_uchar4(o[0], o[1], o[2], (unsigned char)(fminf(fmaxf(c.w, 0.0f), 1.0f) * 255.0f + 0.5f));
        if (upscale == 1) {
            out[i] = px;
        } else {
            const uint32_t w = (uint32_t)width, x = i % w, y = i / w, ow = w * (uint32_t)upscale;
            if (upscale == 2) {
                out[(size_t)(2 * y) * ow + 2 * x] = px;
                out[(size_t)(2 * y + 1) * ow + 2 * x] = px;
                out[(size_t)(2 * y) * ow + 2 * x + 1] = px;
                out[(size_t)(2 * y + 1) * ow + 2 * x + 1] = px;
            } else
                out[(size_t)y * ow + x] = px;
        }
    }
}

// process_samples.comp:106-113 for the frames after the first one since a reset, reprojection_mode == ACCUMULATE, temporal build
__global__ void __launch_bounds__(256) k_reproject(ResolveImages im, float4 *accum, float *display_alpha, float min_sample_weight, int batch) {
    const uint32_t n = (uint32_t)im.w * (uint32_t)im.h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 stored;
        const float4 shown = reproject_and_accumulate(im, accum[i], (int)(i % (uint32_t)im.w), (int)(i / (uint32_t)im.w), min_sample_weight, batch, &stored);
        accum[i] = stored; // the pass reads the accumulator at its own pixel only: in place
        display_alpha[i] = shown.w;
    }
}
__global__ void __launch_bounds__(256) k_copy_alpha(const float4 *accum, float *display_alpha, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) display_alpha[i] = accum[i].w;
}
__global__ void __launch_bounds__(256) k_taa(TaaImages im, uchar4 *out) {
    const uint32_t n = (uint32_t)im.w * (uint32_t)im.h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = process_taa_pixel(im, (int)(i % (uint32_t)im.w), (int)(i / (uint32_t)im.w));
}

// ---------------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

#define RPTR_MAX_PIPES 4
struct rptr_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr; // option "overlap_shadow": the shadow launch of bounce d runs beside the closest-hit launch of d + 1
    cudaEvent_t ev_shade = nullptr, ev_shadow = nullptr;
    int overlap_shadow = 1;
    // sub-waves in flight side by side (render_waves): pipe 0 = stream / stream2 above
    struct Pipe { cudaStream_t s_main = nullptr, s_shadow = nullptr; cudaEvent_t ev_shade = nullptr, ev_shadow = nullptr, ev_done = nullptr; };
    Pipe pipes[RPTR_MAX_PIPES];
    int n_pipes_ready = 0;
    cudaEvent_t ev_round = nullptr;
    // ray reordering between shade and trace (rptr_reorder.cuh): key mode of the bounce queue / the shadow queue; 0 = off, -1 = chosen per scene
    int reorder_bounce = 0, reorder_shadow = 0;
    int tail_kernel = 1; // the last rays of every trace launch go to k_trace_tail, one warp per ray (rptr_trace_tail.cuh)
    int concurrent_waves = 1; // measured: two sub-waves side by side cost 3 % at 64 spp and 9 % at 8 spp on one GPU (profiles/r02_sweeps.md)
    std::string error;
    // framebuffer
    int32_t width = 0, height = 0;
    float4 *accum = nullptr;
    uchar4 *ldr = nullptr;       // the LDR target of this frame, (width * upscale) x (height * upscale)
    // RenderBackendOptions::render_upscale_factor (configure_for / option "render_upscale_factor"); takes effect at the next initialize, as in
    // the reference (render_vulkan.cpp:255-263 sizes the render targets there)
    int upscale_option = 1, upscale = 1;
    // the temporal build (ENABLE_REALTIME_RESOLVE, CMakeLists.txt:98): option "realtime_resolve"
    int realtime_resolve = 0;
    int enable_taa = 0;             // RenderBackendOptions::enable_taa
    float4 *accum_history = nullptr; // accum_buffers[!active_accum_buffer]: the accumulator as the previous frame left it
    ushort4 *nd_history = nullptr;   // aov_buffers[(!active) * AOVBufferCount + AOVNormalDepthIndex]
    float *display_alpha = nullptr;
    bool display_alpha_valid = false;
    uchar4 *ldr_history = nullptr, *ldr_raw = nullptr; // render_targets[!active_render_target]; snapshot of the target before the TAA pass
    bool ldr_valid = false;          // ctx->ldr holds this frame's image (process_taa ran): readback_u8 returns it as it is
    ushort4 *aov_images[3] = {nullptr, nullptr, nullptr}; // RGBA16F: albedo + roughness, normal + depth, motion + jitter (RenderGraphic::AOVBufferIndex)
    float vp[16] = {0.0f}, vp_reference[16] = {0.0f}; // view_params.VP of the current / previous begin_frame (zero before the first: render_vulkan.cpp:103)
    int aov_buffers = 1;                         // option "aov_buffers": the reference always writes them (ENABLE_AOV_BUFFERS)
    // counters protocol (vulkan/render_vulkan.h:166-168)
    uint32_t frame_id = 0, frame_offset = 0, accumulated_spp = 0;
    bool freeze_frame = false;
    // scene
    bool has_scene = false;
    std::vector<void *> scene_allocs;
    SceneDev scene{};
    BvhDev bvh{};
    float scene_extent = 0.0f; // largest |coordinate| of the scene (box padding scale, range of admissible ray origins)
    int32_t n_lights = 0;
    std::vector<rptr_base_material> materials_host; // resolved materials (texture handles folded in), for per-frame host decisions
    bool any_normal_map = false;
    bool any_textured = false;     // some material parameter refers to a texture larger than 1 x 1 (RPTR_FEAT_TEXTURES shade variant)
    bool any_alpha_tested = false; // some triangle needs the stochastic alpha candidate filter (Alpha variants of the trace kernels)
    std::vector<rptr_tri_light_data> lights_host;
    rptr_scene_params scene_params{};
    bool has_scene_params = false;
    // frame
    rptr_camera_params camera{};
    rptr_render_params params{};
    rptr_light_sampling_config lighting{};
    bool in_frame = false;
    // options
    int transmission = 0;
    int rng_variant = 0; // RenderBackendOptions::rng_variant
    uint32_t *pointset_tables[4] = {nullptr, nullptr, nullptr, nullptr}; // device copies, rptr_cuda_set_pointset_table
    // paths per wave: 64 spp of 1920x1080 in one wave (144 B of path state each, 19 GB); every launch of the bounce loop
    // pays a fixed tail (the longest ray of the queue), so few large waves beat many small ones (profiles/r01_wave_sweep.md)
    int64_t wave_paths = 128ll << 20;
    int stage_timing = 0;
    int trace_kernel = 0; // 0 = persistent while-while (rptr_trace_kernels.cuh), 1 = one ray per thread (A/B reference)
    int bvh_builder = 1;  // 1 = binned SAH on the device (rptr_bvh_build.cu), 0 = the same algorithm on the host (rptr_host.cpp)
    double bvh_build_ms = 0.0;
    int tile_rank = 0, tile_world = 1, tile_rows = 8;
    // wave
    Wave wave{};
    size_t wave_capacity = 0;
    int wave_depth = 0;
    std::vector<void *> wave_allocs;
    DevCounters *dcounters = nullptr;
    // stats
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool frame_timed = false;
    float last_render_ms = 0.0f;
    uint64_t launches = 0, trace_launches = 0;
    double ms_trace = 0, ms_shadow = 0, ms_shade = 0, ms_other = 0;
    struct Timed { cudaEvent_t a, b; int stage; int launches; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> event_pool;
    size_t bytes_now = 0, bytes_max = 0, bytes_total = 0;
    std::unordered_map<void *, size_t> alloc_sizes; // RenderStats::device_bytes_currently_allocated
    // view_params.frame_id / frame_offset as the last begin_frame left them (update_view_parameters, render_vulkan.cpp:2907-2910):
    // what the kernels of that frame -- and of ray queries rendered after it -- see, whatever end_frame does to the counters
    uint32_t view_frame_id = 0, view_frame_offset = 0;
    bool has_view = false;
    // ray queries (enable_ray_queries / render_ray_queries, vulkan/render_vulkan.cpp:430-455, 1867-1876)
    int rq_fixed_budget = 0, rq_per_pixel_budget = 0;
    rptr_render_ray_query *rq_queries = nullptr; // ray_query_buffer
    float4 *rq_results = nullptr;                // ray_result_buffer
    size_t rq_capacity = 0;
    std::vector<void *> rq_allocs;
    // multi-GPU: screen-space sharding + one NCCL reduce of the HDR accumulator per readback
    void *comm = nullptr;          // ncclComm_t
    int comm_world = 1, comm_rank = 0;
    float4 *reduced = nullptr;     // result of the last reduce (the accumulator itself must keep this rank's partial image)
    bool reduced_valid = false;    // a reduce has run since the last frame: readback_f32 then returns the whole image
    // scratch of the RaytraceBackend service (rptr_cuda_trace_rays), grown on demand and kept
    std::vector<void *> tr_allocs;
    size_t tr_capacity = 0;
    rptr_render_ray_query *tr_queries = nullptr;
    float4 *tr_ray_o = nullptr, *tr_ray_d = nullptr, *tr_hit = nullptr, *tr_results = nullptr;
    float *tr_t = nullptr;
    uint32_t *tr_counts = nullptr;
};

static int fail(rptr_ctx *ctx, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    else g_create_error = buf;
    return 1;
}

#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T> static cudaError_t dev_alloc(rptr_ctx *ctx, T **p, size_t n, std::vector<void *> &owner) {
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) {
        owner.push_back(*p);
        ctx->alloc_sizes[*p] = n * sizeof(T);
        ctx->bytes_now += n * sizeof(T);
        ctx->bytes_total += n * sizeof(T);
        if (ctx->bytes_now > ctx->bytes_max) ctx->bytes_max = ctx->bytes_now;
    }
    return e;
}
static void free_all(rptr_ctx *ctx, std::vector<void *> &owner) {
    for (void *p : owner) {
        auto it = ctx->alloc_sizes.find(p);
        if (it != ctx->alloc_sizes.end()) {
            ctx->bytes_now -= it->second;
            ctx->alloc_sizes.erase(it);
        }
        cudaFree(p);
    }
    owner.clear();
}

static TileMap make_tilemap(const rptr_ctx *ctx) {
    TileMap t{};
    t.width = ctx->width; t.height = ctx->height;
    t.rank = ctx->tile_rank; t.world = ctx->tile_world; t.rows = ctx->tile_rows;
    int32_t rows = 0;
    for (int32_t y = 0; y < ctx->height; ++y)
        if ((y / t.rows) % t.world == t.rank) rows++;
    t.local_rows = rows;
    t.local_pixels = rows * ctx->width;
    return t;
}

static int grid_for(const rptr_ctx *ctx, int blocks_per_sm) { return ctx->num_sms * blocks_per_sm; }

static int ensure_wave(rptr_ctx *ctx, size_t paths, int depth) {
    const bool need_rng2 = ctx->rng_variant != 0;
    const bool need_foot = ctx->any_textured; // texture footprints: scenes with image textures only
    const bool need_reorder = ctx->reorder_bounce != 0 || ctx->reorder_shadow != 0;
    if (paths <= ctx->wave_capacity && depth <= ctx->wave_depth && (!need_rng2 || ctx->wave.rng2) && (!need_foot || ctx->wave.foot) &&
        (!need_reorder || ctx->wave.bin_hist))
        return 0;
    if (paths < ctx->wave_capacity) paths = ctx->wave_capacity;
    if (depth < ctx->wave_depth) depth = ctx->wave_depth;
    CU(cudaStreamSynchronize(ctx->stream));
    free_all(ctx, ctx->wave_allocs);
    Wave &w = ctx->wave;
    const size_t n = paths;
    CU(dev_alloc(ctx, &w.ray_o, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.ray_d, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.hit, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.thr, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.illum, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.rngb, n, ctx->wave_allocs));
    w.rng2 = nullptr;
    w.rng3 = nullptr;
    if (need_rng2) {
        CU(dev_alloc(ctx, &w.rng2, n, ctx->wave_allocs));
        CU(dev_alloc(ctx, &w.rng3, n, ctx->wave_allocs));
    }
    w.foot = nullptr;
    if (need_foot) CU(dev_alloc(ctx, &w.foot, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.sh_o, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.sh_d, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.sh_c, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.queue[0], n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.queue[1], n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.counts, (size_t)4 * (depth + 2) * RPTR_MAX_PIPES, ctx->wave_allocs)); // per sub-wave
    CU(dev_alloc(ctx, &w.hitq, n, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.hit_counts, (size_t)(depth + 2) * RPTR_MAX_PIPES, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.tail, (size_t)2 * RPTR_TAIL_CAPACITY(ctx->num_sms) * RPTR_MAX_PIPES, ctx->wave_allocs));
    CU(dev_alloc(ctx, &w.tail_counts, (size_t)4 * (depth + 2) * RPTR_MAX_PIPES, ctx->wave_allocs));
    w.keys_b = w.keys_s = nullptr;
    w.queue_sorted = w.sh_perm = w.bin_hist = nullptr;
    if (need_reorder) {
        CU(dev_alloc(ctx, &w.keys_b, n, ctx->wave_allocs));
        CU(dev_alloc(ctx, &w.keys_s, n, ctx->wave_allocs));
        CU(dev_alloc(ctx, &w.queue_sorted, n, ctx->wave_allocs));
        CU(dev_alloc(ctx, &w.sh_perm, n, ctx->wave_allocs));
        CU(dev_alloc(ctx, &w.bin_hist, (size_t)2 * RPTR_BINS * (depth + 2) * RPTR_MAX_PIPES, ctx->wave_allocs));
    }
    ctx->wave_capacity = paths;
    ctx->wave_depth = depth;
    return 0;
}

static int ensure_pipes(rptr_ctx *ctx, int n) {
    if (!ctx->ev_round) CU(cudaEventCreateWithFlags(&ctx->ev_round, cudaEventDisableTiming));
    for (int k = ctx->n_pipes_ready; k < n; ++k) {
        rptr_ctx::Pipe &p = ctx->pipes[k];
        CU(cudaEventCreateWithFlags(&p.ev_done, cudaEventDisableTiming));
        if (k > 0) {
            CU(cudaStreamCreateWithFlags(&p.s_main, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithFlags(&p.s_shadow, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&p.ev_shade, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p.ev_shadow, cudaEventDisableTiming));
        }
        ctx->n_pipes_ready = k + 1;
    }
    return 0;
}

static cudaEvent_t get_event(rptr_ctx *ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
struct StageTimer {
    rptr_ctx *ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    int stage;
    cudaStream_t stream;
    StageTimer(rptr_ctx *c, int s, cudaStream_t st) : ctx(c), stage(s), stream(st) {
        if (ctx->stage_timing && s >= 0) {
            a = get_event(ctx);
            b = get_event(ctx);
            cudaEventRecord(a, stream);
        }
    }
    ~StageTimer() {
        if (a) {
            cudaEventRecord(b, stream);
            ctx->timed.push_back({a, b, stage, 1});
        }
    }
};
static void collect_timers(rptr_ctx *ctx) {
    for (auto &t : ctx->timed) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            if (t.stage == 0) { ctx->ms_trace += ms; ctx->trace_launches += (uint64_t)t.launches; }
            else if (t.stage == 1) ctx->ms_shade += ms;
            else if (t.stage == 2) ctx->ms_shadow += ms;
            else ctx->ms_other += ms;
        }
        ctx->event_pool.push_back(t.a);
        ctx->event_pool.push_back(t.b);
    }
    ctx->timed.clear();
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {
static void free_temporal_buffers(rptr_ctx *ctx);

const char *rptr_cuda_name(void) { return "CUDA wavefront path tracer (sm_100a)"; }

int rptr_cuda_create(int device_ordinal, rptr_ctx **out) {
    rptr_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, "rptr_cuda_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, "no usable CUDA device (%s); this backend has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_ordinal < 0 || device_ordinal >= n) return fail(nullptr, "device ordinal %d out of range (%d devices)", device_ordinal, n);
    e = cudaSetDevice(device_ordinal);
    if (e != cudaSuccess) return fail(nullptr, "cudaSetDevice(%d): %s", device_ordinal, cudaGetErrorString(e));
    ctx = new rptr_ctx();
    ctx->device = device_ordinal;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_ordinal);
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_shade, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_shadow, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&ctx->ev_begin) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_end) != cudaSuccess || cudaMalloc((void **)&ctx->dcounters, sizeof(DevCounters)) != cudaSuccess) {
        fail(nullptr, "CUDA resource creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return 1;
    }
    cudaMemset(ctx->dcounters, 0, sizeof(DevCounters));
    const int top_bytes = (int)RPTR_TRACE_SMEM_BYTES;
    if (cudaFuncSetAttribute(k_trace_persistent<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, top_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_trace_persistent<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, top_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_trace_persistent<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, top_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_trace_persistent<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, top_bytes) != cudaSuccess) {
        fail(nullptr, "cannot reserve %d bytes of shared memory for the trace kernel: %s", top_bytes, cudaGetErrorString(cudaGetLastError()));
        rptr_cuda_destroy(ctx);
        return 1;
    }
    *out = ctx;
    return 0;
}

void rptr_cuda_destroy(rptr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    collect_timers(ctx);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    free_all(ctx, ctx->scene_allocs);
    free_all(ctx, ctx->wave_allocs);
    free_all(ctx, ctx->rq_allocs);
    free_all(ctx, ctx->tr_allocs);
    for (uint32_t *t : ctx->pointset_tables) cudaFree(t);
    rptr_cuda_comm_destroy(ctx);
    cudaFree(ctx->accum);
    cudaFree(ctx->ldr);
    free_temporal_buffers(ctx);
    cudaFree(ctx->reduced);
    for (ushort4 *im : ctx->aov_images) cudaFree(im);
    cudaFree(ctx->dcounters);
    cudaEventDestroy(ctx->ev_begin);
    cudaEventDestroy(ctx->ev_end);
    cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (rptr_ctx::Pipe &p : ctx->pipes) {
        if (p.s_main) cudaStreamDestroy(p.s_main);
        if (p.s_shadow) cudaStreamDestroy(p.s_shadow);
        if (p.ev_shade) cudaEventDestroy(p.ev_shade);
        if (p.ev_shadow) cudaEventDestroy(p.ev_shadow);
        if (p.ev_done) cudaEventDestroy(p.ev_done);
    }
    if (ctx->ev_round) cudaEventDestroy(ctx->ev_round);
    if (ctx->ev_shade) cudaEventDestroy(ctx->ev_shade);
    if (ctx->ev_shadow) cudaEventDestroy(ctx->ev_shadow);
    delete ctx;
}

const char *rptr_cuda_last_error(const rptr_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

static int ensure_ray_query_buffers(rptr_ctx *ctx);

static void free_temporal_buffers(rptr_ctx *ctx) {
    cudaFree(ctx->accum_history); cudaFree(ctx->nd_history); cudaFree(ctx->display_alpha); cudaFree(ctx->ldr_history); cudaFree(ctx->ldr_raw);
    ctx->accum_history = nullptr; ctx->nd_history = nullptr; ctx->display_alpha = nullptr; ctx->ldr_history = nullptr; ctx->ldr_raw = nullptr;
    ctx->display_alpha_valid = false;
    ctx->ldr_valid = false;
}
// history images of the temporal build, allocated when the option is first used with the current frame size
static int ensure_temporal_buffers(rptr_ctx *ctx) {
    if (ctx->accum_history) return 0;
    const size_t n = (size_t)ctx->width * ctx->height, nl = n * ctx->upscale * ctx->upscale;
    CU(cudaMalloc((void **)&ctx->accum_history, n * sizeof(float4)));
    CU(cudaMalloc((void **)&ctx->nd_history, n * sizeof(ushort4)));
    CU(cudaMalloc((void **)&ctx->display_alpha, n * sizeof(float)));
    CU(cudaMalloc((void **)&ctx->ldr_history, nl * sizeof(uchar4)));
    CU(cudaMalloc((void **)&ctx->ldr_raw, nl * sizeof(uchar4)));
    CU(cudaMemsetAsync(ctx->accum_history, 0, n * sizeof(float4), ctx->stream));
    CU(cudaMemsetAsync(ctx->nd_history, 0, n * sizeof(ushort4), ctx->stream));
    CU(cudaMemsetAsync(ctx->display_alpha, 0, n * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(ctx->ldr_history, 0, nl * sizeof(uchar4), ctx->stream));
    CU(cudaMemsetAsync(ctx->ldr_raw, 0, nl * sizeof(uchar4), ctx->stream));
    return 0;
}

int rptr_cuda_initialize(rptr_ctx *ctx, int32_t width, int32_t height) {
    if (!ctx) return 1;
    if (width <= 0 || height <= 0 || (int64_t)width * height > (1ll << 28)) return fail(ctx, "invalid framebuffer size %dx%d", width, height);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->accum);
    cudaFree(ctx->ldr);
    cudaFree(ctx->reduced);
    free_temporal_buffers(ctx);
    ctx->reduced = nullptr;
    ctx->reduced_valid = false;
    for (ushort4 *im : ctx->aov_images) cudaFree(im);
    ctx->accum = nullptr;
    ctx->ldr = nullptr;
    ctx->upscale = ctx->upscale_option;
    ctx->aov_images[0] = ctx->aov_images[1] = ctx->aov_images[2] = nullptr;
    ctx->width = width;
    ctx->height = height;
    const size_t n = (size_t)width * height;
    CU(cudaMalloc((void **)&ctx->accum, n * sizeof(float4)));
    CU(cudaMalloc((void **)&ctx->ldr, n * sizeof(uchar4) * ctx->upscale * ctx->upscale));
    CU(cudaMemsetAsync(ctx->ldr, 0, n * sizeof(uchar4) * ctx->upscale * ctx->upscale, ctx->stream));
    CU(cudaMemsetAsync(ctx->accum, 0, n * sizeof(float4), ctx->stream));
    for (int i = 0; i < 3; ++i) {
        CU(cudaMalloc((void **)&ctx->aov_images[i], n * sizeof(ushort4)));
        CU(cudaMemsetAsync(ctx->aov_images[i], 0, n * sizeof(ushort4), ctx->stream));
    }
    ctx->frame_id = 0; // vulkan/render_vulkan.cpp:245-249
    ctx->frame_offset = 0;
    ctx->accumulated_spp = 0;
    ctx->in_frame = false;
    ctx->has_view = false;
    if (ctx->rq_per_pixel_budget && ensure_ray_query_buffers(ctx)) return 1; // vulkan/render_vulkan.cpp:366-369
    return 0;
}

// Everything set_scene puts on the device.  It is built completely before the previous scene is released, so a failure
// (invalid input, out of memory, a BVH the traversal stack cannot hold) leaves the context with the scene it had.
struct SceneUpload {
    rptr_ctx *ctx;
    std::vector<void *> allocs;
    SceneDev scene{};
    BvhDev bvh{};
    double bvh_build_ms = 0.0;
    float extent = 0.0f;
    explicit SceneUpload(rptr_ctx *c) : ctx(c) {}
    ~SceneUpload() { free_all(ctx, allocs); }
};

static int upload_scene(rptr_ctx *ctx, HostScene &hs, SceneUpload &up) {
    std::vector<uint64_t *> d_qv(hs.qverts.size(), nullptr), d_qn(hs.qnuv.size(), nullptr);
    std::vector<uint8_t *> d_tm(hs.tri_mat.size(), nullptr);
    for (size_t g = 0; g < hs.qverts.size(); ++g) {
        CU(dev_alloc(ctx, &d_qv[g], hs.qverts[g].size(), up.allocs));
        CU(cudaMemcpy(d_qv[g], hs.qverts[g].data(), hs.qverts[g].size() * 8, cudaMemcpyHostToDevice));
        if (!hs.qnuv[g].empty()) {
            CU(dev_alloc(ctx, &d_qn[g], hs.qnuv[g].size(), up.allocs));
            CU(cudaMemcpy(d_qn[g], hs.qnuv[g].data(), hs.qnuv[g].size() * 8, cudaMemcpyHostToDevice));
        }
    }
    for (size_t p = 0; p < hs.tri_mat.size(); ++p)
        if (!hs.tri_mat[p].empty()) {
            CU(dev_alloc(ctx, &d_tm[p], hs.tri_mat[p].size(), up.allocs));
            CU(cudaMemcpy(d_tm[p], hs.tri_mat[p].data(), hs.tri_mat[p].size(), cudaMemcpyHostToDevice));
        }
    std::vector<GeomInst> gi(hs.ginst.size());
    for (size_t i = 0; i < hs.ginst.size(); ++i) {
        gi[i] = hs.ginst[i].g;
        gi[i].qverts = d_qv[hs.ginst[i].geometry];
        gi[i].qnuv = d_qn[hs.ginst[i].geometry];
        gi[i].tri_mat = d_tm[hs.ginst[i].pmesh] ? d_tm[hs.ginst[i].pmesh] + hs.ginst[i].prim_offset : nullptr;
    }
    GeomInst *d_gi;
    rptr_base_material *d_mat;
    rptr_tri_light_data *d_lights;
    CU(dev_alloc(ctx, &d_gi, gi.size(), up.allocs));
    CU(cudaMemcpy(d_gi, gi.data(), gi.size() * sizeof(GeomInst), cudaMemcpyHostToDevice));
    CU(dev_alloc(ctx, &d_mat, hs.materials.size(), up.allocs));
    CU(cudaMemcpy(d_mat, hs.materials.data(), hs.materials.size() * sizeof(rptr_base_material), cudaMemcpyHostToDevice));
    float4 *d_ntex;
    CU(dev_alloc(ctx, &d_ntex, hs.materials.size(), up.allocs));
    CU(cudaMemcpy(d_ntex, hs.normal_texels.data(), hs.materials.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CU(dev_alloc(ctx, &d_lights, hs.lights.size(), up.allocs));
    if (!hs.lights.empty()) CU(cudaMemcpy(d_lights, hs.lights.data(), hs.lights.size() * sizeof(rptr_tri_light_data), cudaMemcpyHostToDevice));
    // textures larger than 1 x 1 (the others were folded into the materials): RGBA8 texels + the table, and the sRGB decode table
    std::vector<TexDev> tex(hs.textures.size());
    for (size_t t = 0; t < hs.textures.size(); ++t) {
        const HostTexture &ht = hs.textures[t];
        tex[t] = TexDev{nullptr, ht.width, ht.height, ht.srgb, ht.levels};
        if (ht.rgba.empty()) continue;
        uchar4 *d_px;
        CU(dev_alloc(ctx, &d_px, ht.rgba.size() / 4, up.allocs));
        CU(cudaMemcpy(d_px, ht.rgba.data(), ht.rgba.size(), cudaMemcpyHostToDevice));
        tex[t].texels = d_px;
    }
    TexDev *d_tex;
    float *d_lut;
    CU(dev_alloc(ctx, &d_tex, tex.size(), up.allocs));
    if (!tex.empty()) CU(cudaMemcpy(d_tex, tex.data(), tex.size() * sizeof(TexDev), cudaMemcpyHostToDevice));
    CU(dev_alloc(ctx, &d_lut, 256, up.allocs));
    CU(cudaMemcpy(d_lut, hs.srgb_lut, sizeof(hs.srgb_lut), cudaMemcpyHostToDevice));
    up.scene = SceneDev{d_gi, d_mat, d_lights, d_ntex, d_tex, d_lut};

    // largest |coordinate| of the scene: scale of the conservative box padding (rptr_host.cpp build_bvh) and of the range of
    // ray origins the slab test is guaranteed for (begin_frame / trace_rays check it)
    float extent = 0.0f, cmin[3] = {1e30f, 1e30f, 1e30f}, cmax[3] = {-1e30f, -1e30f, -1e30f};
    for (const Tri &t : hs.tris) {
        const float v[3][3] = {{t.v0x, t.v0y, t.v0z}, {t.v0x + t.e1x, t.v0y + t.e1y, t.v0z + t.e1z}, {t.v0x + t.e2x, t.v0y + t.e2y, t.v0z + t.e2z}};
        extent = fmaxf(extent, fmaxf(fmaxf(fabsf(t.v0x), fabsf(t.v0y)), fabsf(t.v0z)) + fmaxf(fmaxf(fabsf(t.e1x), fabsf(t.e1y)), fabsf(t.e1z)) +
                                   fmaxf(fmaxf(fabsf(t.e2x), fabsf(t.e2y)), fabsf(t.e2z)));
        for (int k = 0; k < 3; ++k) {
            const float c = 0.5f * (fminf(v[0][k], fminf(v[1][k], v[2][k])) + fmaxf(v[0][k], fmaxf(v[1][k], v[2][k])));
            cmin[k] = fminf(cmin[k], c);
            cmax[k] = fmaxf(cmax[k], c);
        }
    }
    up.extent = extent;
    if (ctx->bvh_builder == 1) { // device builder (rptr_bvh_build.cu)
        DeviceBvh db;
        std::string err;
        const auto t0 = std::chrono::steady_clock::now();
        if (build_bvh_device(hs.tris, extent, cmin, cmax, ctx->stream, ctx->num_sms, db, err)) {
            up.bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            for (void *p : {(void *)db.nodes, (void *)db.tris, (void *)db.top})
                if (p) up.allocs.push_back(p);
            up.bvh = BvhDev{db.nodes, db.tris, db.n_nodes, db.n_tris, db.top, db.top_k};
            return 0;
        }
        // e.g. the collapsed Morton tree is deeper than the traversal stack allows (clustered or coincident centroids): the host
        // SAH builder bounds the depth (it falls back to balanced median splits), so it takes over instead of failing set_scene
        ctx->error = "device builder: " + err + "; fell back to the host SAH builder";
        try {
            build_bvh(hs);
        } catch (const std::exception &e) {
            return fail(ctx, "set_scene: device builder failed (%s) and so did the host builder (%s)", err.c_str(), e.what());
        }
    }
    BvhNode *d_nodes;
    Tri *d_tris;
    up.bvh_build_ms = hs.bvh_build_ms;
    CU(dev_alloc(ctx, &d_nodes, hs.nodes.size(), up.allocs));
    if (!hs.nodes.empty()) CU(cudaMemcpy(d_nodes, hs.nodes.data(), hs.nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
    CU(dev_alloc(ctx, &d_tris, hs.leaf_tris.size(), up.allocs));
    if (!hs.leaf_tris.empty()) CU(cudaMemcpy(d_tris, hs.leaf_tris.data(), hs.leaf_tris.size() * sizeof(Tri), cudaMemcpyHostToDevice));
    // word planes of the top of the (breadth-first ordered) tree for the trace kernel's shared-memory stage
    const int32_t top_k = (int32_t)std::min<size_t>(hs.nodes.size(), RPTR_TOP_NODES_MAX);
    std::vector<float4> top((size_t)RPTR_NODE_WORDS * RPTR_TOP_NODES_MAX, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    for (int32_t i = 0; i < top_k; ++i)
        for (int wd = 0; wd < RPTR_NODE_WORDS; ++wd) memcpy(&top[(size_t)wd * RPTR_TOP_NODES_MAX + i], reinterpret_cast<const unsigned char *>(&hs.nodes[i]) + (wd << 4), 16);
    float4 *d_top;
    CU(dev_alloc(ctx, &d_top, top.size(), up.allocs));
    CU(cudaMemcpy(d_top, top.data(), top.size() * sizeof(float4), cudaMemcpyHostToDevice));
    up.bvh = BvhDev{d_nodes, d_tris, (int32_t)hs.nodes.size(), (int32_t)hs.leaf_tris.size(), d_top, top_k};
    return 0;
}

int rptr_cuda_set_scene(rptr_ctx *ctx, const rptr_scene_desc *desc, const rptr_light_sampling_config *lighting) {
    if (!ctx) return 1;
    if (!desc) return fail(ctx, "set_scene: desc is NULL");
    if (ctx->in_frame) return fail(ctx, "set_scene inside begin_frame/end_frame");
    CU(cudaSetDevice(ctx->device));
    rptr_light_sampling_config ls;
    if (lighting) ls = *lighting;
    else { ls.light_mis_angle = 0.0f; ls.bin_size = 16; ls.min_perceived_receiver_dist = 15.0f; ls.min_radiance = 0.0f; }
    HostScene hs;
    try {
        build_host_scene(*desc, ls, hs, /*with_bvh=*/ctx->bvh_builder == 0);
    } catch (const std::exception &e) {
        return fail(ctx, "set_scene: %s", e.what());
    }
    ctx->error.clear();
    SceneUpload up(ctx);
    if (upload_scene(ctx, hs, up)) return 1; // the previous scene stays in place
    CU(cudaStreamSynchronize(ctx->stream));
    free_all(ctx, ctx->scene_allocs);
    ctx->scene_allocs.swap(up.allocs); // `up` now owns nothing
    ctx->scene = up.scene;
    ctx->bvh = up.bvh;
    ctx->bvh_build_ms = up.bvh_build_ms;
    ctx->scene_extent = up.extent;
    ctx->n_lights = (int32_t)hs.lights.size();
    ctx->lights_host = hs.lights;
    ctx->any_alpha_tested = hs.any_alpha_tested;
    ctx->any_normal_map = hs.any_normal_map;
    ctx->any_textured = hs.any_textured;
    ctx->materials_host = hs.materials;
    ctx->has_scene = true;
    ctx->frame_id = 0; // vulkan/render_vulkan.cpp:1556
    return 0;
}

int32_t rptr_cuda_get_lights(rptr_ctx *ctx, rptr_tri_light_data *out, int32_t max_lights) {
    if (!ctx) return -1;
    const int32_t n = (int32_t)ctx->lights_host.size();
    if (out) memcpy(out, ctx->lights_host.data(), sizeof(rptr_tri_light_data) * (size_t)(n < max_lights ? n : max_lights));
    return n;
}

int rptr_cuda_set_scene_params(rptr_ctx *ctx, const rptr_scene_params *p) {
    if (!ctx) return 1;
    if (!p) return fail(ctx, "set_scene_params: params is NULL");
    ctx->scene_params = *p;
    ctx->has_scene_params = true;
    return 0;
}

int rptr_cuda_set_option(rptr_ctx *ctx, const char *name, int64_t value) {
    if (!ctx) return 1;
    if (!name) return fail(ctx, "set_option: name is NULL");
    if (ctx->in_frame) return fail(ctx, "set_option(%s) inside begin_frame/end_frame", name);
    const std::string n(name);
    if (n == "transmission") ctx->transmission = value != 0;
    else if (n == "rng_variant") {
        if (value < 0 || value > 3) return fail(ctx, "rng_variant must be 0 (UNIFORM), 1 (BN), 2 (SOBOL) or 3 (Z_SBL)");
        ctx->rng_variant = (int)value;
    } else if (n == "wave_paths") {
        // path slots are 32-bit everywhere (k_raygen's slot loop, AovTarget::slot_lo, slots packed into float bits)
        if (value < 1024 || value > 0x7fffffffll) return fail(ctx, "wave_paths must be in [1024, 2^31 - 1]");
        ctx->wave_paths = value;
    } else if (n == "stage_timing") ctx->stage_timing = value != 0;
    else if (n == "aov_buffers") ctx->aov_buffers = value != 0;
    else if (n == "overlap_shadow") ctx->overlap_shadow = value != 0;
    else if (n == "concurrent_waves") {
        if (value < 1 || value > RPTR_MAX_PIPES) return fail(ctx, "concurrent_waves must be in [1, %d]", RPTR_MAX_PIPES);
        ctx->concurrent_waves = (int)value;
    }
    else if (n == "reorder_bounce" || n == "reorder_shadow") {
        if (value < -1 || value > 4) return fail(ctx, "%s must be -1 (chosen per scene), 0 (off) or a key mode 1..4", name);
        (n == "reorder_bounce" ? ctx->reorder_bounce : ctx->reorder_shadow) = (int)value;
    }
    else if (n == "tail_kernel") ctx->tail_kernel = value != 0;
    else if (n == "trace_kernel") ctx->trace_kernel = (int)value;
    else if (n == "bvh_builder") {
        if (value != 0 && value != 1) return fail(ctx, "bvh_builder must be 0 (host builder) or 1 (device builder)");
        ctx->bvh_builder = (int)value; // takes effect at the next set_scene
    }
    else if (n == "realtime_resolve") {
        if (value != 0 && value != 1) return fail(ctx, "realtime_resolve must be 0 or 1");
        if (ctx->in_frame) return fail(ctx, "realtime_resolve cannot change inside a frame");
        ctx->realtime_resolve = (int)value;
        if (!value) free_temporal_buffers(ctx);
    }
    else if (n == "render_upscale_factor") {
        if (value < 1 || value > 8) return fail(ctx, "render_upscale_factor must be in [1, 8]");
        ctx->upscale_option = (int)value; // sizes the LDR target at the next initialize
    }
    else if (n == "tile_rank") ctx->tile_rank = (int)value;
    else if (n == "tile_world") ctx->tile_world = (int)value;
    else if (n == "tile_rows") ctx->tile_rows = (int)value;
    else return fail(ctx, "unknown option '%s'", name);
    if (ctx->tile_world < 1 || ctx->tile_rows < 1) return fail(ctx, "tile_world and tile_rows must be >= 1");
    return 0;
}

static const size_t k_pointset_table_size[4] = {(size_t)RPTR_SOBOL_DIMS * RPTR_SOBOL_MATRIX_SIZE, (size_t)RPTR_SOBOL_TILE * RPTR_SOBOL_TILE,
                                                (size_t)RPTR_BN_SAMPLES * RPTR_BN_DIMS, (size_t)RPTR_BN_TILE * RPTR_BN_TILE * RPTR_BN_SCRAMBLING_DIMS};

int rptr_cuda_set_pointset_table(rptr_ctx *ctx, int32_t table, const uint32_t *data, size_t count) {
    if (!ctx) return 1;
    if (table < 0 || table > 3) return fail(ctx, "set_pointset_table: unknown table %d", table);
    if (!data) return fail(ctx, "set_pointset_table: data is NULL");
    if (count != k_pointset_table_size[table]) return fail(ctx, "set_pointset_table(%d): expected %zu elements, got %zu", table, k_pointset_table_size[table], count);
    if (ctx->in_frame) return fail(ctx, "set_pointset_table inside begin_frame/end_frame");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (!ctx->pointset_tables[table]) CU(cudaMalloc(&ctx->pointset_tables[table], count * sizeof(uint32_t)));
    CU(cudaMemcpy(ctx->pointset_tables[table], data, count * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return 0;
}

int rptr_cuda_begin_frame(rptr_ctx *ctx, const rptr_camera_params *camera, const rptr_render_params *params,
                          const rptr_light_sampling_config *lighting, int32_t reset_accumulation, int32_t freeze_frame, double time) {
    (void)time;
    if (!ctx) return 1;
    if (!camera || !params) return fail(ctx, "begin_frame: camera/params is NULL");
    if (!ctx->accum) return fail(ctx, "begin_frame before initialize");
    if (!ctx->has_scene) return fail(ctx, "begin_frame before set_scene");
    if (!ctx->has_scene_params) return fail(ctx, "begin_frame before set_scene_params (update_config)");
    if (ctx->in_frame) return fail(ctx, "begin_frame called twice without end_frame");
    if (params->batch_spp < 1) return fail(ctx, "batch_spp must be >= 1");
    if (params->max_path_depth < 1 || params->max_path_depth > 64) return fail(ctx, "max_path_depth out of range");
    if (ctx->tile_rank < 0 || ctx->tile_rank >= ctx->tile_world) return fail(ctx, "tile_rank %d outside tile_world %d", ctx->tile_rank, ctx->tile_world);
    if (ctx->realtime_resolve) {
        if (ctx->tile_world != 1) return fail(ctx, "option realtime_resolve needs the whole frame on one GPU (the passes read neighbouring pixels)");
        if (params->reprojection_mode == RPTR_REPROJECTION_MODE_ACCUMULATE && params->spp_accumulation_window < 1)
            return fail(ctx, "spp_accumulation_window must be >= 1");
        CU(cudaSetDevice(ctx->device));
        if (ensure_temporal_buffers(ctx)) return 1;
    }
    {   // The box tests of the traversal are conservative for ray origins within 8 scene extents of the world origin (padding of
        // 2^-16 |coordinate| + 2^-17 extent against the cancellation in org / d - o / d, rptr_host.cpp build_bvh); farther out
        // true hits could be culled silently, so such a camera is refused instead.
        const float reach = 8.0f * fmaxf(ctx->scene_extent, 1e-3f);
        for (int k = 0; k < 3; ++k)
            if (!(fabsf(camera->pos[k]) <= reach))
                return fail(ctx, "begin_frame: camera position (%g, %g, %g) is outside the supported range of %g (8 x the scene extent %g)",
                            camera->pos[0], camera->pos[1], camera->pos[2], reach, ctx->scene_extent);
    }
    ctx->camera = *camera;
    ctx->params = *params;
    if (lighting) ctx->lighting = *lighting;
    else { ctx->lighting.light_mis_angle = 0.0f; ctx->lighting.bin_size = 16; ctx->lighting.min_perceived_receiver_dist = 15.0f; ctx->lighting.min_radiance = 0.0f; }
    if (ctx->lighting.bin_size < 1 || ctx->lighting.bin_size > RPTR_BINNED_LIGHTS_BIN_MAX_SIZE) return fail(ctx, "bin_size must be in [1, %d]", RPTR_BINNED_LIGHTS_BIN_MAX_SIZE);
    ctx->freeze_frame = freeze_frame != 0;
    if (reset_accumulation) { // vulkan/render_vulkan.cpp:1937-1941
        if (!freeze_frame) ctx->frame_offset += ctx->frame_id;
        ctx->frame_id = 0;
    }
    // update_view_parameters (:2880-2941): VP_reference = the VP of the previous begin_frame (:1986-1998 pass last frame's
    // view_params in either reprojection mode), then this frame's VP
    memcpy(ctx->vp_reference, ctx->vp, sizeof(ctx->vp));
    view_projection(ctx->camera, ctx->width, ctx->height, ctx->vp);
    ctx->view_frame_id = ctx->frame_id;
    ctx->view_frame_offset = ctx->frame_offset;
    ctx->has_view = true;
    ctx->in_frame = true;
    return 0;
}

static FrameParams make_frame_params(const rptr_ctx *ctx) {
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.width = ctx->width; fp.height = ctx->height;
    memcpy(fp.cam_pos, ctx->camera.pos, sizeof(fp.cam_pos));
    view_params(ctx->camera, ctx->width, ctx->height, fp.du, fp.dv, fp.tl);
    fp.frame_offset = ctx->view_frame_offset;
    fp.first_sample = ctx->view_frame_id;
    fp.batch = ctx->params.batch_spp;
    fp.max_path_depth = ctx->params.max_path_depth;
    fp.rr_path_depth = ctx->params.rr_path_depth;
    fp.output_channel = ctx->params.output_channel;
    fp.glossy_only_mode = ctx->params.glossy_only_mode;
    fp.enable_raster_taa = ctx->params.enable_raster_taa;
    fp.pixel_radius = ctx->params.pixel_radius;
    fp.image_textures = ctx->any_textured ? 1 : 0;
    if (fp.enable_raster_taa > 0) screen_jitter(ctx->view_frame_offset, ctx->view_frame_id, ctx->width, ctx->height, fp.screen_jitter);
    memcpy(fp.vp, ctx->vp, sizeof(fp.vp));
    memcpy(fp.vp_reference, ctx->vp_reference, sizeof(fp.vp_reference));
    fp.n_lights = ctx->n_lights;
    fp.bin_size = ctx->lighting.bin_size;
    fp.n_bins = (ctx->n_lights + fp.bin_size - 1) / fp.bin_size; // vulkan/pt_megakernel.glsl:102-103
    fp.transmission = ctx->transmission;
    fp.rng_variant = ctx->rng_variant;
    fp.pts.sobol_matrix = ctx->pointset_tables[RPTR_POINTSET_SOBOL_MATRIX];
    fp.pts.sobol_tile_invert = ctx->pointset_tables[RPTR_POINTSET_SOBOL_TILE_INVERT];
    fp.pts.bn_sobol = ctx->pointset_tables[RPTR_POINTSET_BN_SOBOL];
    fp.pts.bn_scrambling = ctx->pointset_tables[RPTR_POINTSET_BN_SCRAMBLING_1SPP];
    fp.sp = ctx->scene_params;
    if (ctx->n_lights > 0) fp.sp.sun_radiance[3] *= 0.5f; // vulkan/render_sky.cpp:67-70
    else fp.sp.sun_radiance[3] = 1.0f;
    return fp;
}

} // extern "C"

// The wavefront: `batch` sample layers of the pixels (or ray queries) of `tm`, in waves of at most wave_paths paths:
//   raygen -> [closest hit -> shade -> shadow] x max_path_depth -> resolve.
// Frames resolve into the accumulator (queries == nullptr); ray queries (tm.query_wgs_x > 0) take their rays from `queries` and
// resolve into `results` (both device arrays of tm.local_pixels entries).  sample_base = sample index of layer 0.
//
// Option "concurrent_waves" = k > 1 splits a wave into k SUB-WAVES of whole sample layers that are enqueued side by side, each on
// its own pair of streams with its own queues and counters, resolved in sample order.  The idea: every launch of the bounce
// loop ends with a tail (the longest rays of its queue) during which most SMs idle, and the persistent grids of the other
// sub-wave could start on the SMs a kernel vacates.  Measured on the B200 it does not pay: the sub-waves double the number of
// launches (each stages 96 KB of BVH per CTA) and halve their queues while the tails stay as long -- 64 spp: 1275 -> 1238
// Msamples/s, 8 spp: 937 -> 851 (profiles/r02_sweeps.md).  The default is therefore 1; the option stays for experiments.
static int render_waves(rptr_ctx *ctx, const FrameParams &fp, const TileMap &tm, int32_t batch, uint32_t sample_base,
                        const rptr_render_ray_query *queries, float4 *results) {
    const int depth = fp.max_path_depth;
    if (tm.local_pixels <= 0) return 0;
    int64_t layers_per_wave = ctx->wave_paths / tm.local_pixels;
    if (layers_per_wave < 1) layers_per_wave = 1;
    if (layers_per_wave > batch) layers_per_wave = batch;
    int n_pipes = ctx->concurrent_waves < 1 ? 1 : (ctx->concurrent_waves > RPTR_MAX_PIPES ? RPTR_MAX_PIPES : ctx->concurrent_waves);
    if (ctx->trace_kernel != 0) n_pipes = 1; // the one-ray-per-thread A/B kernels are not persistent grids: nothing to interleave
    if (n_pipes > layers_per_wave) n_pipes = (int)layers_per_wave;
    const int64_t layers_per_pipe = layers_per_wave / n_pipes; // whole layers per sub-wave
    const size_t pipe_slots = (size_t)layers_per_pipe * tm.local_pixels;
    if (ensure_wave(ctx, pipe_slots * n_pipes, depth)) return 1;
    if (ensure_pipes(ctx, n_pipes)) return 1;
    // Scenes whose materials take several shading code paths (C4: GGX, thick / thin transmission, emitters): the trace kernel
    // hands the shade stage a dense queue of the paths that hit something, which the shade kernel then sorts tile by tile
    // by code path (C4 shade: 27.1 -> 18.8 ms per 33 M samples).  Single-path scenes (C2: all GGX) keep the bounce queue
    // in its screen / compaction order instead -- the tile sort then only moves the misses aside -- because the
    // retire-order hit queue costs more in scattered path-state reads than the dropped misses save (C2: 1242 -> 1170).
    int multi_path = 0;
    {
        int first_key = -1;
        for (const rptr_base_material &m : ctx->materials_host) {
            int key = m.ior > 1.0f ? 2 : 1;
            if (fp.transmission && m.ior > 1.0f && m.specular_transmission > 0.0f) key = (m.flags & RPTR_BASE_MATERIAL_ONESIDED) ? 4 : 3;
            if (m.emission_intensity != 0.0f) key += 5;
            if (first_key < 0) first_key = key;
            else if (key != first_key) { multi_path = 1; break; }
        }
    }
    const int sort_tiles = multi_path ? 2 : 1; // 1: the tile sort only moves the misses aside (key = hit / miss, no material lookup)
    // per-candidate seeds of alpha-tested shadow rays: view_params.frame_id / frame_offset of this frame (pt_megakernel.glsl:252-254)
    const AlphaFilter alpha_filter{ctx->scene, fp.first_sample, fp.frame_offset, 0u};
    const uint32_t alpha_stride = ctx->rng_variant != 0 ? 1u : 2u;
    // trace: one RPTR_TRACE_THREADS CTA per SM; dynamic smem = the staged top of the BVH + the LUT + the shared stack part
    const int g_trace = grid_for(ctx, 8), g_light = grid_for(ctx, 4), g_pt = grid_for(ctx, 1);
    const size_t top_smem = RPTR_TRACE_SMEM_BYTES;
    // smallest compiled shade variant that covers the features this frame uses (rptr_shading.cuh, RPTR_FEAT_*)
    const int feat = (fp.transmission ? RPTR_FEAT_TRANSMISSION : 0) | (fp.n_lights > 0 ? RPTR_FEAT_TRI_LIGHTS : 0) |
                     (fp.output_channel != 0 ? RPTR_FEAT_AOV : 0) | (fp.rng_variant != 0 ? RPTR_FEAT_QMC : 0) |
                     (ctx->any_normal_map ? RPTR_FEAT_NORMAL_MAPS : 0) | (ctx->any_textured ? RPTR_FEAT_TEXTURES : 0);
    const bool overlap = fp.output_channel == 0 && ctx->overlap_shadow && ctx->trace_kernel == 0;
    // ray reordering (rptr_reorder.cuh).  Chosen per scene (-1): bounce rays by origin cell and octant; shadow rays by beam when the
    // sun is the only light NEE samples (parallel rays), by origin cell and octant otherwise.
    ReorderKeys rk0{};
    if (ctx->trace_kernel == 0 && ctx->wave.bin_hist) {
        rk0.mode_b = ctx->reorder_bounce < 0 ? RPTR_KEY_OCTANT_ORIGIN : ctx->reorder_bounce;
        rk0.mode_s = ctx->reorder_shadow < 0 ? (fp.n_lights == 0 ? RPTR_KEY_BEAM : RPTR_KEY_OCTANT_ORIGIN) : ctx->reorder_shadow;
        rk0.center[0] = rk0.center[1] = rk0.center[2] = 0.0f;
        rk0.inv_half = 1.0f / fmaxf(ctx->scene_extent, 1e-20f);
        const float *sd = fp.sp.sun_dir; // beam basis: any orthonormal pair perpendicular to the sun direction
        const int a = fabsf(sd[0]) <= fabsf(sd[1]) && fabsf(sd[0]) <= fabsf(sd[2]) ? 0 : (fabsf(sd[1]) <= fabsf(sd[2]) ? 1 : 2);
        float e[3] = {0.0f, 0.0f, 0.0f};
        e[a] = 1.0f;
        float u[3] = {sd[1] * e[2] - sd[2] * e[1], sd[2] * e[0] - sd[0] * e[2], sd[0] * e[1] - sd[1] * e[0]};
        const float ul = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (ul > 0.0f) {
            for (int k = 0; k < 3; ++k) u[k] /= ul;
            const float v[3] = {sd[1] * u[2] - sd[2] * u[1], sd[2] * u[0] - sd[0] * u[2], sd[0] * u[1] - sd[1] * u[0]};
            const float vl = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-20f);
            for (int k = 0; k < 3; ++k) { rk0.bu[k] = u[k]; rk0.bv[k] = v[k] / vl; }
        } else if (rk0.mode_s == RPTR_KEY_BEAM)
            rk0.mode_s = RPTR_KEY_OCTANT_ORIGIN;
    }
    const bool use_tail = ctx->tail_kernel && ctx->trace_kernel == 0;
    const int g_tail = grid_for(ctx, 64 / RPTR_TAIL_WARPS); // 4 warps x 4 rays per CTA
    const bool sort_b = rk0.mode_b != RPTR_KEY_NONE && fp.output_channel == 0, sort_s = rk0.mode_s != RPTR_KEY_NONE && fp.output_channel == 0;
    const int g_bin = grid_for(ctx, 4);

    struct Sub { // one sub-wave in flight
        Wave w;
        cudaStream_t s_main, s_shadow;
        cudaEvent_t ev_shade, ev_shadow, ev_done;
        int32_t first, nl;
        AovTarget aov;
        bool shadow_pending;
        cudaEvent_t union_a; // stage_timing: start of the interval in which shadow(d) and closest(d + 1) of this sub-wave share the GPU
        int union_launches;
        uint32_t *hitq, *alpha_lcg;
    };
    for (int32_t round_first = 0; round_first < batch;) {
        Sub subs[RPTR_MAX_PIPES];
        int n_sub = 0;
        for (int k = 0; k < n_pipes && round_first < batch; ++k) {
            Sub &sb = subs[n_sub++];
            sb.first = round_first;
            sb.nl = (int32_t)((batch - round_first) < layers_per_pipe ? (batch - round_first) : layers_per_pipe);
            round_first += sb.nl;
            const size_t off = (size_t)k * pipe_slots;
            const Wave &w0 = ctx->wave;
            sb.w = w0;
            sb.w.ray_o += off; sb.w.ray_d += off; sb.w.hit += off; sb.w.thr += off; sb.w.illum += off; sb.w.rngb += off;
            if (w0.rng2) { sb.w.rng2 += off; sb.w.rng3 += off; }
            if (w0.foot) sb.w.foot += off;
            sb.w.sh_o += off; sb.w.sh_d += off; sb.w.sh_c += off;
            sb.w.queue[0] += off; sb.w.queue[1] += off; sb.w.hitq += off;
            if (w0.bin_hist) {
                sb.w.keys_b += off; sb.w.keys_s += off; sb.w.queue_sorted += off; sb.w.sh_perm += off;
                sb.w.bin_hist += (size_t)k * 2 * RPTR_BINS * (depth + 2);
            }
            sb.w.tail += (size_t)k * 2 * RPTR_TAIL_CAPACITY(ctx->num_sms);
            sb.w.tail_counts += (size_t)k * 4 * (depth + 2);
            sb.w.counts += (size_t)k * 4 * (depth + 2);
            sb.w.hit_counts += (size_t)k * (depth + 2);
            sb.s_main = k == 0 ? ctx->stream : ctx->pipes[k].s_main;
            sb.s_shadow = k == 0 ? ctx->stream2 : ctx->pipes[k].s_shadow;
            sb.ev_shade = k == 0 ? ctx->ev_shade : ctx->pipes[k].ev_shade;
            sb.ev_shadow = k == 0 ? ctx->ev_shadow : ctx->pipes[k].ev_shadow;
            sb.ev_done = ctx->pipes[k].ev_done;
            sb.shadow_pending = false;
            sb.union_a = nullptr;
            sb.union_launches = 0;
            sb.hitq = multi_path ? sb.w.hitq : nullptr;
            // closest-hit alpha draws: the path's LCG (UNIFORM), else the separate LCG of pt_megakernel.glsl:354-358
            sb.alpha_lcg = ctx->rng_variant != 0 ? sb.w.rng3 : reinterpret_cast<uint32_t *>(sb.w.rngb);
            // AOV images: written by the first vertex of the frame's last sample layer (last sub-wave, last layer)
            sb.aov = AovTarget{nullptr, nullptr, nullptr, 0u};
            if (ctx->aov_buffers && !queries && sb.first + sb.nl == batch)
                sb.aov = AovTarget{ctx->aov_images[0], ctx->aov_images[1], ctx->aov_images[2], (uint32_t)(sb.nl - 1) * (uint32_t)tm.local_pixels};
        }
        if (n_sub > 1) { // the side streams start after everything enqueued on the context's stream so far
            CU(cudaEventRecord(ctx->ev_round, ctx->stream));
            for (int k = 1; k < n_sub; ++k) CU(cudaStreamWaitEvent(subs[k].s_main, ctx->ev_round, 0));
        }
        auto join_shadow = [&](Sub &sb) -> int { // the sub-wave's main stream waits for its overlapped shadow launch; closes the union interval
            if (!sb.shadow_pending) return 0;
            CU(cudaStreamWaitEvent(sb.s_main, sb.ev_shadow, 0));
            sb.shadow_pending = false;
            if (sb.union_a) {
                cudaEvent_t e = get_event(ctx);
                CU(cudaEventRecord(e, sb.s_main));
                ctx->timed.push_back({sb.union_a, e, 0, sb.union_launches});
                sb.union_a = nullptr;
            }
            return 0;
        };
        for (int k = 0; k < n_sub; ++k) {
            Sub &sb = subs[k];
            CU(cudaMemsetAsync(sb.w.counts, 0, sizeof(uint32_t) * 4 * (depth + 2), sb.s_main));
            CU(cudaMemsetAsync(sb.w.hit_counts, 0, sizeof(uint32_t) * (depth + 2), sb.s_main));
            CU(cudaMemsetAsync(sb.w.tail_counts, 0, sizeof(uint32_t) * 4 * (depth + 2), sb.s_main));
            if (sort_b || sort_s) CU(cudaMemsetAsync(sb.w.bin_hist, 0, sizeof(uint32_t) * 2 * RPTR_BINS * (depth + 2), sb.s_main));
            StageTimer t(ctx, 3, sb.s_main);
            Wave wr = sb.w;
            if (!ctx->any_alpha_tested) wr.rng3 = nullptr; // no alpha-tested triangle: nobody reads the alpha LCG
            k_raygen<<<g_light, 256, 0, sb.s_main>>>(fp, tm, wr, sample_base, sb.first, sb.nl, queries);
            ctx->launches++;
        }
        for (int d = 0; d < depth; ++d)
            for (int k = 0; k < n_sub; ++k) {
                Sub &sb = subs[k];
                Wave &w = sb.w;
                // counts[4d] = live paths entering bounce d, [4d+1] = its shadow rays, [4d+2], [4d+3] = fetch cursors
                const uint32_t *q = d == 0 ? nullptr : w.queue[d & 1];
                const uint32_t *q_trace = d > 0 && sort_b ? w.queue_sorted : q; // same entries, bin order
                uint32_t *nq = w.queue[(d + 1) & 1];
                uint32_t *cn = w.counts + 4 * d;
                ReorderKeys rk = rk0;
                rk.keys_b = sort_b && d + 1 < depth ? w.keys_b : nullptr;
                rk.keys_s = sort_s ? w.keys_s : nullptr;
                {
                    // inside a union interval the closest-hit launch is timed together with the shadow launch it overlaps
                    StageTimer t(ctx, sb.union_a ? -1 : 0, sb.s_main);
                    if (sb.union_a) sb.union_launches++;
                    if (ctx->trace_kernel == 0) {
                        TraceIO io{w.ray_o, w.ray_d, q_trace, cn, cn + 2, w.hit, sb.hitq, w.hit_counts + d, nullptr, nullptr, sb.alpha_lcg, alpha_stride, alpha_filter, tm,
                                   use_tail ? w.tail : nullptr, w.tail_counts + 4 * d, w.tail_counts + 4 * d + 2};
                        auto kernel = ctx->any_alpha_tested ? k_trace_persistent<false, true> : k_trace_persistent<false, false>;
                        kernel<<<g_pt, RPTR_TRACE_THREADS, top_smem, sb.s_main>>>(
                            ctx->bvh, io, &ctx->dcounters->closest_rays, &ctx->dcounters->closest_nodes, &ctx->dcounters->closest_tris);
                        if (use_tail) {
                            auto tail = ctx->any_alpha_tested ? k_trace_tail<false, true> : k_trace_tail<false, false>;
                            tail<<<g_tail, RPTR_TAIL_WARPS * 32, 0, sb.s_main>>>(ctx->bvh, io, &ctx->dcounters->closest_nodes, &ctx->dcounters->closest_tris);
                            ctx->launches++;
                        }
                    } else
                        k_trace<<<g_trace, 128, 0, sb.s_main>>>(ctx->bvh, ctx->scene, w, q, cn, ctx->dcounters, sb.hitq ? w.hit_counts + d : nullptr);
                    ctx->launches++;
                }
                if (join_shadow(sb)) return 1; // shade reads illum and rewrites the shadow queue: the overlapped shadow launch must be done
                {
                    StageTimer t(ctx, 1, sb.s_main);
#define RPTR_SHADE_ARGS fp, ctx->scene, ctx->bvh, w, (sb.hitq ? sb.hitq : q), (sb.hitq ? w.hit_counts + d : cn), nq, cn + 4, cn + 1, ctx->dcounters, (d == 0 ? sb.aov : AovTarget{nullptr, nullptr, nullptr, 0u}), tm, sort_tiles, (d == 0 ? 1 : 0), rk
                    if (feat == 0) k_shade<0><<<g_trace, RPTR_SHADE_THREADS, 0, sb.s_main>>>(RPTR_SHADE_ARGS);
                    else if (feat == RPTR_FEAT_TRI_LIGHTS) k_shade<RPTR_FEAT_TRI_LIGHTS><<<g_trace, RPTR_SHADE_THREADS, 0, sb.s_main>>>(RPTR_SHADE_ARGS);
                    else k_shade<RPTR_FEAT_ALL><<<g_trace, RPTR_SHADE_THREADS, 0, sb.s_main>>>(RPTR_SHADE_ARGS);
#undef RPTR_SHADE_ARGS
                    ctx->launches++;
                }
                if (fp.output_channel != 0 || d + 1 >= depth) continue; // no next-event estimation: no shadow rays
                auto bin_pass = [&](const uint16_t *keys, const uint32_t *values, const uint32_t *count, uint32_t *hist, uint32_t *out, cudaStream_t st) {
                    k_bin_count<<<g_bin, RPTR_BIN_THREADS, 0, st>>>(keys, count, hist);
                    k_bin_scan<<<1, 1024, 0, st>>>(hist);
                    k_bin_scatter<<<g_bin, RPTR_BIN_THREADS, 0, st>>>(keys, values, count, hist, out);
                    ctx->launches += 3;
                };
                TraceIO io{w.sh_o, w.sh_d, sort_s ? w.sh_perm : nullptr, cn + 1, cn + 3, nullptr, nullptr, nullptr, w.sh_c, w.illum, sb.alpha_lcg, alpha_stride, alpha_filter, tm,
                           use_tail ? w.tail + RPTR_TAIL_CAPACITY(ctx->num_sms) : nullptr, w.tail_counts + 4 * d + 1, w.tail_counts + 4 * d + 3};
                auto shadow_tail = [&](cudaStream_t st) {
                    if (!use_tail) return;
                    auto tail = ctx->any_alpha_tested ? k_trace_tail<true, true> : k_trace_tail<true, false>;
                    tail<<<g_tail, RPTR_TAIL_WARPS * 32, 0, st>>>(ctx->bvh, io, &ctx->dcounters->shadow_nodes, &ctx->dcounters->shadow_tris);
                    ctx->launches++;
                };
                if (overlap) {
                    // The shadow rays of bounce d and the closest-hit rays of bounce d + 1 are independent.  Both kernels are
                    // persistent grids of one CTA per SM, so launched on two streams the second fills the SMs the first one
                    // vacates: its head hides the first one's tail (the longest rays of the queue).
                    if (ctx->stage_timing) {
                        sb.union_a = get_event(ctx);
                        CU(cudaEventRecord(sb.union_a, sb.s_main));
                        sb.union_launches = 1;
                    }
                    CU(cudaEventRecord(sb.ev_shade, sb.s_main));
                    CU(cudaStreamWaitEvent(sb.s_shadow, sb.ev_shade, 0));
                    if (sort_s) bin_pass(w.keys_s, nullptr, cn + 1, w.bin_hist + (size_t)(2 * d + 1) * RPTR_BINS, w.sh_perm, sb.s_shadow);
                    if (sort_b) bin_pass(w.keys_b, nq, cn + 4, w.bin_hist + (size_t)(2 * d) * RPTR_BINS, w.queue_sorted, sb.s_main);
                    auto kernel = ctx->any_alpha_tested ? k_trace_persistent<true, true> : k_trace_persistent<true, false>;
                    kernel<<<g_pt, RPTR_TRACE_THREADS, top_smem, sb.s_shadow>>>(
                        ctx->bvh, io, &ctx->dcounters->shadow_rays, &ctx->dcounters->shadow_nodes, &ctx->dcounters->shadow_tris);
                    shadow_tail(sb.s_shadow);
                    CU(cudaEventRecord(sb.ev_shadow, sb.s_shadow));
                    sb.shadow_pending = true;
                } else {
                    StageTimer t(ctx, 2, sb.s_main);
                    if (sort_s) bin_pass(w.keys_s, nullptr, cn + 1, w.bin_hist + (size_t)(2 * d + 1) * RPTR_BINS, w.sh_perm, sb.s_main);
                    if (sort_b) bin_pass(w.keys_b, nq, cn + 4, w.bin_hist + (size_t)(2 * d) * RPTR_BINS, w.queue_sorted, sb.s_main);
                    if (ctx->trace_kernel == 0) {
                        auto kernel = ctx->any_alpha_tested ? k_trace_persistent<true, true> : k_trace_persistent<true, false>;
                        kernel<<<g_pt, RPTR_TRACE_THREADS, top_smem, sb.s_main>>>(
                            ctx->bvh, io, &ctx->dcounters->shadow_rays, &ctx->dcounters->shadow_nodes, &ctx->dcounters->shadow_tris);
                        shadow_tail(sb.s_main);
                    } else
                        k_shadow<<<g_trace, 128, 0, sb.s_main>>>(ctx->bvh, w, cn + 1, ctx->dcounters, alpha_filter, tm);
                }
                ctx->launches++;
            }
        // resolve in sample order: sub-wave k folds its layers after sub-wave k - 1 has
        for (int k = 0; k < n_sub; ++k) {
            Sub &sb = subs[k];
            if (join_shadow(sb)) return 1;
            if (k > 0) CU(cudaStreamWaitEvent(sb.s_main, subs[k - 1].ev_done, 0));
            {
                StageTimer t(ctx, 3, sb.s_main);
                if (queries)
                    k_resolve_queries<<<g_light, 256, 0, sb.s_main>>>(fp, sb.w, results, (uint32_t)tm.local_pixels, sample_base + (uint32_t)sb.first, sb.nl, ctx->dcounters);
                else
                    // the temporal build folds the history in k_reproject (draw_frame), from the frame's own sample
                    k_resolve<<<g_light, 256, 0, sb.s_main>>>(fp, tm, sb.w, ctx->accum, sample_base + (uint32_t)sb.first, sb.nl, ctx->dcounters, sb.aov,
                                                              ctx->params.reprojection_mode == RPTR_REPROJECTION_MODE_DISCARD_HISTORY ||
                                                                  (ctx->realtime_resolve && ctx->params.reprojection_mode == RPTR_REPROJECTION_MODE_ACCUMULATE));
                ctx->launches++;
            }
            if (n_sub > 1) CU(cudaEventRecord(sb.ev_done, sb.s_main));
        }
        if (n_sub > 1) CU(cudaStreamWaitEvent(ctx->stream, subs[n_sub - 1].ev_done, 0)); // the context's stream continues after the whole round
    }
    return 0;
}

extern "C" {

int rptr_cuda_draw_frame(rptr_ctx *ctx, int32_t variant) {
    (void)variant;
    if (!ctx) return 1;
    if (!ctx->in_frame) return fail(ctx, "draw_frame outside begin_frame/end_frame");
    {   // the Sobol / blue-noise samplers read the reference's tables (vulkan/pointsets/render_{sobol,bn}.cpp upload them)
        const int v = ctx->rng_variant;
        const bool ok = v == 0 || (v == 1 && ctx->pointset_tables[2] && ctx->pointset_tables[3]) || (v == 2 && ctx->pointset_tables[0]) ||
                        (v == 3 && ctx->pointset_tables[0] && ctx->pointset_tables[1]);
        if (!ok) return fail(ctx, "rng_variant %d needs its tables: call rptr_cuda_set_pointset_table first", v);
    }
    CU(cudaSetDevice(ctx->device));
    const TileMap tm = make_tilemap(ctx);
    FrameParams fp = make_frame_params(ctx);
    ctx->reduced_valid = false;
    CU(cudaEventRecord(ctx->ev_begin, ctx->stream));
    if (render_waves(ctx, fp, tm, fp.batch, fp.first_sample, nullptr, nullptr)) return 1;
    ctx->ldr_valid = false;
    ctx->display_alpha_valid = false;
    if (ctx->realtime_resolve) {
        const size_t n = (size_t)ctx->width * ctx->height;
        if (ctx->params.reprojection_mode == RPTR_REPROJECTION_MODE_ACCUMULATE && fp.first_sample > 0) { // process_samples.comp:106-113
            if (!ctx->aov_buffers) return fail(ctx, "reprojection_mode ACCUMULATE needs the AOV images (option aov_buffers)");
            const ResolveImages im{ctx->width, ctx->height, ctx->accum_history, ctx->nd_history, ctx->aov_images[1], ctx->aov_images[2]};
            k_reproject<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(im, ctx->accum, ctx->display_alpha,
                                                                  1.0f / (float)ctx->params.spp_accumulation_window, ctx->params.batch_spp);
            ctx->launches++;
            ctx->display_alpha_valid = true;
        }
        // this frame's accumulator and normal / depth image are the next frame's history (the reference swaps two sets of images)
        CU(cudaMemcpyAsync(ctx->accum_history, ctx->accum, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->nd_history, ctx->aov_images[1], n * sizeof(ushort4), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CU(cudaGetLastError());
    return 0;
}

int rptr_cuda_end_frame(rptr_ctx *ctx, int32_t variant) {
    (void)variant;
    if (!ctx) return 1;
    if (!ctx->in_frame) return fail(ctx, "end_frame without begin_frame");
    CU(cudaEventRecord(ctx->ev_end, ctx->stream));
    ctx->frame_timed = true;
    ctx->accumulated_spp = ctx->frame_id + (uint32_t)ctx->params.batch_spp; // vulkan/render_vulkan.cpp:2152-2154
    if (!ctx->freeze_frame) ctx->frame_id += (uint32_t)ctx->params.batch_spp;
    ctx->in_frame = false;
    return 0;
}

int rptr_cuda_flush(rptr_ctx *ctx) {
    if (!ctx) return 1;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    collect_timers(ctx);
    return 0;
}

int rptr_cuda_stats(rptr_ctx *ctx, rptr_render_stats *out) {
    if (!ctx || !out) return 1;
    if (rptr_cuda_flush(ctx)) return 1;
    memset(out, 0, sizeof(*out));
    if (ctx->frame_timed) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) == cudaSuccess) ctx->last_render_ms = ms;
    }
    out->has_valid_frame_stats = ctx->last_render_ms != 0.0f;
    out->render_time = ctx->last_render_ms;
    out->rays_per_second = -1.0f;
    out->frame_stats_delay = 0;
    out->spp = (int32_t)ctx->accumulated_spp;
    out->total_device_bytes_allocated = ctx->bytes_total;
    out->max_device_bytes_allocated = ctx->bytes_max;
    out->device_bytes_currently_allocated = ctx->bytes_now;
    return 0;
}

int rptr_cuda_get_counters(rptr_ctx *ctx, rptr_counters *out) {
    if (!ctx || !out) return 1;
    if (rptr_cuda_flush(ctx)) return 1;
    DevCounters dc;
    CU(cudaMemcpy(&dc, ctx->dcounters, sizeof(dc), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    out->samples = dc.samples;
    out->closest_rays = dc.closest_rays;
    out->shadow_rays = dc.shadow_rays;
    out->shaded_vertices = dc.shaded_vertices;
    out->closest_nodes = dc.closest_nodes;
    out->closest_tris = dc.closest_tris;
    out->shadow_nodes = dc.shadow_nodes;
    out->shadow_tris = dc.shadow_tris;
    out->launches = ctx->launches;
    out->ms_trace = ctx->ms_trace;
    out->ms_shadow = ctx->ms_shadow;
    out->ms_shade = ctx->ms_shade;
    out->ms_other = ctx->ms_other;
    out->trace_launches = ctx->trace_launches;
    out->trace_overlap = (uint64_t)(ctx->overlap_shadow && ctx->trace_kernel == 0);
    out->node_bytes = sizeof(BvhNode);
    out->tri_bytes = sizeof(Tri);
    out->bvh_nodes = (uint64_t)ctx->bvh.n_nodes;
    out->bvh_build_ms = ctx->bvh_build_ms;
    out->num_sms = (uint64_t)ctx->num_sms;
    return 0;
}

int rptr_cuda_reset_counters(rptr_ctx *ctx) {
    if (!ctx) return 1;
    if (rptr_cuda_flush(ctx)) return 1;
    CU(cudaMemset(ctx->dcounters, 0, sizeof(DevCounters)));
    ctx->launches = 0;
    ctx->trace_launches = 0;
    ctx->ms_trace = ctx->ms_shadow = ctx->ms_shade = ctx->ms_other = 0.0;
    return 0;
}

int rptr_cuda_frame_state(rptr_ctx *ctx, uint32_t *frame_id, uint32_t *frame_offset, uint32_t *accumulated_spp) {
    if (!ctx) return 1;
    if (frame_id) *frame_id = ctx->frame_id;
    if (frame_offset) *frame_offset = ctx->frame_offset;
    if (accumulated_spp) *accumulated_spp = ctx->accumulated_spp;
    return 0;
}

int rptr_cuda_framebuffer_size(rptr_ctx *ctx, uint32_t *width, uint32_t *height, uint32_t *channels) {
    if (!ctx) return 1;
    // RenderVulkan::get_framebuffer_size (render_vulkan.cpp:2250-2254): the dimensions of the (upscaled) LDR render target
    if (width) *width = (uint32_t)(ctx->width * ctx->upscale);
    if (height) *height = (uint32_t)(ctx->height * ctx->upscale);
    if (channels) *channels = 4;
    return 0;
}

size_t rptr_cuda_readback_f32(rptr_ctx *ctx, size_t n_elems, float *dst) {
    if (!ctx || !dst || !ctx->accum) return 0;
    const size_t size = (size_t)ctx->width * ctx->height * 4;
    if (n_elems < size) return 0; // vulkan/render_vulkan.cpp:2262-2263
    if (cudaSetDevice(ctx->device) != cudaSuccess) return 0;
    // sharded over several GPUs: after rptr_cuda_reduce_framebuffer the whole image, otherwise this rank's bands
    const float4 *src = (ctx->reduced && ctx->reduced_valid) ? ctx->reduced : ctx->accum;
    if (cudaMemcpyAsync(dst, src, size * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        fail(ctx, "readback failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    return size;
}

// process_samples.comp:138-200 into the LDR target
static void make_ldr(rptr_ctx *ctx) {
    const float scale = exp2f(ctx->params.exposure);
    const bool aov_on = ctx->aov_buffers != 0;
    k_to_srgb8<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(ctx->accum, ctx->display_alpha_valid ? ctx->display_alpha : nullptr,
                                                          aov_on ? ctx->aov_images[0] : nullptr, aov_on ? ctx->aov_images[1] : nullptr,
                                                          aov_on ? ctx->aov_images[2] : nullptr, ctx->ldr, (uint32_t)((size_t)ctx->width * ctx->height), scale,
                                                          ctx->params.output_channel, ctx->params.output_moment, ctx->params.early_tone_mapping_mode,
                                                          (float)ctx->width, (float)ctx->height, ctx->upscale);
    ctx->launches++;
}

size_t rptr_cuda_readback_u8(rptr_ctx *ctx, size_t n_elems, uint8_t *dst) {
    if (!ctx || !dst || !ctx->accum) return 0;
    const size_t size = (size_t)ctx->width * ctx->height * 4 * ctx->upscale * ctx->upscale; // render_vulkan.cpp:255-263: the targets are upscaled
    if (n_elems < size) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return 0;
    if (!ctx->ldr_valid) make_ldr(ctx); // after process_taa the target already holds the frame
    if (cudaMemcpyAsync(dst, ctx->ldr, size, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
    return size;
}

// ProcessTAAVulkan::process (vulkan/processing/process_taa.cpp:93-136), run by the application after end_frame when
// options.enable_taa && params.reprojection_mode != NONE (app.cpp:517-520).
int rptr_cuda_process_taa(rptr_ctx *ctx) {
    if (!ctx) return 1;
    if (!ctx->accum) return fail(ctx, "process_taa before initialize");
    if (ctx->in_frame) return fail(ctx, "process_taa inside begin_frame/end_frame (the pass runs after end_frame)");
    if (!ctx->realtime_resolve) return fail(ctx, "process_taa needs option realtime_resolve (the reference builds the pass with ENABLE_REALTIME_RESOLVE only)");
    if (!ctx->aov_buffers) return fail(ctx, "process_taa needs the motion image (option aov_buffers)");
    CU(cudaSetDevice(ctx->device));
    if (ensure_temporal_buffers(ctx)) return 1;
    if (ctx->ldr_valid) return 0; // already run for this frame
    const size_t nl = (size_t)ctx->width * ctx->height * ctx->upscale * ctx->upscale;
    make_ldr(ctx);
    if (ctx->frame_id > 1) { // process_taa.cpp:95-96
        CU(cudaMemcpyAsync(ctx->ldr_raw, ctx->ldr, nl * sizeof(uchar4), cudaMemcpyDeviceToDevice, ctx->stream));
        const TaaImages im{ctx->width * ctx->upscale, ctx->height * ctx->upscale, ctx->upscale, ctx->width, ctx->height, ctx->ldr_raw, ctx->ldr_history,
                           ctx->aov_images[2]};
        k_taa<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(im, ctx->ldr);
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(ctx->ldr_history, ctx->ldr, nl * sizeof(uchar4), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaGetLastError());
    ctx->ldr_valid = true;
    return 0;
}

size_t rptr_cuda_readback_aov(rptr_ctx *ctx, int32_t aov_index, size_t n_elems, uint16_t *dst) {
    if (!ctx || !dst || !ctx->accum) return 0;
    if (aov_index < 0 || aov_index > 2 || !ctx->aov_buffers) return 0;
    const size_t n = (size_t)ctx->width * ctx->height * 4;
    if (n_elems < n) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return 0;
    if (cudaMemcpyAsync(dst, ctx->aov_images[aov_index], n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
    return n;
}

int rptr_cuda_framebuffer_device_ptr(rptr_ctx *ctx, void **ptr) {
    if (!ctx || !ptr) return 1;
    if (!ctx->accum) return fail(ctx, "framebuffer_device_ptr before initialize");
    if (rptr_cuda_flush(ctx)) return 1;
    *ptr = ctx->accum;
    return 0;
}

int rptr_cuda_stream_handle(rptr_ctx *ctx, void **stream) {
    if (!ctx || !stream) return 1;
    *stream = (void *)ctx->stream;
    return 0;
}

static int origins_in_reach(rptr_ctx *ctx, const rptr_render_ray_query *queries, int32_t n, const char *who, bool honour_skip) {
    const float reach = 8.0f * fmaxf(ctx->scene_extent, 1e-3f); // see begin_frame
    for (int32_t i = 0; i < n; ++i)
        if ((!honour_skip || queries[i].mode_or_data >= 0) &&
            !(fabsf(queries[i].origin[0]) <= reach && fabsf(queries[i].origin[1]) <= reach && fabsf(queries[i].origin[2]) <= reach))
            return fail(ctx, "%s: origin of query %d is outside the supported range of %g (8 x the scene extent)", who, i, reach);
    return 0;
}

int rptr_cuda_trace_rays(rptr_ctx *ctx, const rptr_render_ray_query *queries, int32_t n, float *results, float *hit_t) {
    if (!ctx) return 1;
    if (!ctx->has_scene) return fail(ctx, "trace_rays before set_scene");
    if (n < 0 || (n > 0 && (!queries || !results))) return fail(ctx, "trace_rays: invalid arguments");
    if (n == 0) return 0;
    if (origins_in_reach(ctx, queries, n, "trace_rays", true)) return 1;
    CU(cudaSetDevice(ctx->device));
    if ((size_t)n > ctx->tr_capacity) { // scratch is kept between calls and only ever grows
        CU(cudaStreamSynchronize(ctx->stream));
        free_all(ctx, ctx->tr_allocs);
        ctx->tr_capacity = 0;
        const size_t cap = (size_t)n + (size_t)n / 4;
        CU(dev_alloc(ctx, &ctx->tr_queries, cap, ctx->tr_allocs));
        CU(dev_alloc(ctx, &ctx->tr_ray_o, cap, ctx->tr_allocs));
        CU(dev_alloc(ctx, &ctx->tr_ray_d, cap, ctx->tr_allocs));
        CU(dev_alloc(ctx, &ctx->tr_hit, cap, ctx->tr_allocs));
        CU(dev_alloc(ctx, &ctx->tr_results, cap, ctx->tr_allocs));
        CU(dev_alloc(ctx, &ctx->tr_t, cap, ctx->tr_allocs));
        CU(dev_alloc(ctx, &ctx->tr_counts, 4, ctx->tr_allocs));
        ctx->tr_capacity = cap;
    }
    rptr_render_ray_query *dq = ctx->tr_queries;
    float4 *dr = ctx->tr_results;
    float *dt = ctx->tr_t;
    CU(cudaMemcpyAsync(dq, queries, sizeof(rptr_render_ray_query) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    // slots of skipped queries (mode < 0) keep what the caller's buffers hold (rt_intersect.comp:44-45)
    CU(cudaMemcpyAsync(dr, results, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (hit_t) CU(cudaMemcpyAsync(dt, hit_t, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->trace_kernel == 0) {
        k_rq_prepare<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(dq, n, ctx->tr_ray_o, ctx->tr_ray_d, ctx->tr_counts);
        TraceIO io{ctx->tr_ray_o, ctx->tr_ray_d, nullptr, ctx->tr_counts, ctx->tr_counts + 1, ctx->tr_hit, nullptr, nullptr, nullptr, nullptr, nullptr, 0u,
                   AlphaFilter{SceneDev{}, 0u, 0u, 0u}, TileMap{}, nullptr, nullptr, nullptr};
        k_trace_persistent<false, false><<<grid_for(ctx, 1), RPTR_TRACE_THREADS, RPTR_TRACE_SMEM_BYTES, ctx->stream>>>(
            ctx->bvh, io, &ctx->dcounters->closest_rays, &ctx->dcounters->closest_nodes, &ctx->dcounters->closest_tris);
        k_rq_pack<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(ctx->bvh, dq, n, ctx->tr_hit, dr, hit_t ? dt : nullptr);
        ctx->launches += 3;
    } else {
        k_ray_queries<<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->bvh, dq, n, dr, hit_t ? dt : nullptr);
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(results, dr, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (hit_t) CU(cudaMemcpyAsync(hit_t, dt, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return 0;
}

// ---- ray queries through the integrator -------------------------------------------------------------------------------------
static int ensure_ray_query_buffers(rptr_ctx *ctx) {
    size_t budget = (size_t)(ctx->accum ? (size_t)ctx->width * ctx->height : 0) * (size_t)ctx->rq_per_pixel_budget;
    if (budget < (size_t)ctx->rq_fixed_budget) budget = (size_t)ctx->rq_fixed_budget;
    if (budget == ctx->rq_capacity) return 0;
    CU(cudaStreamSynchronize(ctx->stream));
    free_all(ctx, ctx->rq_allocs);
    ctx->rq_capacity = 0;
    ctx->rq_queries = nullptr;
    ctx->rq_results = nullptr;
    if (budget == 0) return 0;
    CU(dev_alloc(ctx, &ctx->rq_queries, budget, ctx->rq_allocs));
    CU(dev_alloc(ctx, &ctx->rq_results, budget, ctx->rq_allocs));
    CU(cudaMemsetAsync(ctx->rq_results, 0, budget * sizeof(float4), ctx->stream));
    ctx->rq_capacity = budget;
    return 0;
}

int rptr_cuda_enable_ray_queries(rptr_ctx *ctx, int32_t max_queries, int32_t max_queries_per_pixel) {
    if (!ctx) return 1;
    if (max_queries < 0 || max_queries_per_pixel < 0) return fail(ctx, "enable_ray_queries: negative budget");
    CU(cudaSetDevice(ctx->device));
    ctx->rq_fixed_budget = max_queries;
    ctx->rq_per_pixel_budget = max_queries_per_pixel;
    return ensure_ray_query_buffers(ctx); // before initialize() only the fixed budget counts; initialize() sizes them again (:366-369)
}

int rptr_cuda_ray_query_buffers(rptr_ctx *ctx, void **queries, void **results, size_t *capacity) {
    if (!ctx) return 1;
    if (queries) *queries = ctx->rq_queries;
    if (results) *results = ctx->rq_results;
    if (capacity) *capacity = ctx->rq_capacity;
    return 0;
}

int rptr_cuda_write_ray_queries(rptr_ctx *ctx, const rptr_render_ray_query *queries, int32_t first, int32_t n) {
    if (!ctx) return 1;
    if (n < 0 || first < 0 || (n > 0 && !queries)) return fail(ctx, "write_ray_queries: invalid arguments");
    if ((size_t)first + (size_t)n > ctx->rq_capacity)
        return fail(ctx, "write_ray_queries: queries [%d, %d) exceed the budget of %zu set by enable_ray_queries", first, first + n, ctx->rq_capacity);
    if (n == 0) return 0;
    if (ctx->has_scene && origins_in_reach(ctx, queries, n, "write_ray_queries", false)) return 1;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->rq_queries + first, queries, sizeof(rptr_render_ray_query) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // the host buffer is only borrowed for the call
    return 0;
}

int rptr_cuda_read_ray_results(rptr_ctx *ctx, float *results, int32_t first, int32_t n) {
    if (!ctx) return 1;
    if (n < 0 || first < 0 || (n > 0 && !results)) return fail(ctx, "read_ray_results: invalid arguments");
    if ((size_t)first + (size_t)n > ctx->rq_capacity) return fail(ctx, "read_ray_results: range exceeds the ray-query budget of %zu", ctx->rq_capacity);
    if (n == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(results, ctx->rq_results + first, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int rptr_cuda_render_ray_queries(rptr_ctx *ctx, int32_t num_queries, const rptr_render_params *params, int32_t variant) {
    (void)params; // the reference does not read it either: record_frame uses this->params (vulkan/render_vulkan.cpp:1867-1876, 2965)
    (void)variant;
    if (!ctx) return 1;
    if (!ctx->has_scene) return fail(ctx, "render_ray_queries before set_scene");
    if (!ctx->has_view) return fail(ctx, "render_ray_queries before the first begin_frame (it renders with that frame's view and render parameters)");
    if (num_queries < 0) return fail(ctx, "render_ray_queries: negative query count");
    if ((size_t)num_queries > ctx->rq_capacity)
        return fail(ctx, "render_ray_queries: %d queries exceed the budget of %zu set by enable_ray_queries", num_queries, ctx->rq_capacity);
    if (num_queries == 0) return 0; // record_frame then renders a normal frame; callers do not rely on that
    {
        const int v = ctx->rng_variant;
        const bool ok = v == 0 || (v == 1 && ctx->pointset_tables[2] && ctx->pointset_tables[3]) || (v == 2 && ctx->pointset_tables[0]) ||
                        (v == 3 && ctx->pointset_tables[0] && ctx->pointset_tables[1]);
        if (!ok) return fail(ctx, "rng_variant %d needs its tables: call rptr_cuda_set_pointset_table first", v);
    }
    CU(cudaSetDevice(ctx->device));
    FrameParams fp = make_frame_params(ctx);
    TileMap tm{};
    tm.width = ctx->width; tm.height = ctx->height;
    tm.rank = 0; tm.world = 1; tm.rows = 1;
    tm.local_rows = 0;
    tm.local_pixels = num_queries;
    // "dispatch ray queries into a virtual screen square" (record_frame, :3050-3056): ceil(sqrt(n)) invocations per row, in
    // workgroups of 32 x 16 (ComputeRenderPipelineVulkan::dispatch_rays)
    const int dim_x = (int)std::ceil(std::sqrt((float)num_queries));
    tm.query_wgs_x = (dim_x + 31) / 32;
    if (render_waves(ctx, fp, tm, fp.batch, 0u, ctx->rq_queries, ctx->rq_results)) return 1;
    CU(cudaGetLastError());
    return 0;
}

int rptr_cuda_normalize_options(rptr_ctx *ctx, rptr_backend_options *o, int32_t variant) {
    (void)variant;
    if (!ctx) return 1;
    if (!o) return fail(ctx, "normalize_options: options is NULL");
    if (o->rng_variant < 0 || o->rng_variant > RPTR_RNG_VARIANT_Z_SBL) o->rng_variant = RPTR_RNG_VARIANT_UNIFORM;
    if (o->light_sampling_variant < 0 || o->light_sampling_variant > RPTR_LIGHT_SAMPLING_VARIANT_RIS) o->light_sampling_variant = RPTR_LIGHT_SAMPLING_VARIANT_RIS;
    if (o->light_sampling_bucket_count < 1) o->light_sampling_bucket_count = 16;
    if (o->render_upscale_factor < 1) o->render_upscale_factor = 1;
    memset(o->_pad0, 0, sizeof(o->_pad0));
    memset(o->_pad1, 0, sizeof(o->_pad1));
    memset(o->_pad2, 0, sizeof(o->_pad2));
    return 0;
}

int rptr_cuda_configure_for(rptr_ctx *ctx, const rptr_backend_options *o, int32_t variant, rptr_backend_options *available) {
    if (!ctx) return 1;
    if (!o) return fail(ctx, "configure_for: options is NULL");
    if (ctx->in_frame) return fail(ctx, "configure_for inside begin_frame/end_frame");
    rptr_backend_options ok = *o;
    rptr_cuda_normalize_options(ctx, &ok, variant);
    std::string why;
    if (variant != 0) why += "variant index " + std::to_string(variant) + " does not exist (this backend has one integrator, PT_WAVEFRONT); ";
    if (o->rng_variant < 0 || o->rng_variant > RPTR_RNG_VARIANT_Z_SBL) why += "unknown rng_variant; ";
    if (o->light_sampling_variant != RPTR_LIGHT_SAMPLING_VARIANT_RIS) {
        why += "light_sampling_variant must be RIS (binned triangle lights + sun); ";
        ok.light_sampling_variant = RPTR_LIGHT_SAMPLING_VARIANT_RIS;
    }
    if (ok.render_upscale_factor > 8) {
        why += "render_upscale_factor " + std::to_string(o->render_upscale_factor) + " is out of range (1..8); ";
        ok.render_upscale_factor = 8;
    }
    if (o->enable_taa && !ctx->realtime_resolve) {
        why += "enable_taa needs option realtime_resolve (the reference compiles its TAA pass with ENABLE_REALTIME_RESOLVE only); ";
        ok.enable_taa = 0;
    }
    if (ok.rng_variant != RPTR_RNG_VARIANT_UNIFORM) {
        const int v = ok.rng_variant;
        const bool tables = (v == 1 && ctx->pointset_tables[2] && ctx->pointset_tables[3]) || (v == 2 && ctx->pointset_tables[0]) ||
                            (v == 3 && ctx->pointset_tables[0] && ctx->pointset_tables[1]);
        if (!tables) {
            why += "rng_variant " + std::to_string(v) + " needs its tables (rptr_cuda_set_pointset_table) first; ";
            ok.rng_variant = RPTR_RNG_VARIANT_UNIFORM;
        }
    }
    if (available) *available = ok;
    if (!why.empty()) return fail(ctx, "configure_for: %s", why.c_str());
    ctx->rng_variant = ok.rng_variant;
    ctx->upscale_option = ok.render_upscale_factor; // the LDR target is sized by the next initialize (render_vulkan.cpp:255-263; the app
                                                    // re-initialises when the factor changes, app.cpp:434-445)
    ctx->enable_taa = ok.enable_taa;
    // what the reference does at the end of a successful configure_for when built without ENABLE_REALTIME_RESOLVE
    // (render_vulkan.cpp:1911-1915) -- params.reprojection_mode = NONE -- is left to the caller's RenderParams here: without option
    // realtime_resolve the modes NONE and ACCUMULATE are the same running mean
    return 0;
}

// ---- multi-GPU: NCCL, loaded at run time ---------------------------------------------------------------------------------------
// The library does not link NCCL: a process that never shards a frame does not need it, and a host that already carries one
// (PyTorch bundles its own libnccl.so.2) must not get a second copy.  dlopen("libnccl.so.2") returns the copy the process has
// loaded, else the system's; RPTR_NCCL_LIB overrides the path.
namespace {
struct NcclId { char b[128]; }; // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed to ncclCommInitRank by value
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};
NcclApi g_nccl;
bool load_nccl() {
    if (g_nccl.lib) return true;
    const char *names[] = {std::getenv("RPTR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
        g_nccl.error = dlerror();
    }
    if (!g_nccl.lib) return false;
    bool ok = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(g_nccl.lib, name);
        if (!p) { ok = false; g_nccl.error = std::string("missing symbol ") + name; }
        return p;
    };
    *(void **)&g_nccl.GetUniqueId = sym("ncclGetUniqueId");
    *(void **)&g_nccl.CommInitRank = sym("ncclCommInitRank");
    *(void **)&g_nccl.CommInitAll = sym("ncclCommInitAll");
    *(void **)&g_nccl.CommDestroy = sym("ncclCommDestroy");
    *(void **)&g_nccl.Reduce = sym("ncclReduce");
    *(void **)&g_nccl.AllReduce = sym("ncclAllReduce");
    *(void **)&g_nccl.GroupStart = sym("ncclGroupStart");
    *(void **)&g_nccl.GroupEnd = sym("ncclGroupEnd");
    *(void **)&g_nccl.GetErrorString = sym("ncclGetErrorString");
    if (!ok) { dlclose(g_nccl.lib); g_nccl.lib = nullptr; }
    return ok;
}
const int kNcclFloat = 7, kNcclSum = 0; // ncclFloat32, ncclSum (nccl.h)
int attach_comm(rptr_ctx *ctx, void *comm, int world, int rank) {
    if (ctx->in_frame) return fail(ctx, "comm_init inside begin_frame/end_frame");
    if (ctx->comm) g_nccl.CommDestroy(ctx->comm);
    ctx->comm = comm;
    ctx->comm_world = world;
    ctx->comm_rank = rank;
    ctx->tile_world = world; // interleaved bands of tile_rows rows: band b belongs to rank b % world
    ctx->tile_rank = rank;
    return 0;
}
int enqueue_reduce(rptr_ctx *ctx, int root) {
    if (!ctx->accum) return fail(ctx, "reduce_framebuffer before initialize");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, "cudaSetDevice failed");
    const size_t n = (size_t)ctx->width * ctx->height;
    const bool receives = root < 0 || root == ctx->comm_rank;
    if (receives && !ctx->reduced && cudaMalloc((void **)&ctx->reduced, n * sizeof(float4)) != cudaSuccess) return fail(ctx, "out of memory for the reduced framebuffer");
    // ranks own disjoint bands and hold exact zeros elsewhere (initialize clears the accumulator, a rank only ever writes its
    // own pixels): the sum over ranks is a gather, exact in fp32 whatever the reduction order
    const int rc = root < 0 ? g_nccl.AllReduce(ctx->accum, ctx->reduced, 4 * n, kNcclFloat, kNcclSum, ctx->comm, ctx->stream)
                            : g_nccl.Reduce(ctx->accum, receives ? ctx->reduced : nullptr, 4 * n, kNcclFloat, kNcclSum, root, ctx->comm, ctx->stream);
    if (rc != 0) return fail(ctx, "NCCL reduce failed: %s", g_nccl.GetErrorString(rc));
    ctx->reduced_valid = receives;
    return 0;
}
} // namespace

int rptr_cuda_comm_unique_id(void *id, size_t bytes) {
    if (!id || bytes < 128) return fail(nullptr, "comm_unique_id: need a 128-byte buffer");
    if (!load_nccl()) return fail(nullptr, "cannot load NCCL (%s)", g_nccl.error.c_str());
    const int rc = g_nccl.GetUniqueId(id);
    if (rc != 0) return fail(nullptr, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
    return 0;
}

int rptr_cuda_comm_init_rank(rptr_ctx *ctx, int32_t world, int32_t rank, const void *id, size_t bytes) {
    if (!ctx) return 1;
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, "comm_init_rank: rank %d outside world %d", rank, world);
    if (!id || bytes < 128) return fail(ctx, "comm_init_rank: need the 128-byte id of rptr_cuda_comm_unique_id");
    if (!load_nccl()) return fail(ctx, "cannot load NCCL (%s)", g_nccl.error.c_str());
    CU(cudaSetDevice(ctx->device));
    NcclId nid;
    memcpy(nid.b, id, 128);
    void *comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, world, nid, rank);
    if (rc != 0) return fail(ctx, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
    return attach_comm(ctx, comm, world, rank);
}

int rptr_cuda_comm_init_all(rptr_ctx **ctxs, int32_t n) {
    if (!ctxs || n < 1) return 1;
    for (int i = 0; i < n; ++i)
        if (!ctxs[i]) return 1;
    if (!load_nccl()) return fail(ctxs[0], "cannot load NCCL (%s)", g_nccl.error.c_str());
    std::vector<int> devs(n);
    std::vector<void *> comms(n, nullptr);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    const int rc = g_nccl.CommInitAll(comms.data(), n, devs.data());
    if (rc != 0) return fail(ctxs[0], "ncclCommInitAll: %s", g_nccl.GetErrorString(rc));
    for (int i = 0; i < n; ++i)
        if (attach_comm(ctxs[i], comms[i], n, i)) return 1;
    return 0;
}

int rptr_cuda_comm_destroy(rptr_ctx *ctx) {
    if (!ctx) return 1;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->comm_world = 1;
    ctx->comm_rank = 0;
    ctx->reduced_valid = false;
    return 0;
}

int rptr_cuda_reduce_framebuffer(rptr_ctx *ctx, int32_t root) {
    if (!ctx) return 1;
    if (!ctx->comm) return fail(ctx, "reduce_framebuffer without a communicator (rptr_cuda_comm_init_rank / _all)");
    if (root >= ctx->comm_world) return fail(ctx, "reduce_framebuffer: root %d outside world %d", root, ctx->comm_world);
    return enqueue_reduce(ctx, root);
}

int rptr_cuda_reduce_framebuffer_all(rptr_ctx **ctxs, int32_t n, int32_t root) {
    if (!ctxs || n < 1 || !ctxs[0]) return 1;
    for (int i = 0; i < n; ++i)
        if (!ctxs[i] || !ctxs[i]->comm || ctxs[i]->comm_world != n) return fail(ctxs[0], "reduce_framebuffer_all: context %d is not part of a communicator of %d", i, n);
    g_nccl.GroupStart(); // one thread drives every device of the process
    int bad = 0;
    for (int i = 0; i < n; ++i) bad |= enqueue_reduce(ctxs[i], root);
    const int rc = g_nccl.GroupEnd();
    if (rc != 0) return fail(ctxs[0], "ncclGroupEnd: %s", g_nccl.GetErrorString(rc));
    return bad;
}

int rptr_write_pfm(const char *prefix, uint32_t width, uint32_t height, uint32_t channels, const float *pixels) {
    if (!prefix || width == 0 || height == 0 || channels < 3 || !pixels) return 1;
    std::string path = std::string(prefix) + ".pfm";
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return 1;
    fprintf(f, "PF\n%i %i\n-1.0\n", (int)width, (int)height);
    std::vector<float> row((size_t)width * 3);
    for (uint32_t y = 0; y < height; ++y) { // bottom row first
        const float *src = pixels + (size_t)(height - y - 1) * width * channels;
        for (uint32_t x = 0; x < width; ++x)
            for (uint32_t j = 0; j < 3; ++j) row[(size_t)x * 3 + j] = src[(size_t)x * channels + j];
        fwrite(row.data(), sizeof(float), row.size(), f);
    }
    fclose(f);
    return 0;
}

} // extern "C"
