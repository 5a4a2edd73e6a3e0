// rptr_bvh.cuh -- BVH node layout and the closest-hit / any-hit traversal routines of the trace stage.
//
// Replaces the driver-side black box of the reference (rayQueryInitializeEXT/ProceedEXT, vulkan/pt_megakernel.glsl:
// 440-478 and :225-264; acceleration structures from vulkan/vulkanrt_utils.cpp:82-167).  Closest-hit contract
// (DESIGN.md section 5): Moeller-Trumbore on world-space (v0, e1, e2) exactly as written in
// intersect_tri(); no culling; a hit needs tmin < t < tmax; ties in t go to the lowest flattened triangle id.
// Box tests are conservative (padded boxes + widened slab test) so culling never changes that result; they are free to
// use any arithmetic (fma slabs) because they only prune.
#pragma once
#include "rptr_shading.cuh"

namespace rp {

#define RPTR_EMPTY ((int32_t)0x80000000)
#define RPTR_BVH_WIDTH 4
#define RPTR_STACK_SIZE 128
#define RPTR_MAX_BVH_DEPTH 40 // builder guarantee: 3 pushes per level + 1 < RPTR_STACK_SIZE

// 128-byte four-wide node, child boxes as structure of arrays, read as 7 x 128-bit words (the 8th is padding):
//   w0 = lo.x[4]  w1 = lo.y[4]  w2 = lo.z[4]  w3 = hi.x[4]  w4 = hi.y[4]  w5 = hi.z[4]  w6 = child[4]
// child[k] >= 0: inner node index; < 0: leaf reference ~((first_triangle << 2) | (count - 1)), count in [1, 4];
// RPTR_EMPTY: unused slot (its box is inverted so it can never be hit).
struct alignas(128) BvhNode {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int32_t child[4];
    int32_t pad[4];
};
static_assert(sizeof(BvhNode) == 128, "BvhNode must be one 128-byte record");
static_assert(sizeof(Tri) == 48 || sizeof(Tri) == 64, "Tri must be three 128-bit or two 256-bit words");

RPTR_HD int32_t make_leaf_ref(int32_t first, int32_t count) { return ~((first << 2) | (count - 1)); }
RPTR_HD bool is_leaf_ref(int32_t r) { return r < 0 && r != RPTR_EMPTY; }

struct BvhDev {
    const BvhNode *nodes; // breadth-first order: node 0 = root
    const Tri *tris;      // leaf order
    int32_t n_nodes;
    int32_t n_tris;
    // the first top_k nodes again, with the eight 16-byte words of node i stored at word position w ^ (i & 7):
    // the image a CTA of the trace kernel copies into shared memory (the XOR spreads random nodes over the banks)
    const BvhNode *top_swizzled;
    int32_t top_k;
};

#ifndef RPTR_TOP_NODES_MAX
#define RPTR_TOP_NODES_MAX 1024 // 128 KB of shared memory per CTA
#endif

struct HitRec {
    float t, u, v;
    int32_t tri; // index into BvhDev::tris (leaf order), -1 = miss
    int32_t id;  // flattened id of that triangle
};

RPTR_HD float4 ld128(const void *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const float4 *>(p));
#else
    float4 r;
    memcpy(&r, p, 16);
    return r;
#endif
}
RPTR_HD int32_t f2i(float f) { return (int32_t)f2u(f); }

// The ray/triangle routine of the closest-hit contract (bit-exact part: no fma beyond dot/cross of RPTR-FP).
RPTR_HD bool intersect_tri(float3 v0, float3 e1, float3 e2, float3 o, float3 d, float &t, float &u, float &v) {
    float3 p = cross(d, e2);
    float det = dot(e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    float3 s = o - v0;
    u = dot(s, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    float3 q = cross(s, e1);
    v = dot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(e2, q) * inv;
    return true;
}

struct TraceCounters { uint32_t nodes, tris; };

RPTR_HD float slab_safe(float d) { return fabsf(d) > 1e-18f ? d : copysignf(1e-18f, d); }

// Slab test with fma: t = b * inv - o * inv.  Pruning only (conservative: boxes are padded at build time by more than
// the rounding of this expression for origins within ~16x the scene extent; tfar is widened by 4 ulp).
RPTR_HD bool slab(float lox, float loy, float loz, float hix, float hiy, float hiz, float3 inv, float3 ood, float tmin, float tmax,
                  float &tnear) {
    float t0 = fmaf(lox, inv.x, -ood.x), t1 = fmaf(hix, inv.x, -ood.x);
    float tn = fminf(t0, t1), tf = fmaxf(t0, t1);
    t0 = fmaf(loy, inv.y, -ood.y); t1 = fmaf(hiy, inv.y, -ood.y);
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    t0 = fmaf(loz, inv.z, -ood.z); t1 = fmaf(hiz, inv.z, -ood.z);
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    // [tn, tf] x [tmin, tmax] non-empty  <=>  max(tn, tmin) <= min(tf, tmax)
    tf = fminf(tf * 1.0000004f, tmax);
    tn = fmaxf(tn, tmin);
    tnear = tn;
    return tn <= tf;
}

RPTR_HD float comp4(const float4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// Reference traversal (host-executable statement of the contract; the GPU's persistent kernel in
// rptr_trace_kernels.cuh visits the same tree in a different order with the same result).  Any = stop at the first hit.
template <bool Any>
RPTR_HD bool trace_ray(const BvhDev &bvh, float3 o, float3 d, float tmin, float tmax, HitRec &best, TraceCounters &cnt) {
    best.tri = -1;
    best.id = 0x7fffffff;
    best.t = tmax;
    best.u = best.v = 0.0f;
    if (bvh.n_nodes == 0) return false;
    // 1/d for the fma slabs, with |d.k| clamped away from zero: an exactly axis-parallel ray (d.k == +-0, common for
    // sun shadow rays) would otherwise give lo*inf - o*inf = NaN on one side of the slab and a wrong rejection.
    const float3 inv = f3(1.0f / slab_safe(d.x), 1.0f / slab_safe(d.y), 1.0f / slab_safe(d.z));
    const float3 ood = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    int32_t stack[RPTR_STACK_SIZE];
    int sp = 0;
    int32_t cur = 0;
    for (;;) {
        if (cur >= 0) {
            const char *np = reinterpret_cast<const char *>(bvh.nodes + cur);
            const float4 w0 = ld128(np), w1 = ld128(np + 16), w2 = ld128(np + 32), w3 = ld128(np + 48), w4 = ld128(np + 64),
                         w5 = ld128(np + 80), w6 = ld128(np + 96);
            cnt.nodes++;
            // hit children, nearest first into `cur`, the others onto the stack
            int32_t near_ref = RPTR_EMPTY;
            float near_t = 0.0f;
            for (int k = 0; k < RPTR_BVH_WIDTH; ++k) {
                const int32_t ref = f2i(comp4(w6, k));
                float tn;
                if (ref == RPTR_EMPTY) continue;
                if (!slab(comp4(w0, k), comp4(w1, k), comp4(w2, k), comp4(w3, k), comp4(w4, k), comp4(w5, k), inv, ood, tmin, best.t, tn)) continue;
                if (near_ref == RPTR_EMPTY) {
                    near_ref = ref; near_t = tn;
                } else if (tn < near_t) {
                    stack[sp++] = near_ref;
                    near_ref = ref; near_t = tn;
                } else
                    stack[sp++] = ref;
            }
            cur = near_ref != RPTR_EMPTY ? near_ref : (sp > 0 ? stack[--sp] : RPTR_EMPTY);
        } else if (cur != RPTR_EMPTY) {
            const int32_t ref = ~cur;
            const int32_t first = ref >> 2, n = (ref & 3) + 1;
            for (int32_t i = 0; i < n; ++i) {
                const char *tp = reinterpret_cast<const char *>(bvh.tris + first + i);
                const float4 a = ld128(tp), b = ld128(tp + 16), c4 = ld128(tp + 32);
                cnt.tris++;
                float t, u, v;
                if (!intersect_tri(f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c4.x), o, d, t, u, v)) continue;
                if (!(t > tmin && t < tmax)) continue;
                const int32_t id = f2i(c4.y);
                if (Any) {
                    best.t = t; best.u = u; best.v = v; best.tri = first + i; best.id = id;
                    return true;
                }
                if (best.tri < 0 || t < best.t || (t == best.t && id < best.id)) {
                    best.t = t; best.u = u; best.v = v; best.tri = first + i; best.id = id;
                }
            }
            cur = sp > 0 ? stack[--sp] : RPTR_EMPTY;
        } else
            break;
    }
    return best.tri >= 0;
}

} // namespace rp
