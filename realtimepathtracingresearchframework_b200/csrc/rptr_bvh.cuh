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

// 64-byte two-child node (both child boxes in the parent: one fetch decides both descents), read as 4 x 128-bit words:
//   q0 = (c0min.xyz, c0max.x)  q1 = (c0max.yz, c1min.xy)  q2 = (c1min.z, c1max.xyz)  q3 = (c0, c1, n0, n1)
struct alignas(64) BvhNode {
    float c0min[3], c0max[3], c1min[3], c1max[3];
    int32_t c0, c1; // >= 0: inner node index; < 0: leaf, first triangle = ~c
    int32_t n0, n1; // triangle count when the child is a leaf, 0 for an inner child, -1 for "no child"
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be one 64-byte record");
static_assert(sizeof(Tri) == 48, "Tri must be three 128-bit words");

struct BvhDev {
    const BvhNode *nodes; // breadth-first order: node 0 = root
    const Tri *tris;      // leaf order
    int32_t n_nodes;
    int32_t n_tris;
    // the first top_k nodes again, with the four 16-byte words of node i stored at word position w ^ ((i >> 1) & 3):
    // the image a CTA of the trace kernel copies into shared memory (the XOR spreads random nodes over the banks)
    const BvhNode *top_swizzled;
    int32_t top_k;
};

#define RPTR_TOP_NODES_MAX 2048 // 128 KB of shared memory per CTA

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
    tf *= 1.0000004f;
    tnear = tn;
    return tn <= tf && tf >= tmin && tn <= tmax;
}

// generic traversal; Any = stop at the first accepted triangle
template <bool Any>
RPTR_HD bool trace_ray(const BvhDev &bvh, float3 o, float3 d, float tmin, float tmax, HitRec &best, TraceCounters &cnt) {
    best.tri = -1;
    best.id = 0x7fffffff;
    best.t = tmax;
    best.u = best.v = 0.0f;
    if (bvh.n_nodes == 0) return false;
    // 1/d for the fma slabs, with |d.k| clamped away from zero: an exactly axis-parallel ray (d.k == +-0, common for
    // sun shadow rays) would otherwise give lo*inf - o*inf = NaN on one side of the slab and a wrong rejection.  With
    // the clamp the slab interval of such an axis is (-huge, +huge) inside the slab and empty outside, as it should be.
    const float3 inv = f3(1.0f / slab_safe(d.x), 1.0f / slab_safe(d.y), 1.0f / slab_safe(d.z));
    const float3 ood = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    int32_t stack[64];
    int sp = 0;
    int32_t cur = 0;
    for (;;) {
        const char *np = reinterpret_cast<const char *>(bvh.nodes + cur);
        const float4 q0 = ld128(np), q1 = ld128(np + 16), q2 = ld128(np + 32), q3 = ld128(np + 48);
        cnt.nodes++;
        const int32_t c0 = f2i(q3.x), c1 = f2i(q3.y), n0 = f2i(q3.z), n1 = f2i(q3.w);
        float tn0, tn1 = 0.0f;
        bool h0 = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, inv, ood, tmin, best.t, tn0);
        bool h1 = n1 >= 0 && slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, inv, ood, tmin, best.t, tn1);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const bool hs = side == 0 ? h0 : h1;
            const int32_t c = side == 0 ? c0 : c1;
            if (hs && c < 0) {
                const int32_t first = ~c;
                const int32_t n = side == 0 ? n0 : n1;
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
                if (side == 0) h0 = false;
                else h1 = false;
            }
        }
        if (h0 && h1) {
            const bool near0 = tn0 <= tn1;
            stack[sp++] = near0 ? c1 : c0;
            cur = near0 ? c0 : c1;
        } else if (h0) {
            cur = c0;
        } else if (h1) {
            cur = c1;
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    return best.tri >= 0;
}

} // namespace rp
