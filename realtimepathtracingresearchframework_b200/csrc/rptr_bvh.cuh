// rptr_bvh.cuh -- BVH node layout and the closest-hit / any-hit traversal routines of the trace stage.
//
// Replaces the driver-side black box of the reference (rayQueryInitializeEXT/ProceedEXT, vulkan/pt_megakernel.glsl:
// 440-478 and :225-264; acceleration structures from vulkan/vulkanrt_utils.cpp:82-167).  Closest-hit contract
// (DESIGN.md section 5): Moeller-Trumbore on world-space (v0, e1, e2) exactly as written in
// intersect_tri(); no culling; a hit needs tmin < t < tmax; ties in t go to the lowest flattened triangle id.
// Box tests are conservative (padded boxes + widened slab test) so culling never changes that result.
#pragma once
#include "rptr_shading.cuh"

namespace rp {

// 64-byte two-child node (both child boxes in the parent: one fetch decides both descents)
struct BvhNode {
    float c0min[3], c0max[3], c1min[3], c1max[3];
    int32_t c0, c1; // >= 0: inner node index; < 0: leaf, first triangle = ~c
    int32_t n0, n1; // triangle count when the child is a leaf
};

struct BvhDev {
    const BvhNode *nodes;
    const Tri *tris; // leaf order
    int32_t n_nodes;
    int32_t n_tris;
};

struct HitRec {
    float t, u, v;
    int32_t tri; // index into BvhDev::tris (leaf order), -1 = miss
    int32_t id;  // flattened id of that triangle
};

RPTR_HD bool intersect_tri(const Tri &tr, float3 o, float3 d, float &t, float &u, float &v) {
    float3 e1 = f3(tr.e1x, tr.e1y, tr.e1z), e2 = f3(tr.e2x, tr.e2y, tr.e2z);
    float3 p = cross(d, e2);
    float det = dot(e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    float3 s = o - f3(tr.v0x, tr.v0y, tr.v0z);
    u = dot(s, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    float3 q = cross(s, e1);
    v = dot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(e2, q) * inv;
    return true;
}

RPTR_HD bool slab(const float *bmin, const float *bmax, float3 o, float3 inv, float tmin, float tmax, float &tnear) {
    float t0 = (bmin[0] - o.x) * inv.x, t1 = (bmax[0] - o.x) * inv.x;
    float tn = fminf(t0, t1), tf = fmaxf(t0, t1);
    t0 = (bmin[1] - o.y) * inv.y; t1 = (bmax[1] - o.y) * inv.y;
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    t0 = (bmin[2] - o.z) * inv.z; t1 = (bmax[2] - o.z) * inv.z;
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    tf *= 1.0000004f;
    tnear = tn;
    return tn <= tf && tf >= tmin && tn <= tmax;
}

struct TraceCounters { uint32_t nodes, tris; };

// generic traversal; Any = stop at the first accepted triangle
template <bool Any>
RPTR_HD bool trace_ray(const BvhDev &bvh, float3 o, float3 d, float tmin, float tmax, HitRec &best, TraceCounters &cnt) {
    best.tri = -1;
    best.id = 0x7fffffff;
    best.t = tmax;
    best.u = best.v = 0.0f;
    if (bvh.n_nodes == 0) return false;
    float3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int32_t stack[64];
    int sp = 0;
    int32_t cur = 0;
    for (;;) {
        const BvhNode nd = bvh.nodes[cur];
        cnt.nodes++;
        float tn0, tn1;
        bool h0 = slab(nd.c0min, nd.c0max, o, inv, tmin, best.t, tn0);
        bool h1 = nd.n1 >= 0 && slab(nd.c1min, nd.c1max, o, inv, tmin, best.t, tn1);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            bool hs = side == 0 ? h0 : h1;
            int32_t c = side == 0 ? nd.c0 : nd.c1;
            if (hs && c < 0) {
                int32_t first = ~c;
                int32_t n = side == 0 ? nd.n0 : nd.n1;
                for (int32_t i = 0; i < n; ++i) {
                    const Tri &tr = bvh.tris[first + i];
                    cnt.tris++;
                    float t, u, v;
                    if (!intersect_tri(tr, o, d, t, u, v)) continue;
                    if (!(t > tmin && t < tmax)) continue;
                    if (Any) {
                        best.t = t; best.u = u; best.v = v; best.tri = first + i; best.id = tr.id;
                        return true;
                    }
                    if (best.tri < 0 || t < best.t || (t == best.t && tr.id < best.id)) {
                        best.t = t; best.u = u; best.v = v; best.tri = first + i; best.id = tr.id;
                    }
                }
                if (side == 0) h0 = false;
                else h1 = false;
            }
        }
        if (h0 && h1) {
            bool near0 = tn0 <= tn1;
            stack[sp++] = near0 ? nd.c1 : nd.c0;
            cur = near0 ? nd.c0 : nd.c1;
        } else if (h0) {
            cur = nd.c0;
        } else if (h1) {
            cur = nd.c1;
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    return best.tri >= 0;
}

} // namespace rp
