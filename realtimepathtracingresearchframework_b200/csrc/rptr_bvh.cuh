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

#define RPTR_EMPTY ((int32_t)0x80000000) // "no value" marker of the traversal kernels
#define RPTR_BVH_WIDTH 8
#define RPTR_MAX_BVH_DEPTH 32 // builder guarantee (wide levels); the reference traversal pushes <= 7 entries per level
#define RPTR_STACK_SIZE (7 * RPTR_MAX_BVH_DEPTH + 8)

// 96-byte eight-wide node with quantised child boxes (six 128-bit words; three 256-bit loads from global memory):
//   w0 = org.x  org.y  org.z  ext.x
//   w1 = ext.y  ext.z  child_base  tri_base
//   w2 = qlo.x[0..3] qlo.x[4..7] qlo.y[0..3] qlo.y[4..7]
//   w3 = qlo.z[0..3] qlo.z[4..7] qhi.x[0..3] qhi.x[4..7]
//   w4 = qhi.y[0..3] qhi.y[4..7] qhi.z[0..3] qhi.z[4..7]
//   w5 = masks  pad  pad  pad
// Slot k (0..7) is an inner child (bit k of imask = masks & 0xff), ONE triangle (bit k of lmask = masks >> 8 & 0xff) or
// empty.  The inner children of a node are consecutive nodes starting at child_base, its triangles consecutive records of the
// leaf-order triangle array starting at tri_base, both in slot order: slot k -> child_base + popc(imask & ((1 << k) - 1)),
// tri_base + popc(lmask & ((1 << k) - 1)).  So a node step produces two 8-bit hit masks instead of sorted references, and the
// traversal stack holds (base, masks) groups -- one entry per node instead of one per child.
// Slots are assigned by the builder so that slot bit a (a = 0, 1, 2 for x, y, z) says on which side of the node's centre the
// child lies along axis a: a ray visits the inner hits in the order of descending (slot ^ octant), octant bit a = (d.a >= 0),
// which is front to back for children that tile the node (Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs
// Through Compressed Wide BVHs", HPG 2017 -- the idea; layout and arithmetic below are ours).
// Child k spans, per axis, [org + f(qlo[k]) * ext, org + f(qhi[k]) * ext] with f(b) = 1 + (b & 127) / 128: every byte is
// stored as 0x80 | q (q in [0, 127]) so that ONE byte permute builds the float 0x3f000000 | b << 16 = f(b) in [1, 2),
// and the slab test is t = fma(f, ext * (1/d), (org - o) * (1/d)) -- no integer-to-float conversion, no cancellation
// against a large bias.  encode_node() rounds lo down / hi up in exact arithmetic, so the decoded boxes contain the builder's
// padded boxes and culling stays conservative (box tests only prune, the closest-hit contract is untouched).
struct alignas(32) BvhNode {
    float org[3];
    float ext[3];
    int32_t child_base;
    int32_t tri_base;
    uint32_t qlo[3][2]; // byte (k & 3) of qlo[a][k >> 2] = 0x80 | quantised lower bound of slot k on axis a
    uint32_t qhi[3][2];
    uint32_t masks;     // imask | lmask << 8
    uint32_t pad[3];
};
static_assert(sizeof(BvhNode) == 96, "BvhNode must be three 256-bit words");
static_assert(sizeof(Tri) == 48 || sizeof(Tri) == 64, "Tri must be three 128-bit or two 256-bit words");
#define RPTR_NODE_WORDS 6 // 16-byte words per node

RPTR_HD uint32_t node_imask(const BvhNode &n) { return n.masks & 0xffu; }
RPTR_HD uint32_t node_lmask(const BvhNode &n) { return (n.masks >> 8) & 0xffu; }
RPTR_HD int popc8(uint32_t m) {
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

struct BvhDev {
    const BvhNode *nodes; // node 0 = root; the inner children of a node are consecutive
    const Tri *tris;      // leaf order: the triangles of a node are consecutive
    int32_t n_nodes;
    int32_t n_tris;
    // the first top_k nodes again, split into six planes of 16-byte words (plane w at byte w * 16 * RPTR_TOP_NODES_MAX
    // holds word w of node 0, 1, ...): the image a CTA of the trace kernel copies into shared memory.  A warp reading
    // word w of 32 random nodes then spreads over all banks, and the six addresses differ by immediates.
    const float4 *top_planes;
    int32_t top_k;
};

#ifndef RPTR_TOP_NODES_MAX
#define RPTR_TOP_NODES_MAX 832 // 78 KB of shared memory per CTA
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

// ---- stochastic alpha: the candidate filter of non-opaque triangles (generate_candidate_hit, pt_megakernel.glsl:153-211) ----
// Only triangles whose material alpha texel is not 255 take part (tri_alpha8).  Closest hit: candidates are
// judged front to back in (t, id) order with the path's own LCG -- rejected iff !(alpha > 0) || (alpha < 1 && u > alpha), one
// draw per candidate with 0 < alpha < 1 -- which the traversal realises as "closest hit AFTER (after_t, after_id)" restarted
// for every rejected candidate.  Visibility rays (:216-272): every candidate gets its own LCG seeded from
// (primitive ^ frame_id, instance ^ frame_offset, pixel), so the verdict does not depend on the traversal order.
struct AlphaFilter {
    SceneDev scene;                  // geometry instances (per-candidate seeds), materials + textures (alpha of textured candidates)
    uint32_t frame_id, frame_offset; // view_params.frame_id / frame_offset of the frame (batch)
    uint32_t pixel_linear;           // gl_GlobalInvocationID.x + y * frame_dims.x
};
RPTR_HD bool alpha_rejects(float alpha, uint32_t &lcg_state) { return !(alpha > 0.0f) || (alpha < 1.0f && lcg_randomf(lcg_state) > alpha); }
// (u, v) = barycentrics of the candidate: only read when its alpha comes from a texture larger than 1 x 1
RPTR_HD bool shadow_candidate_passes(const AlphaFilter &f, int32_t gi_alpha, int32_t prim, float u, float v) {
    const int32_t a8 = (int32_t)(((uint32_t)gi_alpha) >> 24);
    if (a8 == RPTR_TRI_OPAQUE && !(gi_alpha & RPTR_TRI_TEXTURED_ALPHA)) return true;
    const GeomInst &g = f.scene.ginst[gi_alpha & 0x007fffff];
    uint32_t st = lcg_seed((uint32_t)prim ^ f.frame_id, (uint32_t)g.instance ^ f.frame_offset, f.pixel_linear);
    return !alpha_rejects(candidate_alpha(f.scene, gi_alpha, prim, u, v), st);
}

RPTR_HD float slab_safe(float d) { return fabsf(d) > 1e-18f ? d : copysignf(1e-18f, d); }

RPTR_HD float u2f_(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
// f(b) of byte k of a packed word: float bits 0x3f000000 | byte << 16 (PRMT on the device)
RPTR_HD float qfloat(uint32_t word, int k) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(word, 0x3f000000u, 0x7044u | ((uint32_t)k << 8)));
#else
    return u2f_(0x3f000000u | (((word >> (8 * k)) & 0xffu) << 16));
#endif
}

// Per-node part of the slab test: a = ext / d, b = (org - o) / d  (fma form, |error| ~ 2^-22 of the scene scale, covered
// by the padding of the builder's boxes for ray origins within 8 scene extents, which begin_frame / trace_rays enforce).
struct NodeSlab { float ax, ay, az, bx, by, bz; };
RPTR_HD NodeSlab node_slab(float orgx, float orgy, float orgz, float extx, float exty, float extz, float3 inv, float3 ood) {
    NodeSlab n;
    n.ax = extx * inv.x; n.ay = exty * inv.y; n.az = extz * inv.z;
    n.bx = fmaf(orgx, inv.x, -ood.x); n.by = fmaf(orgy, inv.y, -ood.y); n.bz = fmaf(orgz, inv.z, -ood.z);
    return n;
}
// Slab test of slot k (pruning only; tfar is widened by 4 ulp).
RPTR_HD bool slab_q(const NodeSlab &n, const BvhNode &nd, int k, float tmin, float tmax) {
    const int w = k >> 2, j = k & 3;
    float t0 = fmaf(qfloat(nd.qlo[0][w], j), n.ax, n.bx), t1 = fmaf(qfloat(nd.qhi[0][w], j), n.ax, n.bx);
    float tn = fminf(t0, t1), tf = fmaxf(t0, t1);
    t0 = fmaf(qfloat(nd.qlo[1][w], j), n.ay, n.by); t1 = fmaf(qfloat(nd.qhi[1][w], j), n.ay, n.by);
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    t0 = fmaf(qfloat(nd.qlo[2][w], j), n.az, n.bz); t1 = fmaf(qfloat(nd.qhi[2][w], j), n.az, n.bz);
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    // [tn, tf] x [tmin, tmax] non-empty  <=>  max(tn, tmin) <= min(tf, tmax)
    tf = fminf(tf * 1.0000004f, tmax);
    tn = fmaxf(tn, tmin);
    return tn <= tf;
}

// Quantise the boxes of the used slots (lo[k][axis], hi[k][axis]; already padded by the builder) into one node.  kind[k]:
// 0 empty, 1 inner child, 2 triangle.  Exact arithmetic in double (all operands are floats with <= 8 extra bits): decoded
// lower bounds never exceed lo, upper bounds never fall short of hi.
RPTR_HD BvhNode encode_node(const float (*lo)[3], const float (*hi)[3], const int *kind, int32_t child_base, int32_t tri_base) {
    BvhNode nd;
    nd.child_base = child_base;
    nd.tri_base = tri_base;
    nd.pad[0] = nd.pad[1] = nd.pad[2] = 0u;
    uint32_t imask = 0, lmask = 0;
    for (int k = 0; k < RPTR_BVH_WIDTH; ++k) {
        if (kind[k] == 1) imask |= 1u << k;
        if (kind[k] == 2) lmask |= 1u << k;
    }
    nd.masks = imask | (lmask << 8);
    for (int a = 0; a < 3; ++a) {
        double L = 1e300, H = -1e300;
        for (int k = 0; k < RPTR_BVH_WIDTH; ++k) {
            if (kind[k] == 0) continue;
            L = (double)lo[k][a] < L ? (double)lo[k][a] : L;
            H = (double)hi[k][a] > H ? (double)hi[k][a] : H;
        }
        if (L > H) { L = 0.0; H = 0.0; }
        // ext: 127 steps of ext/128 must cover [L, H] from a grid origin org + ext <= L, with slack for the rounding of org
        float ext = (float)((H - L) * (128.0 / 127.0) * 1.000001);
        const float tiny = (float)(fabs(L) + fabs(H)) * 1.1920929e-07f + 1e-30f;
        if (!(ext > tiny)) ext = tiny;
        float org;
        for (;;) {
            org = (float)(L - (double)ext);
            if ((double)org + (double)ext > L) org = nextafterf(org, -3.0e38f);
            if ((double)org + (double)ext * (255.0 / 128.0) >= H) break;
            ext = ext * 1.0009765625f;
        }
        nd.org[a] = org;
        nd.ext[a] = ext;
        uint32_t wlo[2] = {0x80808080u, 0x80808080u}, whi[2] = {0x80808080u, 0x80808080u};
        for (int k = 0; k < RPTR_BVH_WIDTH; ++k) {
            if (kind[k] == 0) continue;
            const double l = (double)lo[k][a], h = (double)hi[k][a], o = (double)org, e = (double)ext;
            int ql = (int)floor((l - o - e) / e * 128.0), qh = (int)ceil((h - o - e) / e * 128.0);
            ql = ql < 0 ? 0 : (ql > 127 ? 127 : ql);
            qh = qh < 0 ? 0 : (qh > 127 ? 127 : qh);
            while (ql > 0 && o + e * ((128.0 + ql) / 128.0) > l) --ql;
            while (qh < 127 && o + e * ((128.0 + qh) / 128.0) < h) ++qh;
            const int sh = 8 * (k & 3);
            wlo[k >> 2] = (wlo[k >> 2] & ~(0xffu << sh)) | ((0x80u | (uint32_t)ql) << sh);
            whi[k >> 2] = (whi[k >> 2] & ~(0xffu << sh)) | ((0x80u | (uint32_t)qh) << sh);
        }
        nd.qlo[a][0] = wlo[0]; nd.qlo[a][1] = wlo[1];
        nd.qhi[a][0] = whi[0]; nd.qhi[a][1] = whi[1];
    }
    return nd;
}

// Slot assignment of a wide node: slot bit a says on which side of the node's centre a child lies along axis a, so that
// (slot ^ ray octant) orders the children front to back.  Greedy: children far from the centre choose first (ties: lower
// index), each takes the free slot whose bits disagree least with its side of the centre, disagreement weighted by the
// distance to the plane (ties: lower slot).  Shared by the host and the device builder.
RPTR_HD void assign_slots(int nk, const float (*clo)[3], const float (*chi)[3], int *slot_of) {
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int k = 0; k < nk; ++k)
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], clo[k][a]); hi[a] = fmaxf(hi[a], chi[k][a]); }
    float off[RPTR_BVH_WIDTH][3], weight[RPTR_BVH_WIDTH];
    bool placed[RPTR_BVH_WIDTH], used[RPTR_BVH_WIDTH];
    for (int k = 0; k < RPTR_BVH_WIDTH; ++k) { placed[k] = false; used[k] = false; }
    for (int k = 0; k < nk; ++k) {
        weight[k] = 0.0f;
        for (int a = 0; a < 3; ++a) {
            off[k][a] = 0.5f * (clo[k][a] + chi[k][a]) - 0.5f * (lo[a] + hi[a]);
            weight[k] = fmaxf(weight[k], fabsf(off[k][a]));
        }
    }
    for (int i = 0; i < nk; ++i) {
        int k = -1;
        for (int j = 0; j < nk; ++j)
            if (!placed[j] && (k < 0 || weight[j] > weight[k])) k = j;
        placed[k] = true;
        int best = -1;
        float best_cost = 1e30f;
        for (int sl = 0; sl < RPTR_BVH_WIDTH; ++sl) {
            if (used[sl]) continue;
            float cost = 0.0f;
            for (int a = 0; a < 3; ++a) {
                const bool high = off[k][a] > 0.0f;
                if (high != (((sl >> a) & 1) != 0)) cost += fabsf(off[k][a]);
            }
            if (cost < best_cost) { best_cost = cost; best = sl; }
        }
        used[best] = true;
        slot_of[k] = best;
    }
}

// Decoded box of slot k (tests / validation).
RPTR_HD void decode_child(const BvhNode &nd, int k, float *lo, float *hi) {
    for (int a = 0; a < 3; ++a) {
        lo[a] = (float)((double)nd.org[a] + (double)qfloat(nd.qlo[a][k >> 2], k & 3) * (double)nd.ext[a]);
        hi[a] = (float)((double)nd.org[a] + (double)qfloat(nd.qhi[a][k >> 2], k & 3) * (double)nd.ext[a]);
    }
}

// octant of a ray for the slot order of a node: bit a set when d.a >= 0 (the near side is then the low side, slot bit 0)
RPTR_HD uint32_t ray_octant(float3 d) { return (d.x >= 0.0f ? 1u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 4u : 0u); }

// Reference traversal (host-executable statement of the contract; the GPU's persistent kernel in
// rptr_trace_kernels.cuh visits the same tree with group entries and the same result).  Any = stop at the first hit.
// Closest: only candidates strictly after (after_t, after_id) in (t, id) order count (after_t = tmin, after_id = INT_MAX: all).
// Any: `filter` (may be null = every triangle opaque) decides whether a non-opaque candidate occludes.
template <bool Any>
RPTR_HD bool trace_ray(const BvhDev &bvh, float3 o, float3 d, float tmin, float tmax, HitRec &best, TraceCounters &cnt, float after_t,
                      int32_t after_id, const AlphaFilter *filter) {
    best.tri = -1;
    best.id = 0x7fffffff;
    best.t = tmax;
    best.u = best.v = 0.0f;
    if (bvh.n_nodes == 0) return false;
    // 1/d for the fma slabs, with |d.k| clamped away from zero: an exactly axis-parallel ray (d.k == +-0, common for
    // sun shadow rays) would otherwise give lo*inf - o*inf = NaN on one side of the slab and a wrong rejection.
    const float3 inv = f3(1.0f / slab_safe(d.x), 1.0f / slab_safe(d.y), 1.0f / slab_safe(d.z));
    const float3 ood = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    const uint32_t oct = ray_octant(d);
    int32_t stack[RPTR_STACK_SIZE];
    int sp = 0;
    int32_t cur = 0;
    for (;;) {
        const BvhNode &nd = bvh.nodes[cur];
        cnt.nodes++;
        const NodeSlab ns = node_slab(nd.org[0], nd.org[1], nd.org[2], nd.ext[0], nd.ext[1], nd.ext[2], inv, ood);
        const uint32_t imask = node_imask(nd), lmask = node_lmask(nd);
        // triangles of this node first (they shorten the ray for the inner children)
        for (int k = 0; k < RPTR_BVH_WIDTH; ++k) {
            if (!((lmask >> k) & 1u) || !slab_q(ns, nd, k, tmin, best.t)) continue;
            const int32_t ti = nd.tri_base + popc8(lmask & ((1u << k) - 1u));
            const char *tp = reinterpret_cast<const char *>(bvh.tris + ti);
            const float4 a = ld128(tp), b = ld128(tp + 16), c4 = ld128(tp + 32);
            cnt.tris++;
            float t, u, v;
            if (!intersect_tri(f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c4.x), o, d, t, u, v)) continue;
            if (!(t > tmin && t < tmax)) continue;
            const int32_t id = f2i(c4.y);
            if (Any) {
                if (filter && !shadow_candidate_passes(*filter, f2i(c4.z), f2i(c4.w), u, v)) continue;
                best.t = t; best.u = u; best.v = v; best.tri = ti; best.id = id;
                return true;
            }
            if (!(t > after_t || (t == after_t && id > after_id))) continue;
            if (best.tri < 0 || t < best.t || (t == best.t && id < best.id)) {
                best.t = t; best.u = u; best.v = v; best.tri = ti; best.id = id;
            }
        }
        // inner children: pushed in ascending priority (slot ^ oct), so the pop order is front to back
        for (uint32_t p = 0; p < RPTR_BVH_WIDTH; ++p) {
            const int k = (int)(p ^ oct);
            if (!((imask >> k) & 1u) || !slab_q(ns, nd, k, tmin, best.t)) continue;
            stack[sp++] = nd.child_base + popc8(imask & ((1u << k) - 1u));
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return best.tri >= 0;
}
template <bool Any>
RPTR_HD bool trace_ray(const BvhDev &bvh, float3 o, float3 d, float tmin, float tmax, HitRec &best, TraceCounters &cnt) {
    return trace_ray<Any>(bvh, o, d, tmin, tmax, best, cnt, tmin, 0x7fffffff, nullptr);
}

// Closest hit with the front-to-back candidate filter; lcg_state is the path's LCG (alpha_rng == rng for the LCG pointset,
// pt_megakernel.glsl:354-355) and advances by one draw per candidate with 0 < alpha < 1.
RPTR_HD bool closest_hit_filtered(const BvhDev &bvh, const SceneDev &sc, float3 o, float3 d, float tmin, float tmax, uint32_t &lcg_state, HitRec &best,
                                 TraceCounters &cnt) {
    float after_t = tmin;
    int32_t after_id = 0x7fffffff;
    for (;;) {
        if (!trace_ray<false>(bvh, o, d, tmin, tmax, best, cnt, after_t, after_id, nullptr)) return false;
        const Tri &tr = bvh.tris[best.tri];
        if (tri_alpha8(tr) == RPTR_TRI_OPAQUE && !(tr.gi_alpha & RPTR_TRI_TEXTURED_ALPHA)) return true;
        if (!alpha_rejects(candidate_alpha(sc, tr.gi_alpha, tr.prim, best.u, best.v), lcg_state)) return true;
        after_t = best.t;
        after_id = best.id;
    }
}

} // namespace rp
