// rptr_trace_tail.cuh -- the tail of a trace launch: the last rays of a queue, eight lanes per ray.
//
// A launch of the persistent kernel (rptr_trace_kernels.cuh) cannot end before its longest ray has, and in the per-lane state
// machine that ray advances one node at a time at the latency of a lone warp (~0.9 us per node step: a dependent L2 fetch plus
// several hundred instructions issued back to back) while 31 lanes idle -- a floor of 0.23-0.29 ms per launch, 17 launches per
// frame, which is what caps strong scaling over GPUs (profiles/r02_sweeps.md).  So a warp of the persistent kernel that finds the
// queue drained and is down to its last RPTR_TAIL_LIVE rays hands them over instead of finishing them: it appends a record per
// ray -- slot, best hit so far, alpha-filter state, and the (base, masks) groups the lane still had pending: its node-group stack,
// its triangle backlog, its two current groups -- to the launch's tail list and exits.  k_trace_tail runs right behind it on the
// same stream and gives every record eight lanes (four rays per warp, records claimed from a cursor).  The ray RESUMES: the
// lanes first take the record's groups (the pending children of the node groups go onto the ray's frontier in shared memory,
// the pending triangles are tested), then walk on breadth-wise -- every step each lane takes one node off the frontier (eight
// nodes per step instead of one), tests its eight slots with the node step of the persistent kernel, the inner children that
// were hit go back onto the frontier (prefix sum over the group), the triangle slots that were hit are intersected by the lane
// that found them, and one arg-min by (t, id) over the group per step shortens the ray.  Same box tests, same intersect_tri,
// same tie-break, and the closest-hit / any-hit result does not depend on the order of the walk (culling only prunes):
// bit-identical images with the tail kernel on or off (tests/test_gpu_parity.py).
// Versions measured on the way (profiles/r02_sweeps.md): four nodes per step with eight lanes per NODE (no faster than the state
// machine it relieved); one warp per ray, 32 nodes per step (12-14 lanes active: most steps of a ray have fewer than a dozen
// nodes on the frontier); eight lanes per ray restarting at the root (every handed-over ray paid its whole walk again).
#pragma once
#include "rptr_trace_kernels.cuh"

namespace rp {

#ifndef RPTR_TAIL_GROUP
#define RPTR_TAIL_GROUP 8         // lanes per ray
#endif
#define RPTR_TAIL_WARPS (RPTR_TAIL_GROUP >= 8 ? 4 : 2) // warps per CTA of the tail kernel (32 KB of frontiers)
#define RPTR_TAIL_FRONTIER 512    // node indices per ray, shared memory
#define RPTR_TAIL_RESERVE (7 * (RPTR_MAX_BVH_DEPTH + 2) + 8 * RPTR_TAIL_GROUP)
// records one launch can hand over: every warp of the persistent grid, RPTR_TAIL_LIVE rays each
#define RPTR_TAIL_CAPACITY(num_sms) ((size_t)(num_sms) * (RPTR_TRACE_THREADS / 32) * RPTR_TAIL_LIVE)

#if defined(__CUDACC__)

// priority permutation of an 8-bit hit mask (the LUT of the persistent kernel, computed; an involution): bit (s ^ oct) <-> bit s
__device__ __forceinline__ uint32_t order_hits(uint32_t m, uint32_t oct) {
    if (oct & 1u) m = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u);
    if (oct & 2u) m = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u);
    if (oct & 4u) m = ((m & 0x0fu) << 4) | ((m >> 4) & 0x0fu);
    return m;
}

template <bool Any, bool Alpha>
__global__ void __launch_bounds__(RPTR_TAIL_WARPS * 32) k_trace_tail(BvhDev bvh, TraceIO io, unsigned long long *c_nodes, unsigned long long *c_tris) {
    constexpr int G = RPTR_TAIL_GROUP, NG = 32 / G;
    __shared__ int32_t s_frontier[RPTR_TAIL_WARPS][NG][RPTR_TAIL_FRONTIER];
    __shared__ uint32_t hc_word;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & (G - 1), grp = lane / G;
    const unsigned gmask = ((1u << G) - 1u) << (grp * G); // the lanes of this lane's group
    const uint32_t n = *io.tail_count;
    if (n == 0u) return;
    if (threadIdx.x == 0) hc_word = 0x3f000000u;
    __syncthreads();
    const uint32_t hc = hc_word; // read back on purpose: see the persistent kernel (the PRMT selectors stay immediates)
    int32_t *frontier = s_frontier[warp][grp];
    unsigned long long n_nodes = 0, n_tris = 0;
    // the group's ray (the same values in its G lanes)
    bool active = false, exhausted = false;
    uint32_t slot = 0u, pixel_linear = 0u;
    float3 o = f3(0.0f), d = f3(0.0f), inv = f3(0.0f), ood = f3(0.0f);
    float tmin = 0.0f, tmax = 0.0f, best_t = 0.0f, best_u = 0.0f, best_v = 0.0f;
    int32_t after_id = 0x7fffffff, best_tri = -1, best_id = 0x7fffffff;
    int32_t cs = 0; // nodes on the group's frontier
    // a ray that was handed over with its pending groups resumes: while resume_at < resume_n the lanes take record entries
    // (node groups -> their pending children go onto the frontier; triangle groups -> their pending triangles are tested)
    // instead of nodes
    const TailRec *resume_rec = nullptr;
    int32_t resume_at = 0, resume_n = 0, resume_nodes = 0;
    for (;;) {
        // ---- groups without a ray take the next record of the tail list ----
        if (__any_sync(FULL, !active && !exhausted)) {
            uint32_t rec = 0xffffffffu;
            if (!active && !exhausted && sub == 0) rec = atomicAdd(io.tail_cursor, 1u);
            rec = __shfl_sync(FULL, rec, grp * G);
            if (!active && !exhausted) {
                if (rec >= n) exhausted = true;
                else {
                    const TailRec tr = io.tail[rec];
                    slot = tr.slot;
                    const float4 ro = io.ray_o[slot], rd = io.ray_d[slot];
                    o = f3(ro.x, ro.y, ro.z); d = f3(rd.x, rd.y, rd.z);
                    inv = f3(slab_rcp(slab_safe(d.x)), slab_rcp(slab_safe(d.y)), slab_rcp(slab_safe(d.z)));
                    ood = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
                    tmin = tr.tmin; tmax = rd.w; after_id = tr.after_id;
                    best_t = tr.best_t; best_u = tr.best_u; best_v = tr.best_v; best_tri = tr.best_tri; best_id = tr.best_id;
                    if (Alpha && Any) pixel_linear = tile_pixel_linear(io.tm, __float_as_uint(io.sh_c[slot].w) % (uint32_t)io.tm.local_pixels);
                    if (tr.n_node_groups < 0) { // not recorded: from the root
                        cs = bvh.n_nodes > 0 ? 1 : 0;
                        if (sub == 0) frontier[0] = 0;
                        resume_n = 0;
                    } else {
                        cs = 0;
                        resume_rec = io.tail + rec;
                        resume_nodes = tr.n_node_groups;
                        resume_n = tr.n_node_groups + tr.n_tri_groups;
                    }
                    resume_at = 0;
                    active = true;
                }
            }
        }
        if (!__any_sync(FULL, active)) break;
        __syncwarp(); // pushes of the previous step before the reads below
        // ---- one step: every lane of a group takes one node off the top of the group's frontier (one lane only while the frontier
        //      is nearly full: a depth-first walk adds at most seven entries per level of the tree, and that much is kept in reserve) ----
        const bool resuming = active && resume_at < resume_n;
        const int width = cs > RPTR_TAIL_FRONTIER - RPTR_TAIL_RESERVE ? 1 : G;
        const int take = (!active || resuming) ? 0 : (cs < width ? cs : width);
        const int32_t node = sub < take ? frontier[cs - 1 - sub] : -1;
        cs -= take;
        __syncwarp();
        if (sub == 0) n_nodes += (unsigned long long)take;
        uint32_t ih = 0u, th = 0u, imask = 0u, lmask = 0u;
        int32_t child_base = 0, tri_base = 0;
        if (resuming) { // lane `sub` takes entry resume_at + sub of the record
            const int32_t e = resume_at + sub;
            if (e < resume_n) {
                const uint32_t ex = resume_rec->gx[e], ey = resume_rec->gy[e];
                if (e < resume_nodes) { // pending inner children, in priority order for closest-hit rays
                    const uint32_t oct = Any ? 0u : ray_octant(d);
                    imask = (ey >> 8) & 0xffu; ih = Any ? (ey & 0xffu) : order_hits(ey & 0xffu, oct);
                    child_base = (int32_t)ex;
                } else { // pending triangles
                    lmask = (ey >> 8) & 0xffu; th = ey & 0xffu;
                    tri_base = (int32_t)ex;
                }
            }
            resume_at += G;
        }
        if (node >= 0) {
            const bool sx = inv.x < 0.0f, sy = inv.y < 0.0f, sz = inv.z < 0.0f;
            const char *np = reinterpret_cast<const char *>(bvh.nodes + node);
            float4 w0, w1, w2, w3, w4, w5;
            ld256(np, w0, w1);
            ld256(np + 32, w2, w3);
            ld256(np + 64, w4, w5);
            const NodeSlab ns = node_slab(w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, inv, ood);
            const uint32_t lx0 = __float_as_uint(w2.x), lx1 = __float_as_uint(w2.y), ly0 = __float_as_uint(w2.z), ly1 = __float_as_uint(w2.w);
            const uint32_t lz0 = __float_as_uint(w3.x), lz1 = __float_as_uint(w3.y), hx0 = __float_as_uint(w3.z), hx1 = __float_as_uint(w3.w);
            const uint32_t hy0 = __float_as_uint(w4.x), hy1 = __float_as_uint(w4.y), hz0 = __float_as_uint(w4.z), hz1 = __float_as_uint(w4.w);
            const uint32_t nx0 = sx ? hx0 : lx0, fx0 = sx ? lx0 : hx0, nx1 = sx ? hx1 : lx1, fx1 = sx ? lx1 : hx1;
            const uint32_t ny0 = sy ? hy0 : ly0, fy0 = sy ? ly0 : hy0, ny1 = sy ? hy1 : ly1, fy1 = sy ? ly1 : hy1;
            const uint32_t nz0 = sz ? hz0 : lz0, fz0 = sz ? lz0 : hz0, nz1 = sz ? hz1 : lz1, fz1 = sz ? lz1 : hz1;
            uint32_t miss8 = 0u;
            miss8 = shift_in_sign(miss8, slab_k<3>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
            miss8 = shift_in_sign(miss8, slab_k<2>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
            miss8 = shift_in_sign(miss8, slab_k<1>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
            miss8 = shift_in_sign(miss8, slab_k<0>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
            miss8 = shift_in_sign(miss8, slab_k<3>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
            miss8 = shift_in_sign(miss8, slab_k<2>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
            miss8 = shift_in_sign(miss8, slab_k<1>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
            miss8 = shift_in_sign(miss8, slab_k<0>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
            const uint32_t hit8 = ~miss8 & 0xffu, masks = __float_as_uint(w5.x);
            imask = masks & 0xffu; lmask = (masks >> 8) & 0xffu;
            ih = hit8 & imask; th = hit8 & lmask;
            child_base = __float_as_int(w1.z);
            tri_base = __float_as_int(w1.w);
        }
        // ---- the inner children that were hit go onto the group's frontier: prefix sum of the counts over the group ----
        {
            const int mine = __popc(ih);
            int incl = mine;
#pragma unroll
            for (int off = 1; off < G; off <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, off, G);
                if (sub >= off) incl += v;
            }
            const int total = __shfl_sync(FULL, incl, G - 1, G);
            int at = cs + incl - mine;
            for (uint32_t m = ih; m != 0u; m &= m - 1u) {
                const uint32_t s = (uint32_t)__ffs((int)m) - 1u;
                frontier[at++] = child_base + __popc(imask & ((1u << s) - 1u));
            }
            cs += total;
        }
        // ---- the triangle slots that were hit: every lane tests those of its own node, one after the other ----
        float t = 3.0e38f, u = 0.0f, v = 0.0f;
        int32_t id = 0x7fffffff, ti = -1;
        bool blocker = false;
        n_tris += (unsigned long long)__popc(th);
        for (uint32_t m = th; m != 0u; m &= m - 1u) {
            const uint32_t s = (uint32_t)__ffs((int)m) - 1u;
            const int32_t tri_index = tri_base + __popc(lmask & ((1u << s) - 1u));
            const char *tp = reinterpret_cast<const char *>(bvh.tris + tri_index);
            const float4 ta = ld128(tp), tb = ld128(tp + 16), tc = ld128(tp + 32);
            float tt, uu, vv;
            if (!intersect_tri(f3(ta.x, ta.y, ta.z), f3(ta.w, tb.x, tb.y), f3(tb.z, tb.w, tc.x), o, d, tt, uu, vv)) continue;
            const int32_t tid = f2i(tc.y);
            if (!((Alpha && !Any) ? (tt > tmin || (tt == tmin && tid > after_id)) : tt > tmin)) continue;
            if (Any) {
                if (tt < best_t) { // best_t stays the ray's t_max for any-hit rays
                    bool passes = true;
                    if (Alpha) {
                        AlphaFilter af = io.alpha;
                        af.pixel_linear = pixel_linear;
                        passes = shadow_candidate_passes(af, f2i(tc.z), f2i(tc.w), uu, vv);
                    }
                    if (passes) { blocker = true; ti = tri_index; }
                }
            } else {
                // better than the ray's best so far AND than this lane's own candidate of the step
                const bool beats_best = best_tri < 0 ? tt < best_t : (tt < best_t || (tt == best_t && tid < best_id));
                if (beats_best && (ti < 0 || tt < t || (tt == t && tid < id))) { t = tt; u = uu; v = vv; id = tid; ti = tri_index; }
            }
        }
        bool occluded = false;
        if (Any) {
            const unsigned m = __ballot_sync(FULL, blocker) & gmask;
            if (m != 0u) occluded = true; // (best_tri only has to be >= 0)
            if (occluded) best_tri = 0;
        } else {
            const unsigned m = __ballot_sync(FULL, ti >= 0);
            if (m != 0u) { // some group found candidates: arg-min by (t, id) within every group (a group without any keeps ti < 0)
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) {
                    const float ot = __shfl_xor_sync(FULL, t, off, G), ou = __shfl_xor_sync(FULL, u, off, G), ov = __shfl_xor_sync(FULL, v, off, G);
                    const int32_t oid = __shfl_xor_sync(FULL, id, off, G), oti = __shfl_xor_sync(FULL, ti, off, G);
                    if (oti >= 0 && (ti < 0 || ot < t || (ot == t && oid < id))) { t = ot; u = ou; v = ov; id = oid; ti = oti; }
                }
                if (ti >= 0) { best_t = t; best_u = u; best_v = v; best_tri = ti; best_id = id; } // every candidate beat the previous best
            }
        }
        // ---- a ray whose frontier is empty (or that is occluded) is finished ----
        if (active && ((cs == 0 && resume_at >= resume_n) || occluded)) {
            bool again = false;
            if (Alpha && !Any && best_tri >= 0) { // the verdict of the alpha filter on the closest candidate (as at the retire step of the persistent kernel)
                const int32_t ga = bvh.tris[best_tri].gi_alpha;
                if ((((uint32_t)ga) >> 24) != RPTR_TRI_OPAQUE || (ga & RPTR_TRI_TEXTURED_ALPHA)) {
                    uint32_t *ap = io.alpha_lcg + (size_t)slot * io.alpha_stride;
                    uint32_t st = *ap; // every lane of the group draws the same number; its first lane stores the state
                    const uint32_t before = st;
                    const bool rejected = alpha_rejects(candidate_alpha(io.alpha.scene, ga, bvh.tris[best_tri].prim, best_u, best_v), st);
                    __syncwarp(gmask);
                    if (sub == 0 && st != before) *ap = st;
                    if (rejected) { // look for the closest hit after this candidate
                        tmin = best_t; after_id = best_id;
                        best_t = tmax; best_u = 0.0f; best_v = 0.0f; best_tri = -1; best_id = 0x7fffffff;
                        cs = 1;
                        if (sub == 0) frontier[0] = 0;
                        resume_n = 0; resume_at = 0;
                        again = true;
                    }
                }
            }
            if (!again) {
                if (sub == 0) {
                    if (Any) {
                        if (best_tri < 0) { // unoccluded: add the pending NEE contribution
                            const float4 c = io.sh_c[slot];
                            const uint32_t ps = __float_as_uint(c.w);
                            float4 il = io.illum[ps];
                            il.x = il.x + c.x; il.y = il.y + c.y; il.z = il.z + c.z;
                            io.illum[ps] = il;
                        }
                    } else {
                        io.hit[slot] = f4(best_t, best_u, best_v, __int_as_float(best_tri));
                        if (io.hitq && best_tri >= 0) io.hitq[atomicAdd(io.hit_count, 1u)] = slot;
                    }
                }
                active = false;
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_nodes += __shfl_down_sync(FULL, n_nodes, off);
        n_tris += __shfl_down_sync(FULL, n_tris, off);
    }
    if (lane == 0) {
        if (n_nodes) atomicAdd(c_nodes, n_nodes);
        if (n_tris) atomicAdd(c_tris, n_tris);
    }
}

#endif // __CUDACC__

} // namespace rp
