// rptr_trace_kernels.cuh -- the trace stage of the wavefront: persistent-threads BVH traversal.
//
// Ray lengths in the target scenes are close to exponentially distributed (a random triangle soup: neighbouring pixels
// stop at unrelated depths), so a "one ray per thread, wait for the warp" kernel runs at ~14 % SIMD efficiency (ncu:
// 4.6 active threads per instruction, profiles/r01_trace_v1.txt).  This kernel is a per-lane state machine instead; one
// trip through its loop is at most one node step and one leaf step for the whole warp:
//   * every lane owns one ray.  Lanes that have finished are refilled from the ray queue once at least
//     RPTR_REFILL_LANES of them are idle (warp-aggregated fetch from a per-warp chunk: one global atomic per 256 rays);
//   * NODE STEP: every lane whose current item is an inner node fetches it (64 bytes: 2 x 256-bit), slab-tests the four
//     quantised child boxes and picks the next item;
//   * a lane that reaches a leaf parks it in a register and keeps walking inner nodes from its stack (speculative
//     traversal), so nearly all lanes take part in every node step;
//   * LEAF STEP: run only when at least RPTR_LEAF_LANES lanes hold a parked leaf (warp ballot) or nobody has inner-node
//     work left, so the ~4x more expensive triangle code also runs at high lane utilisation.
// Semantics are those of trace_ray<> in rptr_bvh.cuh (same intersect_tri, same tie-break, order independent), which
// stays as the host-executable statement of the contract.
#pragma once
#include "rptr_bvh.cuh"

namespace rp {

#ifndef RPTR_FETCH_CHUNK
#define RPTR_FETCH_CHUNK 256
#endif
#ifndef RPTR_REFILL_LANES
#define RPTR_REFILL_LANES 4
#endif
#ifndef RPTR_LEAF_LANES
#define RPTR_LEAF_LANES 8
#endif

struct TraceIO {
    // rays: closest -> Wave ray_o/ray_d indexed by path slot through `queue` (or identity); shadow -> sh_o/sh_d by index
    const float4 *ray_o;
    const float4 *ray_d;
    const uint32_t *queue; // may be nullptr (identity)
    const uint32_t *count; // number of rays (device resident)
    uint32_t *work;        // global fetch cursor (zeroed before launch)
    float4 *hit;           // closest: (t,u,v,bits(tri)) per path slot
    uint32_t *hitq;        // closest, optional: compacted slots of the rays that hit something (input of the shade stage when set)
    uint32_t *hit_count;
    const float4 *sh_c;    // shadow: (contribution.rgb, bits(path slot))
    float4 *illum;         // shadow: illum.rgb += contribution when unoccluded
    // stochastic alpha (kernels instantiated with Alpha = true only; scenes without alpha-tested triangles never pay for it)
    uint32_t *alpha_lcg;   // closest: LCG state the candidate filter draws from, word alpha_lcg[slot * alpha_stride]: the path's own
    uint32_t alpha_stride; //          LCG with the UNIFORM pointset (stride 2: Wave::rngb), a separate one otherwise (stride 1: Wave::rng3)
    AlphaFilter alpha;     // shadow: per-candidate LCG seeds (pixel_linear is filled in per ray)
    TileMap tm;            // shadow: path slot -> pixel that seeds the per-candidate LCG (frame pixel or ray-query invocation)
};


#if defined(__CUDACC__)

// 256-bit read-only global load (sm_100: LDG.E.ENL2.256.CONSTANT); p must be 32-byte aligned
__device__ __forceinline__ void ld256(const void *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// ---- TMA bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier, raw PTX ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// the same, opaque to ptxas: the value is computed once and kept (S2R SR_CgaCtaId + LEA per use otherwise)
__device__ __forceinline__ uint32_t smem_u32_pinned(const void *p) {
    uint32_t r;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float slab_rcp(float x) {
#if RPTR_FAST_RCP
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

#ifndef RPTR_HC_PER_THREAD
#define RPTR_HC_PER_THREAD 1
#endif
#ifndef RPTR_ANY_UNSORTED
#define RPTR_ANY_UNSORTED 1 // shadow rays: skip the nearest-first sorting network (any hit terminates the ray)
#endif
#ifndef RPTR_LEAF_ONE_PER_TRIP
#define RPTR_LEAF_ONE_PER_TRIP 1 // leaf step tests one triangle of the parked leaf per trip instead of looping over it
#endif
#ifndef RPTR_FAST_RCP
#define RPTR_FAST_RCP 1 // 1/d of the slab test through MUFU.RCP alone (boxes only prune and are padded far beyond 1 ulp of 1/d)
#endif
#ifndef RPTR_PIN_SMEM_BASE
#define RPTR_PIN_SMEM_BASE 1 // compute the shared-window addresses once (asm volatile) instead of letting ptxas rematerialise them per node step
#endif
#ifndef RPTR_CHUNKS_PER_WARP
#define RPTR_CHUNKS_PER_WARP 4 // target number of queue fetches per warp (tail balance) before the chunk is shortened
#endif
#ifndef RPTR_TRACE_THREADS
#define RPTR_TRACE_THREADS 896 // one CTA per SM: 28 warps (<= 72 registers each) share one 64 KB image of the top of the BVH
#endif
#define RPTR_TOP_PLANE_BYTES (RPTR_TOP_NODES_MAX * 16)
// Traversal stack: the first RPTR_SMEM_STACK entries of every thread live in shared memory, laid out [entry][thread] so
// that the bank only depends on the lane (any mix of stack depths in a warp is conflict free: one wavefront per push /
// pop instead of up to 32 sectors through local memory); deeper entries spill to a local-memory array.
#ifndef RPTR_SMEM_STACK
#define RPTR_SMEM_STACK 16
#endif
#define RPTR_TRACE_SMEM_BYTES ((size_t)RPTR_TOP_NODES_MAX * sizeof(BvhNode) + (size_t)RPTR_SMEM_STACK * RPTR_TRACE_THREADS * sizeof(int32_t))
// shared-window accesses by 32-bit address (a generic pointer would cost a window-base computation per push / pop)
__device__ __forceinline__ void sts32(uint32_t addr, int32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int32_t lds32(uint32_t addr) {
    int32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
// byte-permute with the selector as the immediate operand (nvcc otherwise keeps the constant as the immediate and
// re-materialises every selector in a register)
template <int K>
__device__ __forceinline__ float qfloat_k(uint32_t word, uint32_t hi_const) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(hi_const), "n"(0x7044 | (K << 8)));
    return __uint_as_float(r);
}
// slab test of child K against the per-node ray constants, branch free: returns the entry distance, or +inf on a miss.
// qn* / qf* are the packed bounds on the near / far side of each axis (picked per node from the sign of the direction:
// fma is monotonic, so this equals the min / max form of slab_q() bit for bit).
template <int K>
__device__ __forceinline__ float slab_k(const NodeSlab &n, uint32_t qnx, uint32_t qny, uint32_t qnz, uint32_t qfx, uint32_t qfy,
                                        uint32_t qfz, uint32_t hc, float tmin, float tmax, int32_t ref) {
    const float nx = fmaf(qfloat_k<K>(qnx, hc), n.ax, n.bx), fx = fmaf(qfloat_k<K>(qfx, hc), n.ax, n.bx);
    const float ny = fmaf(qfloat_k<K>(qny, hc), n.ay, n.by), fy = fmaf(qfloat_k<K>(qfy, hc), n.ay, n.by);
    const float nz = fmaf(qfloat_k<K>(qnz, hc), n.az, n.bz), fz = fmaf(qfloat_k<K>(qfz, hc), n.az, n.bz);
    float tf = fminf(fminf(fx, fy), fz);
    float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    tf = fminf(tf * 1.0000004f, tmax);
    return (tn <= tf) & (ref != RPTR_EMPTY) ? tn : __int_as_float(0x7f800000);
}
#define RPTR_STACK_STRIDE ((uint32_t)(RPTR_TRACE_THREADS * sizeof(int32_t)))
#define RPTR_PUSH(v)                                                                  \
    {                                                                                 \
        if (sp < RPTR_SMEM_STACK) sts32(sst + (uint32_t)sp * RPTR_STACK_STRIDE, (v)); \
        else lstack[sp - RPTR_SMEM_STACK] = (v);                                      \
        ++sp;                                                                         \
    }
#define RPTR_POP() (sp > 0 ? (--sp, sp < RPTR_SMEM_STACK ? lds32(sst + (uint32_t)sp * RPTR_STACK_STRIDE) : lstack[sp - RPTR_SMEM_STACK]) : RPTR_EMPTY)

// Alpha = true adds the candidate filter of non-opaque triangles (rptr_bvh.cuh, AlphaFilter).  Closest hit: a lane whose
// traversal ended on an alpha-tested triangle draws from its path's LCG when it would retire; if the candidate is rejected
// the lane restarts its traversal for the closest hit AFTER (t, id) of that candidate instead of retiring (front-to-back
// order of DESIGN.md section 5 without a second wavefront pass).  Any hit: a candidate occludes iff its own seeded draw says so.
template <bool Any, bool Alpha>
__global__ void __launch_bounds__(RPTR_TRACE_THREADS, 1) k_trace_persistent(BvhDev bvh, TraceIO io, unsigned long long *c_rays,
                                                                            unsigned long long *c_nodes, unsigned long long *c_tris) {
    extern __shared__ __align__(128) unsigned char smem_top[]; // four word planes of the top_k first nodes, then the stacks
    __shared__ __align__(8) uint64_t top_bar;
    __shared__ uint32_t hc_word;
    const uint32_t n = *io.count;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // ---- stage the top of the tree: TMA bulk copies issued by one thread, completion through an mbarrier ----
    const uint32_t plane_bytes = (uint32_t)bvh.top_k * 16u; // one 16-byte word of each staged node
    if (threadIdx.x == 0) {
        mbar_init(&top_bar, 1);
        hc_word = 0x3f000000u;
    }
    __syncthreads();
    if (threadIdx.x == 0 && n > 0 && plane_bytes > 0) {
        mbar_expect_tx(&top_bar, 4u * plane_bytes);
        for (uint32_t w = 0; w < 4; ++w)
            tma_bulk_g2s(smem_top + w * RPTR_TOP_PLANE_BYTES, reinterpret_cast<const unsigned char *>(bvh.top_planes) + w * RPTR_TOP_PLANE_BYTES,
                         plane_bytes, &top_bar);
    }
    if (n > 0 && plane_bytes > 0) mbar_wait(&top_bar, 0);
    const int32_t top_k = bvh.top_k;
#if RPTR_PIN_SMEM_BASE
    const uint32_t top_base = smem_u32_pinned(smem_top);
#else
    const uint32_t top_base = smem_u32(smem_top);
#endif
    // Rays per queue fetch: RPTR_FETCH_CHUNK for long queues (few atomics, coherent warps); short queues (late bounces)
    // are cut finer so that every warp of the grid gets work instead of a few warps walking 256 rays 32 at a time.
    // Guided self-scheduling: the size is recomputed from what is left of the queue at every fetch.
    const uint32_t chunk_div = gridDim.x * (RPTR_TRACE_THREADS / 32) * RPTR_CHUNKS_PER_WARP;
    uint32_t chunk = min((uint32_t)RPTR_FETCH_CHUNK, max(32u, (n / chunk_div) & ~31u));
    uint32_t pool_pos = 0, pool_end = 0; // per-warp pool of ray indices (warp-uniform)
    bool drained = false;                // the global queue has been exhausted (warp-uniform)

    // per-lane ray state
    bool have = false;
    uint32_t ray_index = 0, slot = 0;
    float3 o = f3(0.0f), d = f3(0.0f), inv = f3(0.0f), ood = f3(0.0f);
    float tmin = 0.0f, tmax = 0.0f;
    float best_t = 0.0f, best_u = 0.0f, best_v = 0.0f;
    int32_t best_tri = -1, best_id = 0x7fffffff;
    int32_t after_id = 0x7fffffff; // Alpha closest: candidates must come after (tmin, after_id) in (t, id) order; tmin doubles as after_t
    uint32_t pixel_linear = 0;     // Alpha any-hit: global pixel of the path the shadow ray belongs to
    int32_t node = RPTR_EMPTY; // current item: inner node (>= 0), leaf reference (< 0) or RPTR_EMPTY
    int32_t leaf = 0;          // parked leaf reference (< 0) or 0 = none
    int32_t lstack[RPTR_STACK_SIZE - RPTR_SMEM_STACK]; // overflow part of the traversal stack (local memory)
#if RPTR_PIN_SMEM_BASE
    const uint32_t sst = top_base + (uint32_t)(RPTR_TOP_NODES_MAX * sizeof(BvhNode)) + threadIdx.x * (uint32_t)sizeof(int32_t);
#else
    const uint32_t sst = smem_u32(smem_top + (size_t)RPTR_TOP_NODES_MAX * sizeof(BvhNode)) + threadIdx.x * (uint32_t)sizeof(int32_t);
#endif
    // high bytes of the decoded box coordinates, kept opaque to ptxas so that it stays in a register and the selectors
    // become immediates (n is never 2^32 - 1)
#if RPTR_HC_PER_THREAD
    // read back from shared memory on purpose: a value ptxas can prove constant or warp-uniform takes the immediate /
    // uniform-register slot of PRMT, and all 24 selectors of a node step are then materialised in registers instead
    const uint32_t hc = (uint32_t)lds32(smem_u32(&hc_word));
#else
    const uint32_t hc = n == 0xffffffffu ? 0u : 0x3f000000u;
#endif
    int sp = 0;
    uint32_t n_nodes = 0, n_tris = 0, n_rays = 0;

    for (;;) {
        __syncwarp();
        // ---- retire + refill --------------------------------------------------------------------------------------
        bool done = have && node == RPTR_EMPTY && leaf == 0;
        if (Alpha && !Any && done && best_tri >= 0 && after_id != RPTR_EMPTY) {
            const int32_t a8 = tri_alpha8(bvh.tris[best_tri]);
            if (a8 != RPTR_TRI_OPAQUE) {
                uint32_t *ap = io.alpha_lcg + (size_t)slot * io.alpha_stride;
                uint32_t st = *ap;
                const uint32_t before = st;
                const bool rejected = alpha_rejects(alpha8_to_float(a8), st);
                if (st != before) *ap = st;
                if (rejected) { // look for the closest hit after this candidate
                    tmin = best_t; after_id = best_id;
                    best_t = tmax; best_u = 0.0f; best_v = 0.0f; best_tri = -1; best_id = 0x7fffffff;
                    sp = 0;
                    node = 0;
                    done = false;
                }
            }
            if (done) after_id = RPTR_EMPTY; // verdict reached: no second draw while the lane waits for the next refill
        }
        const unsigned idle = __ballot_sync(FULL, !have || done);
        const unsigned inner0 = __ballot_sync(FULL, have && node >= 0);
        const unsigned parked0 = __ballot_sync(FULL, have && leaf != 0);
        if (__popc(idle) >= RPTR_REFILL_LANES || (inner0 == 0 && parked0 == 0)) {
            if (done) {
                if (Any) {
                    if (best_tri < 0) { // unoccluded: add the pending NEE contribution (one shadow ray per path and bounce)
                        const float4 c = io.sh_c[ray_index];
                        const uint32_t ps = __float_as_uint(c.w);
                        float4 il = io.illum[ps];
                        il.x = il.x + c.x; il.y = il.y + c.y; il.z = il.z + c.z;
                        io.illum[ps] = il;
                    }
                } else {
                    io.hit[slot] = f4(best_t, best_u, best_v, __int_as_float(best_tri));
                    // paths that left the scene end here (their sky term is added by the resolve kernel); the others
                    // go to the shade stage through a dense queue: one atomic per warp and retire event
                    if (io.hitq) { // warp-uniform
                        const unsigned hits = __ballot_sync(__activemask(), best_tri >= 0);
                        if (best_tri >= 0) {
                            const int leader = __ffs(hits) - 1;
                            uint32_t base = 0;
                            if (lane == leader) base = atomicAdd(io.hit_count, (uint32_t)__popc(hits));
                            base = __shfl_sync(hits, base, leader);
                            io.hitq[base + __popc(hits & ((1u << lane) - 1u))] = slot;
                        }
                    }
                }
                have = false;
            }
            if (pool_pos >= pool_end && !drained) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(io.work, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) drained = true;
                else {
                    pool_pos = base; pool_end = min(base + chunk, n);
                    chunk = min((uint32_t)RPTR_FETCH_CHUNK, max(32u, ((n - pool_end) / chunk_div) & ~31u));
                }
            }
            const uint32_t avail = pool_end > pool_pos ? pool_end - pool_pos : 0u;
            const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
            if (!have && rank < avail) {
                ray_index = pool_pos + rank;
                slot = (Any || !io.queue) ? ray_index : io.queue[ray_index];
                const float4 ro = io.ray_o[Any ? ray_index : slot], rd = io.ray_d[Any ? ray_index : slot];
                o = f3(ro.x, ro.y, ro.z); tmin = ro.w;
                d = f3(rd.x, rd.y, rd.z); tmax = rd.w;
                inv = f3(slab_rcp(slab_safe(d.x)), slab_rcp(slab_safe(d.y)), slab_rcp(slab_safe(d.z)));
                ood = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
                best_t = tmax; best_u = 0.0f; best_v = 0.0f; best_tri = -1; best_id = 0x7fffffff;
                if (Alpha && !Any) after_id = 0x7fffffff;
                if (Alpha && Any) { // global pixel of the path: slot -> local pixel -> row band of this rank
                    pixel_linear = tile_pixel_linear(io.tm, __float_as_uint(io.sh_c[ray_index].w) % (uint32_t)io.tm.local_pixels);
                }
                sp = 0;
                leaf = 0;
                node = bvh.n_nodes > 0 ? 0 : RPTR_EMPTY;
                have = true;
                n_rays++;
            }
            pool_pos += min(avail, (uint32_t)__popc(idle));
            if (drained && !__any_sync(FULL, have)) break;
        }
        // ---- issue every load of this trip up front: the node of lanes with node work AND the first triangle of lanes
        //      whose parked leaf is processed in this trip, so the two dependent fetches overlap instead of queueing ----
        const bool leaf_turn = __popc(parked0) >= RPTR_LEAF_LANES || inner0 == 0;
        const bool do_node = have && node >= 0;
        const bool do_leaf = leaf_turn && have && leaf != 0;
        float4 w0, w1, w2, w3, ta, tb, tc;
        int32_t lf_first = 0, lf_cnt = 0;
        if (do_leaf) {
            const int32_t ref = ~leaf;
            lf_first = ref >> 2;
            lf_cnt = (ref & 3) + 1;
            const char *tp = reinterpret_cast<const char *>(bvh.tris + lf_first);
            ta = ld128(tp); tb = ld128(tp + 16); tc = ld128(tp + 32);
        }
        if (do_node) {
            if (node < top_k) { // top of the tree: shared memory, one LDS.128 per word plane (address = base + 16 * node)
                const uint32_t a = top_base + ((uint32_t)node << 4);
                w0 = lds128<0>(a); w1 = lds128<RPTR_TOP_PLANE_BYTES>(a);
                w2 = lds128<2 * RPTR_TOP_PLANE_BYTES>(a); w3 = lds128<3 * RPTR_TOP_PLANE_BYTES>(a);
            } else { // 2 x 256-bit loads: both 32-byte sectors of the node pass through L1 once
                const char *np = reinterpret_cast<const char *>(bvh.nodes + node);
                ld256(np, w0, w1);
                ld256(np + 32, w2, w3);
            }
        }
        // ---- node step ----------------------------------------------------------------------------------------------
        if (do_node) {
            n_nodes++;
            // four slab tests on the quantised child boxes; a missed or unused child gets key +inf and reference EMPTY
            const float INF = __int_as_float(0x7f800000);
            const NodeSlab ns = node_slab(w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, inv, ood);
            const uint32_t qlx = __float_as_uint(w1.z), qly = __float_as_uint(w1.w), qlz = __float_as_uint(w2.x);
            const uint32_t qhx = __float_as_uint(w2.y), qhy = __float_as_uint(w2.z), qhz = __float_as_uint(w2.w);
            const bool sx = inv.x < 0.0f, sy = inv.y < 0.0f, sz = inv.z < 0.0f;
            const uint32_t qnx = sx ? qhx : qlx, qfx = sx ? qlx : qhx;
            const uint32_t qny = sy ? qhy : qly, qfy = sy ? qly : qhy;
            const uint32_t qnz = sz ? qhz : qlz, qfz = sz ? qlz : qhz;
            int32_t r0 = f2i(w3.x), r1 = f2i(w3.y), r2 = f2i(w3.z), r3 = f2i(w3.w);
            float t0 = slab_k<0>(ns, qnx, qny, qnz, qfx, qfy, qfz, hc, tmin, best_t, r0);
            float t1 = slab_k<1>(ns, qnx, qny, qnz, qfx, qfy, qfz, hc, tmin, best_t, r1);
            float t2 = slab_k<2>(ns, qnx, qny, qnz, qfx, qfy, qfz, hc, tmin, best_t, r2);
            float t3 = slab_k<3>(ns, qnx, qny, qnz, qfx, qfy, qfz, hc, tmin, best_t, r3);
            // 5-comparator sorting network on (t, ref): nearest first
#define RPTR_CSWAP(ta_, ra, tb_, rb)                                 \
    {                                                               \
        const bool sw_ = tb_ < ta_;                                 \
        const float tl_ = sw_ ? tb_ : ta_, th_ = sw_ ? ta_ : tb_;   \
        const int32_t rl_ = sw_ ? rb : ra, rh_ = sw_ ? ra : rb;     \
        ta_ = tl_; tb_ = th_; ra = rl_; rb = rh_;                   \
    }
            if (!(Any && RPTR_ANY_UNSORTED)) {
                RPTR_CSWAP(t0, r0, t1, r1)
                RPTR_CSWAP(t2, r2, t3, r3)
                RPTR_CSWAP(t0, r0, t2, r2)
                RPTR_CSWAP(t1, r1, t3, r3)
                RPTR_CSWAP(t1, r1, t2, r2)
            }
#undef RPTR_CSWAP
            // continue with the nearest hit, push the others farthest first (a key of +inf marks a missed / unused child)
            if (sp + 3 <= RPTR_SMEM_STACK) { // common case, branch free: store unconditionally, advance when the entry is valid
                sts32(sst + (uint32_t)sp * RPTR_STACK_STRIDE, r3); sp += t3 < INF;
                sts32(sst + (uint32_t)sp * RPTR_STACK_STRIDE, r2); sp += t2 < INF;
                sts32(sst + (uint32_t)sp * RPTR_STACK_STRIDE, r1); sp += t1 < INF;
            } else {
                if (t3 < INF) RPTR_PUSH(r3);
                if (t2 < INF) RPTR_PUSH(r2);
                if (t1 < INF) RPTR_PUSH(r1);
            }
            node = t0 < INF ? r0 : RPTR_POP();
            // park a leaf and go on with whatever the stack holds (speculative traversal); when this lane's parked leaf
            // is being processed in this trip the slot frees up below
            if (node < 0 && node != RPTR_EMPTY && leaf == 0) {
                leaf = node;
                node = RPTR_POP();
            }
        }
        // ---- leaf step ------------------------------------------------------------------------------------------------
#if RPTR_LEAF_ONE_PER_TRIP
        // one triangle of the parked leaf per trip (its words were requested at the top of the trip); the leaf reference
        // is advanced in place: ~((first + 1) << 2 | (count - 2)) == leaf - 3
        if (do_leaf) {
            bool occluded = false;
            n_tris++;
            float t, u, v;
            if (intersect_tri(f3(ta.x, ta.y, ta.z), f3(ta.w, tb.x, tb.y), f3(tb.z, tb.w, tc.x), o, d, t, u, v) && t < tmax &&
                ((Alpha && !Any) ? (t > tmin || (t == tmin && f2i(tc.y) > after_id)) : t > tmin)) {
                const int32_t id = f2i(tc.y);
                if (Any) {
                    bool passes = true;
                    if (Alpha) {
                        AlphaFilter af = io.alpha;
                        af.pixel_linear = pixel_linear;
                        passes = shadow_candidate_passes(af, f2i(tc.z), f2i(tc.w));
                    }
                    if (passes) {
                        best_tri = lf_first;
                        occluded = true;
                    }
                } else if (best_tri < 0 || t < best_t || (t == best_t && id < best_id)) {
                    best_t = t; best_u = u; best_v = v; best_tri = lf_first; best_id = id;
                }
            }
            leaf = lf_cnt > 1 ? leaf - 3 : 0;
            if (Any && occluded) { // drop the rest of the traversal
                sp = 0;
                leaf = 0;
                node = RPTR_EMPTY;
            } else if (leaf == 0 && node < 0 && node != RPTR_EMPTY) { // the current item was a second leaf waiting for the slot
                leaf = node;
                node = RPTR_POP();
            }
        }
#else
        // software-pipelined over the (<= 4, contiguous) triangles of the parked leaf
        if (do_leaf) {
            bool occluded = false;
            for (int32_t i = 0; i < lf_cnt; ++i) {
                const float4 a = ta, b = tb, c4 = tc;
                if (i + 1 < lf_cnt) { // fetch the next triangle while this one is tested
                    const char *tp = reinterpret_cast<const char *>(bvh.tris + lf_first + i + 1);
                    ta = ld128(tp); tb = ld128(tp + 16); tc = ld128(tp + 32);
                }
                n_tris++;
                float t, u, v;
                if (!intersect_tri(f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c4.x), o, d, t, u, v)) continue;
                if (!(t < tmax && ((Alpha && !Any) ? (t > tmin || (t == tmin && f2i(c4.y) > after_id)) : t > tmin))) continue;
                const int32_t id = f2i(c4.y);
                if (Any) {
                    if (Alpha) {
                        AlphaFilter af = io.alpha;
                        af.pixel_linear = pixel_linear;
                        if (!shadow_candidate_passes(af, f2i(c4.z), f2i(c4.w))) continue;
                    }
                    best_tri = lf_first + i;
                    occluded = true;
                    break;
                }
                if (best_tri < 0 || t < best_t || (t == best_t && id < best_id)) {
                    best_t = t; best_u = u; best_v = v; best_tri = lf_first + i; best_id = id;
                }
            }
            leaf = 0;
            if (Any && occluded) { // drop the rest of the traversal
                sp = 0;
                node = RPTR_EMPTY;
            } else if (node < 0 && node != RPTR_EMPTY) { // the current item was a second leaf waiting for the slot
                leaf = node;
                node = RPTR_POP();
            }
        }
#endif
    }
    // ---- counters ----
    unsigned long long a = n_rays, b = n_nodes, c = n_tris;
    for (int s = 16; s > 0; s >>= 1) {
        a += __shfl_down_sync(FULL, a, s);
        b += __shfl_down_sync(FULL, b, s);
        c += __shfl_down_sync(FULL, c, s);
    }
    if (lane == 0) {
        if (a) atomicAdd(c_rays, a);
        if (b) atomicAdd(c_nodes, b);
        if (c) atomicAdd(c_tris, c);
    }
}

#endif // __CUDACC__

} // namespace rp
