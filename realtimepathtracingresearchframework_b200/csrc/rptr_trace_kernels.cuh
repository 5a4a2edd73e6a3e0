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

// A ray the persistent kernel hands over to the tail kernel (rptr_trace_tail.cuh): where it is read from and what is known so far
#define RPTR_TAIL_GROUPS 16 // >= 2 + RPTR_SMEM_STACK + RPTR_TRI_BACKLOG: what a lane can have pending without its local-memory overflow stack
struct TailRec {
    uint32_t slot;    // closest: path slot; shadow: index of the shadow ray
    float tmin;       // closest with the alpha filter: t of the last rejected candidate
    int32_t after_id; //                                its id
    float best_t, best_u, best_v;
    int32_t best_tri, best_id;
    // where the traversal stood: node groups (bottom of the stack first, the lane's current group last), then triangle groups, as
    // (base, masks) pairs of the persistent kernel.  n_node_groups < 0: not recorded (deep stack) -- the tail kernel restarts at the root
    int32_t n_node_groups, n_tri_groups;
    uint32_t gx[RPTR_TAIL_GROUPS], gy[RPTR_TAIL_GROUPS];
};

struct TraceIO {
    // rays: closest -> Wave ray_o/ray_d indexed by path slot through `queue` (or identity); shadow -> sh_o/sh_d by index
    const float4 *ray_o;
    const float4 *ray_d;
    const uint32_t *queue; // order in which the rays are fetched (path slots / shadow-ray indices); may be nullptr (identity)
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
    AlphaFilter alpha;     // scene tables for the alpha of textured candidates; shadow: per-candidate LCG seeds (pixel_linear is filled in per ray)
    TileMap tm;            // shadow: path slot -> pixel that seeds the per-candidate LCG (frame pixel or ray-query invocation)
    // tail hand-over (nullptr: the persistent kernel finishes every ray itself)
    TailRec *tail;
    uint32_t *tail_count;
    uint32_t *tail_cursor; // next record the tail kernel hands to a group of lanes (zero at launch)
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
#ifndef RPTR_NO_OOD
#define RPTR_NO_OOD 1 // o / d is recomputed per node step (three multiplies on the idle FMA pipe) instead of living in three registers
#endif
#ifndef RPTR_FAST_RCP
#define RPTR_FAST_RCP 1 // 1/d of the slab test through MUFU.RCP alone (boxes only prune and are padded far beyond 1 ulp of 1/d)
#endif
#ifndef RPTR_SIGN_MASK
#define RPTR_SIGN_MASK 1 // hit mask of a node step from the sign bits of (tfar - tnear)
#endif
#ifndef RPTR_NO_WIDEN
#define RPTR_NO_WIDEN 0 // 1: tfar is not widened by 4 ulp (the builder's padding alone keeps the slab test conservative)
#endif
#ifndef RPTR_TAIL_LIVE
#define RPTR_TAIL_LIVE 8 // a drained warp hands its rays over to the tail kernel once at most this many are alive (sweep: profiles/r02_sweeps.md)
#endif
#ifndef RPTR_CHUNKS_PER_WARP
#define RPTR_CHUNKS_PER_WARP 4 // target number of queue fetches per warp (tail balance) before the chunk is shortened
#endif
#ifndef RPTR_TRACE_THREADS
#define RPTR_TRACE_THREADS 896 // one CTA per SM: 28 warps share one 96 KB image of the top of the BVH
#endif
#define RPTR_TOP_PLANE_BYTES (RPTR_TOP_NODES_MAX * 16)
#define RPTR_TOP_BYTES (RPTR_NODE_WORDS * RPTR_TOP_PLANE_BYTES)
#define RPTR_LUT_BYTES 2048
// Traversal stack of (base, masks) groups: the first RPTR_SMEM_STACK entries of every thread live in shared memory as two
// planes of 32-bit words laid out [entry][thread], so that the bank only depends on the lane (any mix of stack depths in a
// warp is conflict free); deeper entries spill to a local-memory array.
#ifndef RPTR_SMEM_STACK
#define RPTR_SMEM_STACK 8
#endif
#define RPTR_STACK_PLANE_BYTES ((uint32_t)(RPTR_SMEM_STACK * RPTR_TRACE_THREADS * sizeof(uint32_t)))
#ifndef RPTR_TRI_BACKLOG
#define RPTR_TRI_BACKLOG 3 // with 832 staged nodes the CTA stays inside the 164 KB shared-memory configuration: 92 KB of L1 remain (profiles/r02_sweeps.md)
#endif
// the backlog of triangle groups: shared memory as well, same [entry][thread] layout (it lived in local memory first: every pop
// was an L1 / L2 round trip the whole warp waited for at the top of the next trip -- 6 % of all stall samples, profiles/r02_trace_source_stalls.md)
#define RPTR_TRI_PLANE_BYTES ((uint32_t)(RPTR_TRI_BACKLOG * RPTR_TRACE_THREADS * sizeof(uint32_t)))
static_assert(2 + RPTR_SMEM_STACK + RPTR_TRI_BACKLOG <= RPTR_TAIL_GROUPS, "TailRec holds the node-group stack, the backlog and the two current groups");
#define RPTR_TRACE_SMEM_BYTES ((size_t)RPTR_TOP_BYTES + RPTR_LUT_BYTES + 2 * (size_t)RPTR_STACK_PLANE_BYTES + 2 * (size_t)RPTR_TRI_PLANE_BYTES)
#ifndef RPTR_NODE_REPS
#define RPTR_NODE_REPS 2 // node steps per trip of the loop (the ballots / refill checks of a trip are paid once)
#endif
// shared-window accesses by 32-bit address (a generic pointer would cost a window-base computation per push / pop)
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
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
// slab test of the child in byte K of the six words, branch free: a value whose sign bit is CLEAR when the padded box meets the
// ray inside (tmin, tmax].  The caller collects the eight sign bits with one funnel shift each (instead of a compare, a select
// and a third of an add per slot: the kernel is bound by instruction issue, and the subtraction runs on the idle FMA pipe).
// qn* / qf* are the packed bounds on the near / far side of each axis (picked per node from the sign of the direction:
// fma is monotonic, so this equals the min / max form of slab_q() bit for bit).
template <int K>
__device__ __forceinline__ float slab_k(const NodeSlab &n, uint32_t qnx, uint32_t qny, uint32_t qnz, uint32_t qfx, uint32_t qfy,
                                        uint32_t qfz, uint32_t hc, float tmin, float tmax) {
    const float nx = fmaf(qfloat_k<K>(qnx, hc), n.ax, n.bx), fx = fmaf(qfloat_k<K>(qfx, hc), n.ax, n.bx);
    const float ny = fmaf(qfloat_k<K>(qny, hc), n.ay, n.by), fy = fmaf(qfloat_k<K>(qfy, hc), n.ay, n.by);
    const float nz = fmaf(qfloat_k<K>(qnz, hc), n.az, n.bz), fz = fmaf(qfloat_k<K>(qfz, hc), n.az, n.bz);
    float tf = fminf(fminf(fx, fy), fz);
    const float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
#if RPTR_NO_WIDEN
    tf = fminf(tf, tmax);
#else
    tf = fminf(tf * 1.0000004f, tmax);
#endif
#if RPTR_SIGN_MASK
    return tf - tn; // hit <=> tn <= tf <=> the sign bit of tf - tn is clear (all operands are finite; x - x = +0)
#else
    return tn <= tf ? 0.0f : -1.0f;
#endif
}
// the sign bit of x shifted into m from the right (one funnel shift)
__device__ __forceinline__ uint32_t shift_in_sign(uint32_t m, float x) { return __funnelshift_l(__float_as_uint(x), m, 1); }

// Alpha = true adds the candidate filter of non-opaque triangles (rptr_bvh.cuh, AlphaFilter).  Closest hit: a lane whose
// traversal ended on an alpha-tested triangle draws from its path's LCG when it would retire; if the candidate is rejected
// the lane restarts its traversal for the closest hit AFTER (t, id) of that candidate instead of retiring (front-to-back
// order of DESIGN.md section 5 without a second wavefront pass).  Any hit: a candidate occludes iff its own seeded draw says so.
//
// Per-lane state machine over the eight-wide tree (rptr_bvh.cuh).  A lane holds one NODE GROUP G = (child_base, imask << 8 |
// pending inner hits in priority order) and one TRIANGLE GROUP T = (tri_base, lmask << 8 | pending triangle hits) in
// registers and further groups on its stack.  One trip through the loop is at most one node step and one triangle step for
// the warp: the node step takes the highest-priority pending child of G, slab-tests the eight slots of that node and
// replaces G (the rest of the old group goes to the stack); triangle hits become T, or a stack entry while T is busy.  The
// triangle step tests ONE pending triangle of T and runs only when at least RPTR_LEAF_LANES lanes hold one (warp ballot)
// or no lane has node work, so that the triangle code runs at high lane utilisation too.
template <bool Any, bool Alpha>
__global__ void __launch_bounds__(RPTR_TRACE_THREADS, 1) k_trace_persistent(BvhDev bvh, TraceIO io, unsigned long long *c_rays,
                                                                            unsigned long long *c_nodes, unsigned long long *c_tris) {
    extern __shared__ __align__(128) unsigned char smem_top[]; // six word planes of the top_k first nodes, the LUT, the stacks
    __shared__ __align__(8) uint64_t top_bar;
    __shared__ uint32_t hc_word;
    __shared__ uint32_t s_cnt[RPTR_TRACE_THREADS / 32][4]; // per warp: rays, node steps, triangle tests
    if (threadIdx.x < RPTR_TRACE_THREADS / 32) s_cnt[threadIdx.x][0] = s_cnt[threadIdx.x][1] = s_cnt[threadIdx.x][2] = 0u;
    const uint32_t n = *io.count;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // ---- stage the top of the tree: TMA bulk copies issued by one thread, completion through an mbarrier ----
    const uint32_t plane_bytes = (uint32_t)bvh.top_k * 16u; // one 16-byte word of each staged node
    if (threadIdx.x == 0) {
        mbar_init(&top_bar, 1);
        hc_word = 0x3f000000u;
    }
    // priority permutation of an 8-bit hit mask: lut[oct << 8 | m] has bit (s ^ oct) set iff bit s of m is set
    for (uint32_t i = threadIdx.x; i < RPTR_LUT_BYTES; i += RPTR_TRACE_THREADS) {
        const uint32_t oct = i >> 8, m = i & 0xffu;
        uint32_t r = 0;
#pragma unroll
        for (uint32_t sl = 0; sl < 8; ++sl) r |= ((m >> sl) & 1u) << (sl ^ oct);
        smem_top[RPTR_TOP_BYTES + i] = (unsigned char)r;
    }
    __syncthreads();
    if (threadIdx.x == 0 && n > 0 && plane_bytes > 0) {
        mbar_expect_tx(&top_bar, (uint32_t)RPTR_NODE_WORDS * plane_bytes);
        for (uint32_t w = 0; w < RPTR_NODE_WORDS; ++w)
            tma_bulk_g2s(smem_top + w * RPTR_TOP_PLANE_BYTES, reinterpret_cast<const unsigned char *>(bvh.top_planes) + w * RPTR_TOP_PLANE_BYTES,
                         plane_bytes, &top_bar);
    }
    if (n > 0 && plane_bytes > 0) mbar_wait(&top_bar, 0);
    const int32_t top_k = bvh.top_k;
    const uint32_t top_base = smem_u32_pinned(smem_top);
    const uint32_t lut_base = top_base + (uint32_t)RPTR_TOP_BYTES;
    // Rays per queue fetch: RPTR_FETCH_CHUNK for long queues (few atomics, coherent warps); short queues (late bounces)
    // are cut finer so that every warp of the grid gets work instead of a few warps walking 256 rays 32 at a time.
    // Guided self-scheduling: the size is recomputed from what is left of the queue at every fetch.
    const uint32_t chunk_div = gridDim.x * (RPTR_TRACE_THREADS / 32) * RPTR_CHUNKS_PER_WARP;
    uint32_t chunk = min((uint32_t)RPTR_FETCH_CHUNK, max(32u, (n / chunk_div) & ~31u));
    uint32_t pool_pos = 0, pool_end = 0; // per-warp pool of ray indices (warp-uniform)
    bool drained = false;                // the global queue has been exhausted (warp-uniform)

    // per-lane ray state
    bool have = false;
    uint32_t slot = 0;
    float3 o = f3(0.0f), d = f3(0.0f), inv = f3(0.0f);
#if !RPTR_NO_OOD
    float3 ood = f3(0.0f); // o / d, kept per ray (RPTR_NO_OOD: recomputed per node step -- three multiplies for three registers)
#endif
    float tmin = 0.0f;
    float best_t = 0.0f, best_u = 0.0f, best_v = 0.0f;
    int32_t best_tri = -1, best_id = 0x7fffffff;
    int32_t after_id = 0x7fffffff; // Alpha closest: candidates must come after (tmin, after_id) in (t, id) order; tmin doubles as after_t
    uint32_t pixel_linear = 0;     // Alpha any-hit: pixel of the path the shadow ray belongs to
    uint32_t gx = 0, gy = 0;       // node group: first child node, imask << 8 | pending inner hits (priority order: bit p = slot ^ oct)
    uint32_t tx = 0, ty = 0;       // triangle group: first triangle, lmask << 8 | pending triangle hits (slot order)
    uint32_t oct = 0;              // ray octant (closest hit only: any-hit rays take the children in slot order)
    uint32_t lstack_x[RPTR_MAX_BVH_DEPTH + 2 - RPTR_SMEM_STACK], lstack_y[RPTR_MAX_BVH_DEPTH + 2 - RPTR_SMEM_STACK]; // overflow part of the node-group stack (local memory)
    // triangle groups that arrive while T is busy (shared memory, two planes [entry][thread]); a lane whose backlog is full pauses
    // its node steps until the triangle step has caught up, so the backlog is bounded
    int tsp = 0;
    const uint32_t sst = lut_base + (uint32_t)RPTR_LUT_BYTES + threadIdx.x * (uint32_t)sizeof(uint32_t);
    const uint32_t tst = sst + 2u * RPTR_STACK_PLANE_BYTES;
    // high bytes of the decoded box coordinates: read back from shared memory on purpose -- a value ptxas can prove constant or
    // warp-uniform takes the immediate / uniform-register slot of PRMT, and all 48 selectors of a node step are then
    // materialised in registers instead
#if RPTR_HC_PER_THREAD
    const uint32_t hc = lds32(smem_u32(&hc_word));
#else
    const uint32_t hc = n == 0xffffffffu ? 0u : 0x3f000000u;
#endif
    int sp = 0;
    // statistics (rays, node steps, triangle tests): ballot counts added to per-warp words in shared memory by lane 0 (reductions
    // without a return value: nothing waits for them) -- per-lane counters in registers were spilled to local memory by the
    // register allocator and every increment stalled the warp on a local-memory round trip (7 % of all stall samples)
#define RPTR_COUNT(k, v)                                                          \
    {                                                                            \
        const uint32_t v_ = (uint32_t)(v); /* evaluated by the whole warp (ballots) */ \
        if (lane == 0) atomicAdd(&s_cnt[threadIdx.x >> 5][k], v_);                \
    }

#define RPTR_STACK_STRIDE ((uint32_t)(RPTR_TRACE_THREADS * sizeof(uint32_t)))
#define RPTR_PUSH(vx, vy)                                                        \
    {                                                                            \
        if (sp < RPTR_SMEM_STACK) {                                              \
            const uint32_t a_ = sst + (uint32_t)sp * RPTR_STACK_STRIDE;          \
            sts32(a_, (vx));                                                     \
            sts32(a_ + RPTR_STACK_PLANE_BYTES, (vy));                            \
        } else {                                                                 \
            lstack_x[sp - RPTR_SMEM_STACK] = (vx);                               \
            lstack_y[sp - RPTR_SMEM_STACK] = (vy);                               \
        }                                                                        \
        ++sp;                                                                    \
    }
#define RPTR_POP(vx, vy)                                                         \
    {                                                                            \
        --sp;                                                                    \
        if (sp < RPTR_SMEM_STACK) {                                              \
            const uint32_t a_ = sst + (uint32_t)sp * RPTR_STACK_STRIDE;          \
            (vx) = lds32(a_);                                                    \
            (vy) = lds32(a_ + RPTR_STACK_PLANE_BYTES);                           \
        } else {                                                                 \
            (vx) = lstack_x[sp - RPTR_SMEM_STACK];                               \
            (vy) = lstack_y[sp - RPTR_SMEM_STACK];                               \
        }                                                                        \
    }

    for (;;) {
        __syncwarp();
        // ---- retire + refill --------------------------------------------------------------------------------------
        bool done = have && (gy & 0xffu) == 0u && (ty & 0xffu) == 0u && sp == 0 && tsp == 0;
        if (Alpha && !Any && done && best_tri >= 0 && after_id != RPTR_EMPTY) {
            const int32_t ga = bvh.tris[best_tri].gi_alpha;
            if ((((uint32_t)ga) >> 24) != RPTR_TRI_OPAQUE || (ga & RPTR_TRI_TEXTURED_ALPHA)) {
                uint32_t *ap = io.alpha_lcg + (size_t)slot * io.alpha_stride;
                uint32_t st = *ap;
                const uint32_t before = st;
                const bool rejected = alpha_rejects(candidate_alpha(io.alpha.scene, ga, bvh.tris[best_tri].prim, best_u, best_v), st);
                if (st != before) *ap = st;
                if (rejected) { // look for the closest hit after this candidate
                    tmin = best_t; after_id = best_id;
                    best_t = io.ray_d[slot].w; best_u = 0.0f; best_v = 0.0f; best_tri = -1; best_id = 0x7fffffff;
                    gx = 0u; gy = 0x100u | (1u << oct);
                    done = false;
                }
            }
            if (done) after_id = RPTR_EMPTY; // verdict reached: no second draw while the lane waits for the next refill
        }
        const unsigned idle = __ballot_sync(FULL, !have || done);
        const unsigned inner0 = __ballot_sync(FULL, have && (gy & 0xffu) != 0u && tsp < RPTR_TRI_BACKLOG);
        const unsigned parked0 = __ballot_sync(FULL, have && (ty & 0xffu) != 0u);
        if (__popc(idle) >= RPTR_REFILL_LANES || (inner0 == 0 && parked0 == 0)) {
            if (done) {
                if (Any) {
                    if (best_tri < 0) { // unoccluded: add the pending NEE contribution (one shadow ray per path and bounce)
                        const float4 c = io.sh_c[slot];
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
                const uint32_t ray_index = pool_pos + rank;
                slot = io.queue ? io.queue[ray_index] : ray_index; // closest: path slot; shadow: index of the shadow ray
                const float4 ro = io.ray_o[slot], rd = io.ray_d[slot];
                o = f3(ro.x, ro.y, ro.z); tmin = ro.w;
                d = f3(rd.x, rd.y, rd.z);
                inv = f3(slab_rcp(slab_safe(d.x)), slab_rcp(slab_safe(d.y)), slab_rcp(slab_safe(d.z)));
#if !RPTR_NO_OOD
                ood = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
#endif
                best_t = rd.w; best_u = 0.0f; best_v = 0.0f; best_tri = -1; best_id = 0x7fffffff;
                if (Alpha && !Any) after_id = 0x7fffffff;
                if (Alpha && Any) pixel_linear = tile_pixel_linear(io.tm, __float_as_uint(io.sh_c[slot].w) % (uint32_t)io.tm.local_pixels);
                oct = Any ? 0u : ray_octant(d);
                sp = 0;
                tsp = 0;
                tx = 0u; ty = 0u;
                gx = 0u;
                gy = bvh.n_nodes > 0 ? (0x100u | (1u << oct)) : 0u; // the root as the only child of a virtual group: slot 0, priority 0 ^ oct
                have = true;
            }
            RPTR_COUNT(0, min(avail, (uint32_t)__popc(idle)));
            pool_pos += min(avail, (uint32_t)__popc(idle));
            if (drained && !__any_sync(FULL, have)) break;
        }
        // ---- the tail of the launch: the queue is empty and this warp is down to its last rays -> they go to the tail kernel, where
        //      a whole warp works on each (rptr_trace_tail.cuh), and the warp's slot on the SM is free for the next launch ----
        if (io.tail && drained) { // warp-uniform
            const unsigned live = __ballot_sync(FULL, have);
            if (__popc(live) <= RPTR_TAIL_LIVE) {
                if (have) {
                    TailRec tr;
                    tr.slot = slot; tr.tmin = tmin; tr.after_id = (Alpha && !Any) ? after_id : 0x7fffffff;
                    tr.best_t = best_t; tr.best_u = best_u; tr.best_v = best_v; tr.best_tri = best_tri; tr.best_id = best_id;
                    TailRec *out = io.tail + atomicAdd(io.tail_count, 1u);
                    int ng = 0, nt = 0;
                    if (sp > RPTR_SMEM_STACK) ng = -1; // part of the stack is in local memory: restart at the root
                    else {
                        for (int i = 0; i < sp; ++i) {
                            out->gx[ng] = lds32(sst + (uint32_t)i * RPTR_STACK_STRIDE);
                            out->gy[ng] = lds32(sst + (uint32_t)i * RPTR_STACK_STRIDE + RPTR_STACK_PLANE_BYTES);
                            ++ng;
                        }
                        if ((gy & 0xffu) != 0u) { out->gx[ng] = gx; out->gy[ng] = gy; ++ng; }
                        for (int i = 0; i < tsp; ++i) {
                            out->gx[ng + nt] = lds32(tst + (uint32_t)i * RPTR_STACK_STRIDE);
                            out->gy[ng + nt] = lds32(tst + (uint32_t)i * RPTR_STACK_STRIDE + RPTR_TRI_PLANE_BYTES);
                            ++nt;
                        }
                        if ((ty & 0xffu) != 0u) { out->gx[ng + nt] = tx; out->gy[ng + nt] = ty; ++nt; }
                    }
                    tr.n_node_groups = ng; tr.n_tri_groups = nt;
                    // header only (the groups were written in place)
                    out->slot = tr.slot; out->tmin = tr.tmin; out->after_id = tr.after_id; out->best_t = tr.best_t; out->best_u = tr.best_u;
                    out->best_v = tr.best_v; out->best_tri = tr.best_tri; out->best_id = tr.best_id; out->n_node_groups = ng; out->n_tri_groups = nt;
                }
                break;
            }
        }
        // ---- the triangle of lanes whose triangle group is processed in this trip is requested first, so that its latency
        //      hides behind the node step ----
        const bool leaf_turn = __popc(parked0) >= RPTR_LEAF_LANES || inner0 == 0;
        const bool do_leaf = leaf_turn && have && (ty & 0xffu) != 0u;
        if (leaf_turn && parked0 != 0u) RPTR_COUNT(2, __popc(parked0));
        float4 ta, tb, tc;
        int32_t tri_index = 0;
        if (do_leaf) {
            const uint32_t s = (uint32_t)__ffs((int)(ty & 0xffu)) - 1u;
            tri_index = (int32_t)(tx + (uint32_t)__popc((ty >> 8) & ((1u << s) - 1u)));
            ty &= ~(1u << s);
            const char *tp = reinterpret_cast<const char *>(bvh.tris + tri_index);
            ta = ld128(tp); tb = ld128(tp + 16); tc = ld128(tp + 32);
        }
        // ---- node steps ---------------------------------------------------------------------------------------------
#pragma unroll
        for (int rep = 0; rep < RPTR_NODE_REPS; ++rep) {
            RPTR_COUNT(1, __popc(__ballot_sync(FULL, have && (gy & 0xffu) != 0u && tsp < RPTR_TRI_BACKLOG)));
            if (have && (gy & 0xffu) != 0u && tsp < RPTR_TRI_BACKLOG) {
                // the highest-priority pending child of the group; what is left of the group goes to the stack
                const uint32_t p = 31u - (uint32_t)__clz((int)(gy & 0xffu));
                gy &= ~(1u << p);
                const uint32_t sl = Any ? p : (p ^ oct);
                const int32_t node = (int32_t)(gx + (uint32_t)__popc((gy >> 8) & ((1u << sl) - 1u)));
                if ((gy & 0xffu) != 0u) RPTR_PUSH(gx, gy);
                float4 w0, w1, w2, w3, w4, w5;
                if (node < top_k) { // top of the tree: shared memory, one LDS.128 per word plane (address = base + 16 * node)
                    const uint32_t a = top_base + ((uint32_t)node << 4);
                    w0 = lds128<0>(a); w1 = lds128<RPTR_TOP_PLANE_BYTES>(a); w2 = lds128<2 * RPTR_TOP_PLANE_BYTES>(a);
                    w3 = lds128<3 * RPTR_TOP_PLANE_BYTES>(a); w4 = lds128<4 * RPTR_TOP_PLANE_BYTES>(a); w5 = lds128<5 * RPTR_TOP_PLANE_BYTES>(a);
                } else { // 3 x 256-bit loads: every 32-byte sector of the node passes through L1 once
                    const char *np = reinterpret_cast<const char *>(bvh.nodes + node);
                    ld256(np, w0, w1);
                    ld256(np + 32, w2, w3);
                    ld256(np + 64, w4, w5);
                }
#if RPTR_NO_OOD
                const float3 ood = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
#endif
                const NodeSlab ns = node_slab(w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, inv, ood);
                // packed bounds: qlo x / y / z = (w2.x w2.y) (w2.z w2.w) (w3.x w3.y), qhi x / y / z = (w3.z w3.w) (w4.x w4.y) (w4.z w4.w)
                const bool sx = inv.x < 0.0f, sy = inv.y < 0.0f, sz = inv.z < 0.0f;
                const uint32_t lx0 = __float_as_uint(w2.x), lx1 = __float_as_uint(w2.y), ly0 = __float_as_uint(w2.z), ly1 = __float_as_uint(w2.w);
                const uint32_t lz0 = __float_as_uint(w3.x), lz1 = __float_as_uint(w3.y), hx0 = __float_as_uint(w3.z), hx1 = __float_as_uint(w3.w);
                const uint32_t hy0 = __float_as_uint(w4.x), hy1 = __float_as_uint(w4.y), hz0 = __float_as_uint(w4.z), hz1 = __float_as_uint(w4.w);
                const uint32_t nx0 = sx ? hx0 : lx0, fx0 = sx ? lx0 : hx0, nx1 = sx ? hx1 : lx1, fx1 = sx ? lx1 : hx1;
                const uint32_t ny0 = sy ? hy0 : ly0, fy0 = sy ? ly0 : hy0, ny1 = sy ? hy1 : ly1, fy1 = sy ? ly1 : hy1;
                const uint32_t nz0 = sz ? hz0 : lz0, fz0 = sz ? lz0 : hz0, nz1 = sz ? hz1 : lz1, fz1 = sz ? lz1 : hz1;
                uint32_t miss8 = 0u; // slot 7 first: bit k of the result belongs to slot k
                miss8 = shift_in_sign(miss8, slab_k<3>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
                miss8 = shift_in_sign(miss8, slab_k<2>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
                miss8 = shift_in_sign(miss8, slab_k<1>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
                miss8 = shift_in_sign(miss8, slab_k<0>(ns, nx1, ny1, nz1, fx1, fy1, fz1, hc, tmin, best_t));
                miss8 = shift_in_sign(miss8, slab_k<3>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
                miss8 = shift_in_sign(miss8, slab_k<2>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
                miss8 = shift_in_sign(miss8, slab_k<1>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
                miss8 = shift_in_sign(miss8, slab_k<0>(ns, nx0, ny0, nz0, fx0, fy0, fz0, hc, tmin, best_t));
                const uint32_t hit8 = ~miss8 & 0xffu;
                const uint32_t masks = __float_as_uint(w5.x);
                const uint32_t ih = hit8 & masks, th = hit8 & (masks >> 8); // empty slots have neither bit
                gx = __float_as_uint(w1.z);
                gy = ((masks & 0xffu) << 8) | (Any ? ih : lds8(lut_base + ((oct << 8) | ih)));
                if (th != 0u) {
                    const uint32_t nty = (masks & 0xff00u) | th;
                    if ((ty & 0xffu) != 0u) {
                        const uint32_t a_ = tst + (uint32_t)tsp * RPTR_STACK_STRIDE;
                        sts32(a_, __float_as_uint(w1.w));
                        sts32(a_ + RPTR_TRI_PLANE_BYTES, nty);
                        ++tsp;
                    }
                    else { tx = __float_as_uint(w1.w); ty = nty; }
                }
                if ((gy & 0xffu) == 0u && sp > 0) RPTR_POP(gx, gy); // nothing hit: back to the closest pending group
            }
        }
        // ---- triangle step: one triangle of the group per trip (its words were requested at the top of the trip) -------------
        if (do_leaf) {
            float t, u, v;
            if (intersect_tri(f3(ta.x, ta.y, ta.z), f3(ta.w, tb.x, tb.y), f3(tb.z, tb.w, tc.x), o, d, t, u, v) &&
                ((Alpha && !Any) ? (t > tmin || (t == tmin && f2i(tc.y) > after_id)) : t > tmin)) {
                const int32_t id = f2i(tc.y);
                if (Any) {
                    if (t < best_t) { // best_t stays the ray's t_max for any-hit rays
                        bool passes = true;
                        if (Alpha) {
                            AlphaFilter af = io.alpha;
                            af.pixel_linear = pixel_linear;
                            passes = shadow_candidate_passes(af, f2i(tc.z), f2i(tc.w), u, v);
                        }
                        if (passes) { // occluded: drop the rest of the traversal
                            best_tri = tri_index;
                            sp = 0;
                            tsp = 0;
                            gy = 0u; ty = 0u;
                        }
                    }
                } else if (best_tri < 0 ? t < best_t : (t < best_t || (t == best_t && id < best_id))) {
                    best_t = t; best_u = u; best_v = v; best_tri = tri_index; best_id = id;
                }
            }
        }
        if (have && (ty & 0xffu) == 0u && tsp > 0) {
            --tsp;
            const uint32_t a_ = tst + (uint32_t)tsp * RPTR_STACK_STRIDE;
            tx = lds32(a_);
            ty = lds32(a_ + RPTR_TRI_PLANE_BYTES);
        }
    }
#undef RPTR_PUSH
#undef RPTR_POP
    // ---- counters ----
#undef RPTR_COUNT
    if (lane == 0) {
        const uint32_t *c = s_cnt[threadIdx.x >> 5];
        if (c[0]) atomicAdd(c_rays, (unsigned long long)c[0]);
        if (c[1]) atomicAdd(c_nodes, (unsigned long long)c[1]);
        if (c[2]) atomicAdd(c_tris, (unsigned long long)c[2]);
    }
}

#endif // __CUDACC__

} // namespace rp
