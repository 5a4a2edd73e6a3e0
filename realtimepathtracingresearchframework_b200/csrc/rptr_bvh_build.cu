// rptr_bvh_build.cu -- device-side BVH builder (option "bvh_builder" = 1): a top-down binned-SAH build run entirely on the GPU,
// level by level, followed by the SAH-optimal collapse into the eight-wide layout the trace kernels read (rptr_bvh.cuh).
//
// Stands in for vkCmdBuildAccelerationStructuresKHR (vulkan/vulkanrt_utils.cpp:82-167,241-300): the reference contains no
// BVH algorithm of its own, so only closest-hit RESULTS have to match -- which they do for any
// conservative tree, so images rendered with this builder are bit-identical to those of the host SAH builder
// (tests/test_gpu_parity.py::test_device_builder_gives_identical_images).
//
// The algorithm is the one of the host builder (rptr_host.cpp, Builder::build: 16 centroid bins on three axes for nodes of more
// than eight triangles, exact sweep below that), restated breadth-first so that one level of the tree is a handful of launches
// over the whole triangle array.  No sort and no library call is involved:
//   k_prims        padded triangle boxes, root bounds (warp-reduced atomics on order-preserving integer keys)
//   per level      k_bins_clear, k_bin (shared-memory bins per 2048-triangle chunk when the chunk lies inside one node, global
//                  atomics otherwise), k_split (one thread per node: SAH sweep over 3 x 15 bin boundaries, children appended to the
//                  next level / the small-node list / linked as single triangles), then a STABLE partition of every node's
//                  triangle range: k_side_sums + k_scan_sums + k_scan_write (exclusive scan of the "goes left" predicate over the
//                  whole array) and k_scatter
//   k_sweep        one thread per node of at most eight triangles: exact sweep SAH down to single triangles, collapse costs
//   k_collapse_dp  dynamic programme of the collapse, level by level from the bottom
//   k_collapse     wide nodes breadth-first: slots, consecutive children, consecutive leaf-order triangle records
//   k_top_planes   shared-memory image of the first nodes
// The binary tree always ends in single triangles; the node between positions (mid - 1, mid) of the final triangle order has the
// index mid - 1, so node indices need no allocation and the tree -- every box, every link -- is the same on every run.
#include <cuda_runtime.h>

#include <vector>

#include "rptr_bvh_build.hpp"

namespace rp {

namespace {

struct Box { float lo[3], hi[3]; };
struct Node2 { int32_t left, right; }; // child >= 0: inner node; < 0: ~(position in the final triangle order)

#define RPTR_SAH_BINS 16
#define RPTR_SAH_SMALL 8       // nodes of at most this many triangles are finished by k_sweep
#define RPTR_SAH_DEPTH_LIMIT 32 // below this level nodes are halved by position (as the host builder does)
#define RPTR_SAH_MAX_LEVELS 96
#define RPTR_BIN_WORDS 13      // per (axis, bin): 6 minima (box lo, centroid lo), 6 maxima (box hi, centroid hi), count
#define RPTR_NODE_BIN_WORDS (3 * RPTR_SAH_BINS * RPTR_BIN_WORDS)
#define RPTR_BIN_CHUNK 2048

struct alignas(16) Prim { float lo[3]; int32_t id; float hi[3]; int32_t node; }; // node: position in the level's node list or -1

// a node of the level being split.  Bounds are kept as order-preserving integer keys so that they can be accumulated by atomics.
struct Seg {
    int32_t lo, n, link;       // triangle range; link = parent * 2 + side (-1: root)
    uint32_t kb[12];           // keys of box lo, box hi, centroid lo, centroid hi
    // written by k_split:
    int32_t axis, bin, n_left; // axis < 0: halved by position
    float cm, sc;
    int32_t child_pos[2];      // position of the children in the next level's list (-1: small / single triangle)
};
struct Small { int32_t lo, n, link; };

__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float funkey(uint32_t k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }
__device__ __forceinline__ int bin_of(float c, float cm, float sc) {
    const int q = (int)((c - cm) * sc);
    return q < 0 ? 0 : (q > RPTR_SAH_BINS - 1 ? RPTR_SAH_BINS - 1 : q);
}
__device__ __forceinline__ float half_area3(const float *lo, const float *hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ float half_area(const Box &b) { return half_area3(b.lo, b.hi); }
__device__ __forceinline__ void set_link(Node2 *nodes, int32_t *root, int32_t link, int32_t value) {
    if (link < 0) *root = value;
    else if (link & 1) nodes[link >> 1].right = value;
    else nodes[link >> 1].left = value;
}

__global__ void k_prims(const Tri *tris, int32_t n, float abs_pad, Box *boxes, Prim *prims, Seg *root_seg, int32_t root_node) {
    uint32_t mn[6], mx[6];
    for (int k = 0; k < 6; ++k) { mn[k] = 0xffffffffu; mx[k] = 0u; }
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Tri t = tris[i];
        const float v[3][3] = {{t.v0x, t.v0y, t.v0z}, {t.v0x + t.e1x, t.v0y + t.e1y, t.v0z + t.e1z}, {t.v0x + t.e2x, t.v0y + t.e2y, t.v0z + t.e2z}};
        Box b;
        Prim p;
        for (int k = 0; k < 3; ++k) {
            const float lo = fminf(v[0][k], fminf(v[1][k], v[2][k])), hi = fmaxf(v[0][k], fmaxf(v[1][k], v[2][k]));
            const float pad = 1.52587890625e-05f * fmaxf(fabsf(lo), fabsf(hi)) + abs_pad + 1e-30f; // same rule as the host builder
            b.lo[k] = lo - pad;
            b.hi[k] = hi + pad;
            p.lo[k] = b.lo[k];
            p.hi[k] = b.hi[k];
            const uint32_t kc = fkey(0.5f * (b.lo[k] + b.hi[k]));
            mn[k] = min(mn[k], fkey(b.lo[k])); mx[k] = max(mx[k], fkey(b.hi[k]));
            mn[3 + k] = min(mn[3 + k], kc); mx[3 + k] = max(mx[3 + k], kc);
        }
        p.id = i;
        p.node = root_node;
        boxes[i] = b;
        prims[i] = p;
    }
    for (int k = 0; k < 6; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = min(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = max(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; ++k) {
            atomicMin(&root_seg->kb[k], mn[k]); atomicMax(&root_seg->kb[3 + k], mx[k]);
            atomicMin(&root_seg->kb[6 + k], mn[3 + k]); atomicMax(&root_seg->kb[9 + k], mx[3 + k]);
        }
}

__global__ void k_bins_clear(uint32_t *bins, int32_t n_nodes) {
    const int64_t total = (int64_t)n_nodes * RPTR_NODE_BIN_WORDS;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x)
        bins[w] = (w % RPTR_BIN_WORDS) < 6 ? 0xffffffffu : 0u;
}

template <class Bins> __device__ __forceinline__ void bin_add(Bins *b, const Prim &p, const float *cmin, const float *sc) {
    float c[3];
    for (int k = 0; k < 3; ++k) c[k] = 0.5f * (p.lo[k] + p.hi[k]);
    for (int ax = 0; ax < 3; ++ax) {
        if (!(sc[ax] > 0.0f)) continue;
        uint32_t *w = b + (ax * RPTR_SAH_BINS + bin_of(c[ax], cmin[ax], sc[ax])) * RPTR_BIN_WORDS;
        for (int k = 0; k < 3; ++k) {
            atomicMin(w + k, fkey(p.lo[k])); atomicMin(w + 3 + k, fkey(c[k]));
            atomicMax(w + 6 + k, fkey(p.hi[k])); atomicMax(w + 9 + k, fkey(c[k]));
        }
        atomicAdd(w + 12, 1u);
    }
}
__device__ __forceinline__ void seg_scales(const Seg &s, float *cmin, float *sc) {
    for (int ax = 0; ax < 3; ++ax) {
        cmin[ax] = funkey(s.kb[6 + ax]);
        const float ext = funkey(s.kb[9 + ax]) - cmin[ax];
        sc[ax] = ext > 0.0f ? (float)RPTR_SAH_BINS / ext : 0.0f;
    }
}

// one block per chunk of RPTR_BIN_CHUNK consecutive triangles.  A chunk inside one node (the rule near the root, where all the
// triangles meet in a few hundred counters) is binned in shared memory and flushed once.
__global__ void __launch_bounds__(256) k_bin(const Prim *prims, int32_t n, const Seg *segs, uint32_t *bins) {
    __shared__ uint32_t s_bins[RPTR_NODE_BIN_WORDS];
    const int32_t base = blockIdx.x * RPTR_BIN_CHUNK;
    const int32_t last = min(base + RPTR_BIN_CHUNK, n) - 1;
    const int32_t first_node = prims[base].node;
    const bool uniform = first_node >= 0 && prims[last].node == first_node; // node ranges are contiguous
    if (uniform) {
        for (int w = threadIdx.x; w < RPTR_NODE_BIN_WORDS; w += 256) s_bins[w] = (w % RPTR_BIN_WORDS) < 6 ? 0xffffffffu : 0u;
        __syncthreads();
        float cmin[3], sc[3];
        seg_scales(segs[first_node], cmin, sc);
        for (int32_t i = base + threadIdx.x; i <= last; i += 256) bin_add(s_bins, prims[i], cmin, sc);
        __syncthreads();
        uint32_t *g = bins + (size_t)first_node * RPTR_NODE_BIN_WORDS;
        for (int w = threadIdx.x; w < RPTR_NODE_BIN_WORDS; w += 256) {
            const int k = w % RPTR_BIN_WORDS;
            const uint32_t v = s_bins[w];
            if (k < 6) { if (v != 0xffffffffu) atomicMin(g + w, v); }
            else if (k < 12) { if (v != 0u) atomicMax(g + w, v); }
            else if (v) atomicAdd(g + w, v);
        }
    } else {
        for (int32_t i = base + threadIdx.x; i <= last; i += 256) {
            const Prim p = prims[i];
            if (p.node < 0) continue;
            float cmin[3], sc[3];
            seg_scales(segs[p.node], cmin, sc);
            bin_add(bins + (size_t)p.node * RPTR_NODE_BIN_WORDS, p, cmin, sc);
        }
    }
}

// counters: [0] nodes of the next level, [1] small nodes (all levels), [2] root of the binary tree
__global__ void k_split(Seg *segs, int32_t n_segs, const uint32_t *bins, bool sah, Seg *next, Small *small, uint32_t *counters, Node2 *nodes,
                        Box *node_boxes, int32_t *level_nodes) {
    for (int32_t s_i = blockIdx.x * blockDim.x + threadIdx.x; s_i < n_segs; s_i += gridDim.x * blockDim.x) {
        Seg s = segs[s_i];
        const uint32_t *b = bins + (size_t)s_i * RPTR_NODE_BIN_WORDS;
        float cmin[3], sc[3];
        seg_scales(s, cmin, sc);
        int best_axis = -1, best_bin = -1;
        float best_cost = 1e30f;
        if (sah) {
            for (int ax = 0; ax < 3; ++ax) {
                if (!(sc[ax] > 0.0f)) continue;
                const uint32_t *ba = b + ax * RPTR_SAH_BINS * RPTR_BIN_WORDS;
                float ra[RPTR_SAH_BINS];
                int rc[RPTR_SAH_BINS];
                uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
                int c = 0;
                for (int q = RPTR_SAH_BINS - 1; q > 0; --q) {
                    const uint32_t *w = ba + q * RPTR_BIN_WORDS;
                    for (int k = 0; k < 3; ++k) { mn[k] = min(mn[k], w[k]); mx[k] = max(mx[k], w[6 + k]); }
                    c += (int)w[12];
                    float lo[3], hi[3];
                    for (int k = 0; k < 3; ++k) { lo[k] = funkey(mn[k]); hi[k] = funkey(mx[k]); }
                    ra[q] = c ? half_area3(lo, hi) : 0.0f;
                    rc[q] = c;
                }
                for (int k = 0; k < 3; ++k) { mn[k] = 0xffffffffu; mx[k] = 0u; }
                c = 0;
                for (int q = 0; q < RPTR_SAH_BINS - 1; ++q) {
                    const uint32_t *w = ba + q * RPTR_BIN_WORDS;
                    for (int k = 0; k < 3; ++k) { mn[k] = min(mn[k], w[k]); mx[k] = max(mx[k], w[6 + k]); }
                    c += (int)w[12];
                    if (c == 0 || rc[q + 1] == 0) continue;
                    float lo[3], hi[3];
                    for (int k = 0; k < 3; ++k) { lo[k] = funkey(mn[k]); hi[k] = funkey(mx[k]); }
                    const float cost = half_area3(lo, hi) * (float)c + ra[q + 1] * (float)rc[q + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = q; }
                }
            }
        }
        Seg child[2];
        for (int side = 0; side < 2; ++side)
            for (int k = 0; k < 12; ++k) child[side].kb[k] = (k < 3 || (k >= 6 && k < 9)) ? 0xffffffffu : 0u;
        if (best_axis < 0) { // coincident centroids or past the depth limit: halve by position, k_scatter accumulates the bounds
            s.axis = -1; s.bin = 0; s.cm = 0.0f; s.sc = 0.0f;
            s.n_left = s.n / 2;
        } else {
            s.axis = best_axis; s.bin = best_bin; s.cm = cmin[best_axis]; s.sc = sc[best_axis];
            int c = 0;
            for (int q = 0; q < RPTR_SAH_BINS; ++q) {
                const uint32_t *w = b + (best_axis * RPTR_SAH_BINS + q) * RPTR_BIN_WORDS;
                if (w[12] == 0u) continue;
                Seg &ch = child[q <= best_bin ? 0 : 1];
                if (q <= best_bin) c += (int)w[12];
                for (int k = 0; k < 3; ++k) {
                    ch.kb[k] = min(ch.kb[k], w[k]); ch.kb[3 + k] = max(ch.kb[3 + k], w[6 + k]);
                    ch.kb[6 + k] = min(ch.kb[6 + k], w[3 + k]); ch.kb[9 + k] = max(ch.kb[9 + k], w[9 + k]);
                }
            }
            s.n_left = c;
        }
        const int32_t idx = s.lo + s.n_left - 1;
        level_nodes[s_i] = idx;
        Box nb;
        for (int k = 0; k < 3; ++k) { nb.lo[k] = funkey(s.kb[k]); nb.hi[k] = funkey(s.kb[3 + k]); }
        node_boxes[idx] = nb;
        set_link(nodes, (int32_t *)(counters + 2), s.link, idx);
        for (int side = 0; side < 2; ++side) {
            const int32_t clo = side ? s.lo + s.n_left : s.lo, cn = side ? s.n - s.n_left : s.n_left, link = idx * 2 + side;
            s.child_pos[side] = -1;
            if (cn == 1) {
                set_link(nodes, nullptr, link, ~clo);
            } else if (cn <= RPTR_SAH_SMALL) {
                small[atomicAdd(counters + 1, 1u)] = Small{clo, cn, link};
            } else {
                const int32_t pos = (int32_t)atomicAdd(counters + 0, 1u);
                child[side].lo = clo; child[side].n = cn; child[side].link = link;
                next[pos] = child[side];
                s.child_pos[side] = pos;
            }
        }
        segs[s_i] = s;
    }
}

__device__ __forceinline__ int side_of(const Prim &p, int32_t i, const Seg &s) {
    if (s.axis < 0) return i - s.lo < s.n_left ? 0 : 1;
    return bin_of(0.5f * (p.lo[s.axis] + p.hi[s.axis]), s.cm, s.sc) <= s.bin ? 0 : 1;
}

// exclusive scan of "triangle i goes to the left child of its node" over the whole array, in three steps
#define RPTR_SCAN_BLOCK 256
#define RPTR_SCAN_ITEMS 4
__global__ void __launch_bounds__(RPTR_SCAN_BLOCK) k_side_sums(const Prim *prims, int32_t n, const Seg *segs, uint32_t *block_sums) {
    __shared__ uint32_t s_warp[RPTR_SCAN_BLOCK / 32];
    const int32_t base = blockIdx.x * RPTR_SCAN_BLOCK * RPTR_SCAN_ITEMS + threadIdx.x * RPTR_SCAN_ITEMS;
    uint32_t v = 0;
    for (int k = 0; k < RPTR_SCAN_ITEMS; ++k)
        if (base + k < n) {
            const Prim p = prims[base + k];
            if (p.node >= 0 && side_of(p, base + k, segs[p.node]) == 0) v++;
        }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < RPTR_SCAN_BLOCK / 32; ++w) t += s_warp[w];
        block_sums[blockIdx.x] = t;
    }
}
__global__ void k_scan_sums(uint32_t *block_sums, int32_t nb) { // one warp; in place, exclusive
    uint32_t carry = 0;
    for (int32_t b0 = 0; b0 < nb; b0 += 32) {
        const int32_t i = b0 + (int32_t)threadIdx.x;
        const uint32_t v = i < nb ? block_sums[i] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)threadIdx.x >= o) inc += t;
        }
        if (i < nb) block_sums[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}
__global__ void __launch_bounds__(RPTR_SCAN_BLOCK) k_scan_write(const Prim *prims, int32_t n, const Seg *segs, const uint32_t *block_sums, uint32_t *scan) {
    __shared__ uint32_t s_warp[RPTR_SCAN_BLOCK / 32];
    const int32_t base = blockIdx.x * RPTR_SCAN_BLOCK * RPTR_SCAN_ITEMS + threadIdx.x * RPTR_SCAN_ITEMS;
    uint32_t f[RPTR_SCAN_ITEMS], v = 0;
    for (int k = 0; k < RPTR_SCAN_ITEMS; ++k) {
        f[k] = 0;
        if (base + k < n) {
            const Prim p = prims[base + k];
            if (p.node >= 0 && side_of(p, base + k, segs[p.node]) == 0) f[k] = 1;
        }
        v += f[k];
    }
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)(threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t run = block_sums[blockIdx.x] + inc - v;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) run += s_warp[w];
    for (int k = 0; k < RPTR_SCAN_ITEMS; ++k)
        if (base + k < n) { scan[base + k] = run; run += f[k]; }
}

__global__ void k_scatter(const Prim *in, Prim *out, int32_t n, const Seg *segs, const uint32_t *scan, Seg *next) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Prim p = in[i];
        if (p.node < 0) { out[i] = p; continue; }
        const Seg &s = segs[p.node];
        const int side = side_of(p, i, s);
        const int32_t rank_left = (int32_t)(scan[i] - scan[s.lo]);
        const int32_t dest = side == 0 ? s.lo + rank_left : s.lo + s.n_left + (i - s.lo - rank_left);
        const int32_t axis = s.axis;
        p.node = s.child_pos[side];
        out[dest] = p;
        if (axis < 0 && p.node >= 0) {
            uint32_t *kb = next[p.node].kb;
            for (int k = 0; k < 3; ++k) {
                const uint32_t kc = fkey(0.5f * (p.lo[k] + p.hi[k]));
                atomicMin(kb + k, fkey(p.lo[k])); atomicMax(kb + 3 + k, fkey(p.hi[k]));
                atomicMin(kb + 6 + k, kc); atomicMax(kb + 9 + k, kc);
            }
        }
    }
}

// SAH-optimal collapse by dynamic programming (the same recurrence as collapse() of rptr_host.cpp, after Ylitie et al. 2017,
// section 3): cost[i - 1] = cheapest representation of the node's sub-tree as a forest of at most i roots, a root being a triangle
// slot (area x c_tri) or a wide node (area x c_node + the best distribution of <= 8 roots over the two sub-trees).
struct Dp { float cost[7]; uint8_t split[8]; }; // split[i - 1]: 0 = as for i - 1 roots (i == 1: one root); k > 0 = k roots left; split[7]: the wide node's own split
#define RPTR_COLLAPSE_C_NODE 1.0f
#define RPTR_COLLAPSE_C_TRI 0.6f
__device__ __forceinline__ float dp_cost(const Dp *dp, const Box *boxes, const uint32_t *sorted, int32_t id, int i) {
    if (id < 0) return half_area(boxes[sorted[~id]]) * RPTR_COLLAPSE_C_TRI; // a triangle is one root whatever the budget
    return dp[id].cost[i - 1];
}
__device__ void dp_node(int32_t idx, const Node2 *nodes, const Box *node_boxes, const Box *boxes, const uint32_t *sorted, Dp *dp) {
    const int32_t l = nodes[idx].left, r = nodes[idx].right;
    float lc[7], rc[7];
    for (int i = 1; i <= 7; ++i) { lc[i - 1] = dp_cost(dp, boxes, sorted, l, i); rc[i - 1] = dp_cost(dp, boxes, sorted, r, i); }
    Dp e;
    float wide = 3.0e38f;
    int wide_k = 1;
    for (int k = 1; k <= 7; ++k) {
        const float c = lc[k - 1] + rc[8 - k - 1];
        if (c < wide) { wide = c; wide_k = k; }
    }
    e.split[7] = (uint8_t)wide_k;
    e.cost[0] = half_area(node_boxes[idx]) * RPTR_COLLAPSE_C_NODE + wide;
    e.split[0] = 0;
    for (int i = 2; i <= 7; ++i) {
        float best = e.cost[i - 2];
        int best_k = 0;
        for (int k = 1; k < i; ++k) {
            const float c = lc[k - 1] + rc[i - k - 1];
            if (c < best) { best = c; best_k = k; }
        }
        e.cost[i - 1] = best;
        e.split[i - 1] = (uint8_t)best_k;
    }
    dp[idx] = e;
}
__global__ void k_collapse_dp(const int32_t *level_nodes, int32_t count, const Node2 *nodes, const Box *node_boxes, const Box *boxes,
                              const uint32_t *sorted, Dp *dp) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        dp_node(level_nodes[i], nodes, node_boxes, boxes, sorted, dp);
}

__global__ void k_order(const Prim *prims, int32_t n, uint32_t *sorted) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) sorted[i] = (uint32_t)prims[i].id;
}

// one thread per node of 2..8 triangles: the exact sweep of the host builder (every boundary of the centroid order of every axis,
// ties by triangle index), recursively down to single triangles, then the collapse costs of the nodes it made, children first
__global__ void k_sweep(const Small *small, int32_t n_small, Prim *prims, uint32_t *sorted, Node2 *nodes, Box *node_boxes, const Box *boxes, Dp *dp,
                        int32_t *root) {
    for (int32_t s_i = blockIdx.x * blockDim.x + threadIdx.x; s_i < n_small; s_i += gridDim.x * blockDim.x) {
        const Small sm = small[s_i];
        Prim P[RPTR_SAH_SMALL], tmp[RPTR_SAH_SMALL];
        for (int i = 0; i < sm.n; ++i) P[i] = prims[sm.lo + i];
        int st_a[RPTR_SAH_SMALL], st_b[RPTR_SAH_SMALL], st_link[RPTR_SAH_SMALL];
        int32_t made[RPTR_SAH_SMALL];
        int sp = 0, n_made = 0;
        st_a[sp] = 0; st_b[sp] = sm.n; st_link[sp++] = sm.link;
        while (sp > 0) {
            const int a = st_a[--sp], b = st_b[sp], link = st_link[sp], m = b - a;
            if (m == 1) { set_link(nodes, root, link, ~(sm.lo + a)); continue; }
            float best = 1e30f;
            int bsplit = 1;
            int border[RPTR_SAH_SMALL];
            Box nb;
            for (int ax = 0; ax < 3; ++ax) {
                int o[RPTR_SAH_SMALL];
                float key[RPTR_SAH_SMALL];
                for (int i = 0; i < m; ++i) { // insertion sort by (centroid, triangle index)
                    const float c = 0.5f * (P[a + i].lo[ax] + P[a + i].hi[ax]);
                    int j = i;
                    while (j > 0 && (key[j - 1] > c || (key[j - 1] == c && P[o[j - 1]].id > P[a + i].id))) { key[j] = key[j - 1]; o[j] = o[j - 1]; --j; }
                    key[j] = c; o[j] = a + i;
                }
                float ra[RPTR_SAH_SMALL];
                float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
                for (int i = m - 1; i > 0; --i) {
                    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], P[o[i]].lo[k]); mx[k] = fmaxf(mx[k], P[o[i]].hi[k]); }
                    ra[i] = half_area3(mn, mx);
                }
                for (int k = 0; k < 3; ++k) { nb.lo[k] = fminf(mn[k], P[o[0]].lo[k]); nb.hi[k] = fmaxf(mx[k], P[o[0]].hi[k]); }
                for (int k = 0; k < 3; ++k) { mn[k] = 1e30f; mx[k] = -1e30f; }
                bool better = false;
                for (int i = 0; i < m - 1; ++i) {
                    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], P[o[i]].lo[k]); mx[k] = fmaxf(mx[k], P[o[i]].hi[k]); }
                    const float cost = half_area3(mn, mx) * (float)(i + 1) + ra[i + 1] * (float)(m - i - 1);
                    if (cost < best) { best = cost; bsplit = i + 1; better = true; }
                }
                if (better || ax == 0)
                    for (int i = 0; i < m; ++i) border[i] = o[i];
            }
            for (int i = 0; i < m; ++i) tmp[i] = P[border[i]];
            for (int i = 0; i < m; ++i) P[a + i] = tmp[i];
            const int mid = a + bsplit;
            const int32_t idx = sm.lo + mid - 1;
            node_boxes[idx] = nb;
            set_link(nodes, root, link, idx);
            made[n_made++] = idx;
            st_a[sp] = mid; st_b[sp] = b; st_link[sp++] = idx * 2 + 1;
            st_a[sp] = a; st_b[sp] = mid; st_link[sp++] = idx * 2;
        }
        for (int i = 0; i < sm.n; ++i) {
            P[i].node = -1;
            prims[sm.lo + i] = P[i];
            sorted[sm.lo + i] = (uint32_t)P[i].id;
        }
        for (int i = n_made - 1; i >= 0; --i) dp_node(made[i], nodes, node_boxes, boxes, sorted, dp);
    }
}

// one wide node per queue entry: the roots the dynamic programme chose for its two sub-trees, assigns slots, claims consecutive node
// indices of the next level for its inner children and consecutive records of the leaf-order triangle array for its triangles
__global__ void k_collapse(const int32_t *cur, int32_t cur_count, int32_t level_base, int32_t next_base, int32_t *next, uint32_t *next_count,
                           uint32_t *tri_count, const Node2 *nodes, const Dp *dp, const Box *node_boxes, const Box *boxes, const uint32_t *sorted,
                           const Tri *tris, Tri *leaf_tris, BvhNode *out, int32_t n) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cur_count; i += gridDim.x * blockDim.x) {
        int32_t kids[RPTR_BVH_WIDTH]; // >= 0: binary inner node, < 0: ~(position in the sorted triangle array)
        int nk = 0;
        const int32_t root = cur[i];
        if (n == 1) {
            kids[nk++] = ~0;
        } else { // expand(left, k) + expand(right, 8 - k), iteratively: (sub-tree, budget) pairs
            int32_t st_node[RPTR_BVH_WIDTH];
            int st_budget[RPTR_BVH_WIDTH];
            int sp = 0;
            const int k0 = dp[root].split[7];
            st_node[sp] = nodes[root].right; st_budget[sp++] = 8 - k0;
            st_node[sp] = nodes[root].left; st_budget[sp++] = k0;
            while (sp > 0) {
                const int32_t m = st_node[--sp];
                int i = st_budget[sp];
                if (m < 0) { kids[nk++] = m; continue; }
                while (i > 1 && dp[m].split[i - 1] == 0) --i;
                if (i == 1) { kids[nk++] = m; continue; }
                const int k = dp[m].split[i - 1];
                st_node[sp] = nodes[m].right; st_budget[sp++] = i - k;
                st_node[sp] = nodes[m].left; st_budget[sp++] = k;
            }
        }
        float klo[RPTR_BVH_WIDTH][3], khi[RPTR_BVH_WIDTH][3];
        int n_inner = 0, n_tri = 0;
        for (int k = 0; k < nk; ++k) {
            const Box b = kids[k] < 0 ? boxes[sorted[~kids[k]]] : node_boxes[kids[k]];
            for (int a = 0; a < 3; ++a) { klo[k][a] = b.lo[a]; khi[k][a] = b.hi[a]; }
            if (kids[k] < 0) n_tri++;
            else n_inner++;
        }
        int slot_of[RPTR_BVH_WIDTH], kid_in[RPTR_BVH_WIDTH];
        assign_slots(nk, klo, khi, slot_of);
        for (int sl = 0; sl < RPTR_BVH_WIDTH; ++sl) kid_in[sl] = -1;
        for (int k = 0; k < nk; ++k) kid_in[slot_of[k]] = k;
        const uint32_t inner_pos = n_inner ? atomicAdd(next_count, (uint32_t)n_inner) : 0u;
        const uint32_t tri_pos = n_tri ? atomicAdd(tri_count, (uint32_t)n_tri) : 0u;
        float slo[RPTR_BVH_WIDTH][3], shi[RPTR_BVH_WIDTH][3];
        int kind[RPTR_BVH_WIDTH];
        uint32_t ri = 0, rt = 0;
        for (int sl = 0; sl < RPTR_BVH_WIDTH; ++sl) {
            const int k = kid_in[sl];
            kind[sl] = 0;
            for (int a = 0; a < 3; ++a) { slo[sl][a] = 0.0f; shi[sl][a] = 0.0f; }
            if (k < 0) continue;
            for (int a = 0; a < 3; ++a) { slo[sl][a] = klo[k][a]; shi[sl][a] = khi[k][a]; }
            if (kids[k] >= 0) {
                kind[sl] = 1;
                next[inner_pos + ri++] = kids[k];
            } else {
                kind[sl] = 2;
                leaf_tris[tri_pos + rt++] = tris[sorted[~kids[k]]];
            }
        }
        out[level_base + i] = encode_node(slo, shi, kind, next_base + (int32_t)inner_pos, (int32_t)tri_pos);
    }
}

__global__ void k_top_planes(const BvhNode *nodes, int32_t top_k, float4 *planes) {
    // one thread per 16-byte word: word w of node i goes to plane w, slot i
    for (int32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < top_k * RPTR_NODE_WORDS; t += gridDim.x * blockDim.x) {
        const int32_t i = t / RPTR_NODE_WORDS, w = t % RPTR_NODE_WORDS;
        planes[w * RPTR_TOP_NODES_MAX + i] = reinterpret_cast<const float4 *>(nodes + i)[w];
    }
}

#define CU_OK(call)                                         \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) { err = cudaGetErrorString(e_); goto fail; } \
    } while (0)

} // namespace

bool build_bvh_device(const std::vector<Tri> &tris, float extent, const float *cmin, const float *cmax, cudaStream_t stream, int num_sms,
                      DeviceBvh &out, std::string &error) {
    (void)cmin;
    (void)cmax;
    const int32_t n = (int32_t)tris.size();
    out = DeviceBvh();
    const char *err = nullptr;
    Tri *d_tris = nullptr, *d_leaf = nullptr;
    Box *d_boxes = nullptr, *d_nboxes = nullptr;
    Prim *d_prims[2] = {nullptr, nullptr};
    Seg *d_segs[2] = {nullptr, nullptr};
    Small *d_small = nullptr;
    uint32_t *d_bins = nullptr, *d_scan = nullptr, *d_block_sums = nullptr, *d_counters = nullptr;
    uint32_t *d_sorted = nullptr, *d_count = nullptr, *d_tri_count = nullptr;
    int32_t *d_level_nodes = nullptr;
    Node2 *d_nodes2 = nullptr;
    int32_t *d_q[2] = {nullptr, nullptr};
    Dp *d_dp = nullptr;
    int32_t root = 0;
    int cur = 0;
    BvhNode *d_out = nullptr;
    float4 *d_top = nullptr;
    std::vector<int32_t> level_base, level_count;
    const int grid = num_sms * 8;
    const size_t max_nodes = (size_t)(n > 1 ? n : 1); // a wide node has >= 2 children, so there are < n of them
    const int32_t max_active = n / (RPTR_SAH_SMALL + 1) + 1; // nodes of one level that are still split by binning
    if (n == 0) return true;
    CU_OK(cudaMalloc(&d_tris, sizeof(Tri) * n));
    CU_OK(cudaMalloc(&d_leaf, sizeof(Tri) * n));
    CU_OK(cudaMalloc(&d_boxes, sizeof(Box) * n));
    CU_OK(cudaMalloc(&d_nboxes, sizeof(Box) * n));
    CU_OK(cudaMalloc(&d_sorted, 4 * (size_t)n));
    CU_OK(cudaMalloc(&d_count, 4));
    CU_OK(cudaMalloc(&d_tri_count, 4));
    CU_OK(cudaMemsetAsync(d_tri_count, 0, 4, stream));
    CU_OK(cudaMalloc(&d_nodes2, sizeof(Node2) * n));
    CU_OK(cudaMalloc(&d_q[0], 4 * (size_t)n));
    CU_OK(cudaMalloc(&d_q[1], 4 * (size_t)n));
    CU_OK(cudaMalloc(&d_out, sizeof(BvhNode) * max_nodes));
    CU_OK(cudaMalloc(&d_prims[0], sizeof(Prim) * n));
    CU_OK(cudaMalloc(&d_prims[1], sizeof(Prim) * n));
    CU_OK(cudaMalloc(&d_segs[0], sizeof(Seg) * max_active));
    CU_OK(cudaMalloc(&d_segs[1], sizeof(Seg) * max_active));
    CU_OK(cudaMalloc(&d_small, sizeof(Small) * ((size_t)n / 2 + 1)));
    CU_OK(cudaMalloc(&d_counters, 16));
    CU_OK(cudaMemsetAsync(d_counters, 0, 16, stream));
    CU_OK(cudaMalloc(&d_dp, sizeof(Dp) * (size_t)n));
    CU_OK(cudaMalloc(&d_level_nodes, 4 * (size_t)n));
    CU_OK(cudaMemcpyAsync(d_tris, tris.data(), sizeof(Tri) * n, cudaMemcpyHostToDevice, stream));
    {
        // the root: a node of the first level, or -- up to eight triangles -- the only small node
        Seg root_seg = Seg();
        root_seg.lo = 0; root_seg.n = n; root_seg.link = -1;
        for (int k = 0; k < 12; ++k) root_seg.kb[k] = (k < 3 || (k >= 6 && k < 9)) ? 0xffffffffu : 0u;
        CU_OK(cudaMemcpyAsync(d_segs[0], &root_seg, sizeof(Seg), cudaMemcpyHostToDevice, stream));
        const bool binned = n > RPTR_SAH_SMALL;
        const float abs_pad = 7.62939453125e-06f * extent;
        k_prims<<<grid, 256, 0, stream>>>(d_tris, n, abs_pad, d_boxes, d_prims[0], d_segs[0], binned ? 0 : -1);
        CU_OK(cudaStreamSynchronize(stream)); // `root_seg` is a stack variable
        int32_t n_active = binned ? 1 : 0, n_small = 0, nodes_so_far = 0;
        if (!binned && n > 1) {
            const Small sm = Small{0, n, -1};
            const uint32_t one[2] = {0u, 1u};
            CU_OK(cudaMemcpyAsync(d_small, &sm, sizeof(Small), cudaMemcpyHostToDevice, stream));
            CU_OK(cudaMemcpyAsync(d_counters, one, 8, cudaMemcpyHostToDevice, stream));
            CU_OK(cudaStreamSynchronize(stream));
            n_small = 1;
        }
        if (n_active) {
            CU_OK(cudaMalloc(&d_bins, 4 * (size_t)max_active * RPTR_NODE_BIN_WORDS));
            CU_OK(cudaMalloc(&d_scan, 4 * (size_t)n));
            const int32_t nb = (n + RPTR_SCAN_BLOCK * RPTR_SCAN_ITEMS - 1) / (RPTR_SCAN_BLOCK * RPTR_SCAN_ITEMS);
            CU_OK(cudaMalloc(&d_block_sums, 4 * (size_t)nb));
            int sg = 0;
            for (int level = 0; n_active > 0; ++level) {
                if (level >= RPTR_SAH_MAX_LEVELS) { err = "device builder: the tree does not close"; goto fail; }
                const int ng = n_active < 256 * grid ? (n_active + 255) / 256 : grid;
                k_bins_clear<<<grid, 256, 0, stream>>>(d_bins, n_active);
                k_bin<<<(n + RPTR_BIN_CHUNK - 1) / RPTR_BIN_CHUNK, 256, 0, stream>>>(d_prims[cur], n, d_segs[sg], d_bins);
                CU_OK(cudaMemsetAsync(d_counters, 0, 4, stream));
                k_split<<<ng, 256, 0, stream>>>(d_segs[sg], n_active, d_bins, level < RPTR_SAH_DEPTH_LIMIT, d_segs[sg ^ 1], d_small, d_counters, d_nodes2,
                                                d_nboxes, d_level_nodes + nodes_so_far);
                k_side_sums<<<nb, RPTR_SCAN_BLOCK, 0, stream>>>(d_prims[cur], n, d_segs[sg], d_block_sums);
                k_scan_sums<<<1, 32, 0, stream>>>(d_block_sums, nb);
                k_scan_write<<<nb, RPTR_SCAN_BLOCK, 0, stream>>>(d_prims[cur], n, d_segs[sg], d_block_sums, d_scan);
                k_scatter<<<grid, 256, 0, stream>>>(d_prims[cur], d_prims[cur ^ 1], n, d_segs[sg], d_scan, d_segs[sg ^ 1]);
                uint32_t counters[2] = {0u, 0u};
                CU_OK(cudaMemcpyAsync(counters, d_counters, 8, cudaMemcpyDeviceToHost, stream));
                CU_OK(cudaStreamSynchronize(stream));
                level_base.push_back(nodes_so_far);
                level_count.push_back(n_active);
                nodes_so_far += n_active;
                if ((int64_t)counters[0] > max_active || (int64_t)counters[1] > n / 2 + 1) { err = "device builder: inconsistent level"; goto fail; }
                n_active = (int32_t)counters[0];
                n_small = (int32_t)counters[1];
                cur ^= 1;
                sg ^= 1;
            }
        }
        k_order<<<grid, 256, 0, stream>>>(d_prims[cur], n, d_sorted);
        if (n_small > 0)
            k_sweep<<<(n_small + 127) / 128 < grid ? (n_small + 127) / 128 : grid, 128, 0, stream>>>(d_small, n_small, d_prims[cur], d_sorted, d_nodes2, d_nboxes,
                                                                                                     d_boxes, d_dp, (int32_t *)(d_counters + 2));
        for (size_t l = level_base.size(); l-- > 0;) {
            const int32_t c = level_count[l];
            k_collapse_dp<<<(c + 127) / 128 < grid ? (c + 127) / 128 : grid, 128, 0, stream>>>(d_level_nodes + level_base[l], c, d_nodes2, d_nboxes, d_boxes,
                                                                                              d_sorted, d_dp);
        }
        if (n > 1) {
            CU_OK(cudaMemcpyAsync(&root, d_counters + 2, 4, cudaMemcpyDeviceToHost, stream));
            CU_OK(cudaStreamSynchronize(stream));
        }
    }
    {
        // breadth-first collapse, one launch per level
        int32_t cur_count = 1, level_base_w = 0, depth = 0;
        CU_OK(cudaMemcpyAsync(d_q[0], &root, 4, cudaMemcpyHostToDevice, stream));
        CU_OK(cudaStreamSynchronize(stream)); // `root` is a stack variable
        int q = 0;
        while (cur_count > 0) {
            if (++depth > RPTR_MAX_BVH_DEPTH) { err = "device builder: tree deeper than the traversal stack allows"; goto fail; }
            CU_OK(cudaMemsetAsync(d_count, 0, 4, stream));
            const int32_t next_base = level_base_w + cur_count;
            if ((size_t)next_base > max_nodes) { err = "device builder: node overflow"; goto fail; }
            k_collapse<<<grid, 128, 0, stream>>>(d_q[q], cur_count, level_base_w, next_base, d_q[q ^ 1], d_count, d_tri_count, d_nodes2, d_dp, d_nboxes,
                                                 d_boxes, d_sorted, d_tris, d_leaf, d_out, n);
            uint32_t next_count = 0;
            CU_OK(cudaMemcpyAsync(&next_count, d_count, 4, cudaMemcpyDeviceToHost, stream));
            CU_OK(cudaStreamSynchronize(stream));
            level_base_w = next_base;
            cur_count = (int32_t)next_count;
            q ^= 1;
        }
        out.n_nodes = level_base_w;
        out.depth = depth;
    }
    out.top_k = out.n_nodes < RPTR_TOP_NODES_MAX ? out.n_nodes : RPTR_TOP_NODES_MAX;
    CU_OK(cudaMalloc(&d_top, sizeof(float4) * RPTR_NODE_WORDS * RPTR_TOP_NODES_MAX));
    CU_OK(cudaMemsetAsync(d_top, 0, sizeof(float4) * RPTR_NODE_WORDS * RPTR_TOP_NODES_MAX, stream));
    if (out.top_k > 0) k_top_planes<<<64, 256, 0, stream>>>(d_out, out.top_k, d_top);
    CU_OK(cudaStreamSynchronize(stream));
    CU_OK(cudaGetLastError());
    out.nodes = d_out;
    out.tris = d_leaf;
    out.top = d_top;
    out.n_tris = n;
    d_out = nullptr;
    d_leaf = nullptr;
    d_top = nullptr;
fail:
    cudaFree(d_tris); cudaFree(d_leaf); cudaFree(d_boxes); cudaFree(d_nboxes); cudaFree(d_sorted); cudaFree(d_count); cudaFree(d_tri_count);
    cudaFree(d_nodes2); cudaFree(d_q[0]); cudaFree(d_q[1]); cudaFree(d_prims[0]); cudaFree(d_prims[1]); cudaFree(d_segs[0]); cudaFree(d_segs[1]);
    cudaFree(d_small); cudaFree(d_bins); cudaFree(d_scan); cudaFree(d_block_sums); cudaFree(d_counters); cudaFree(d_level_nodes); cudaFree(d_dp);
    cudaFree(d_out); cudaFree(d_top);
    if (err) {
        error = err;
        return false;
    }
    return true;
}

} // namespace rp
