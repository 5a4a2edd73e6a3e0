// rptr_bvh_build.cu -- device-side BVH builder (option "bvh_builder" = 1): Morton-order LBVH built entirely on the GPU,
// emitted directly in the 4-wide breadth-first layout the trace kernels read (rptr_bvh.cuh).
//
// Stands in for vkCmdBuildAccelerationStructuresKHR (vulkan/vulkanrt_utils.cpp:82-167,241-300): the reference contains no
// BVH algorithm of its own, so only closest-hit RESULTS have to match -- which they do for any
// conservative tree, so images rendered with this builder are bit-identical to those of the host SAH builder
// (tests/test_gpu_parity.py::test_device_lbvh_builder_gives_identical_images).
//
// Stages (all kernels below, one CUB radix sort for the Morton keys -- library plumbing of the build, not the hot path):
//   k_morton      padded triangle boxes + 63-bit Morton key of the box centre
//   sort          (key, triangle) pairs
//   k_radix_tree  Karras 2012: one thread per internal node finds its key range and split
//   k_refit       bottom-up boxes through atomic arrival counters
//   k_collapse    level by level: opens the largest child until 4 are held; sub-trees of <= RPTR_LBVH_LEAF_MAX (1) triangles become leaves
//                 (their triangles are contiguous in Morton order, so the leaf triangle array is just the sorted array)
//   k_gather_tris / k_top_planes    leaf-order triangle records, shared-memory image of the first nodes
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <vector>

#include "rptr_bvh_build.hpp"

#ifndef RPTR_LBVH_LEAF_MAX
#define RPTR_LBVH_LEAF_MAX 1 // sub-trees of at most this many triangles become leaves (1 = single-triangle leaves)
#endif

namespace rp {

namespace {

struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ uint64_t expand21(uint32_t v) { // spread 21 bits to every third bit
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const Tri *tris, int32_t n, float abs_pad, float3 cmin, float3 cscale, Box *boxes, uint64_t *keys, uint32_t *vals) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Tri t = tris[i];
        const float v[3][3] = {{t.v0x, t.v0y, t.v0z}, {t.v0x + t.e1x, t.v0y + t.e1y, t.v0z + t.e1z}, {t.v0x + t.e2x, t.v0y + t.e2y, t.v0z + t.e2z}};
        Box b;
        float c[3];
        for (int k = 0; k < 3; ++k) {
            const float lo = fminf(v[0][k], fminf(v[1][k], v[2][k])), hi = fmaxf(v[0][k], fmaxf(v[1][k], v[2][k]));
            const float pad = 1.52587890625e-05f * fmaxf(fabsf(lo), fabsf(hi)) + abs_pad + 1e-30f; // same rule as the host builder
            b.lo[k] = lo - pad;
            b.hi[k] = hi + pad;
            c[k] = 0.5f * (b.lo[k] + b.hi[k]);
        }
        boxes[i] = b;
        const uint32_t qx = (uint32_t)fminf(fmaxf((c[0] - cmin.x) * cscale.x, 0.0f), 2097151.0f);
        const uint32_t qy = (uint32_t)fminf(fmaxf((c[1] - cmin.y) * cscale.y, 0.0f), 2097151.0f);
        const uint32_t qz = (uint32_t)fminf(fmaxf((c[2] - cmin.z) * cscale.z, 0.0f), 2097151.0f);
        keys[i] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
        vals[i] = (uint32_t)i;
    }
}

// binary radix tree: internal nodes [0, n-1), leaves encoded as ~leaf_index
struct Node2 {
    int32_t left, right, parent;
    int32_t first, last; // range of sorted triangles covered
};

__device__ __forceinline__ int delta(const uint64_t *keys, int32_t n, int32_t i, int32_t j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_radix_tree(const uint64_t *keys, int32_t n, Node2 *nodes, int32_t *leaf_parent) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(keys, n, i, i - d);
        int lmax = 2;
        while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t >= 1; t >>= 1)
            if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
        const int32_t j = i + l * d;
        const int dnode = delta(keys, n, i, j);
        int s = 0;
        for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
            if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
            if (t == 1) break;
        }
        const int32_t gamma = i + s * d + min(d, 0);
        const int32_t first = min(i, j), last = max(i, j);
        Node2 nd;
        nd.left = (first == gamma) ? ~gamma : gamma;
        nd.right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
        // nodes[i].parent is written by the thread of the parent node (the root keeps the -1 of the memset)
        nodes[i].left = nd.left;
        nodes[i].right = nd.right;
        nodes[i].first = first;
        nodes[i].last = last;
        if (nd.left >= 0) nodes[nd.left].parent = i;
        else leaf_parent[gamma] = i;
        if (nd.right >= 0) nodes[nd.right].parent = i;
        else leaf_parent[gamma + 1] = i;
    }
}

__device__ __forceinline__ Box load_box_cg(const Box *p) {
    Box b;
    const float *f = reinterpret_cast<const float *>(p);
    for (int k = 0; k < 3; ++k) {
        b.lo[k] = __ldcg(f + k);
        b.hi[k] = __ldcg(f + 3 + k);
    }
    return b;
}

__global__ void k_refit(const Box *boxes, const uint32_t *sorted, int32_t n, const Node2 *nodes, const int32_t *leaf_parent, Box *node_boxes,
                        uint32_t *arrivals) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int32_t cur = leaf_parent[i];
        while (cur >= 0) {
            __threadfence();
            if (atomicAdd(&arrivals[cur], 1u) == 0u) break; // the first child to arrive stops; the second continues with both boxes ready
            const Node2 nd = nodes[cur];
            // boxes of inner children were written by other SMs during this launch: read them through L2 (ld.cg)
            const Box a = nd.left >= 0 ? load_box_cg(node_boxes + nd.left) : boxes[sorted[~nd.left]];
            const Box b = nd.right >= 0 ? load_box_cg(node_boxes + nd.right) : boxes[sorted[~nd.right]];
            Box m;
            for (int k = 0; k < 3; ++k) {
                m.lo[k] = fminf(a.lo[k], b.lo[k]);
                m.hi[k] = fmaxf(a.hi[k], b.hi[k]);
            }
            node_boxes[cur] = m;
            cur = nd.parent;
        }
    }
}

__device__ __forceinline__ float half_area(const Box &b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// one wide node per queue entry: opens the largest inner child until eight are held, assigns slots, claims consecutive node
// indices of the next level for its inner children and consecutive records of the leaf-order triangle array for its triangles
__global__ void k_collapse(const int32_t *cur, int32_t cur_count, int32_t level_base, int32_t next_base, int32_t *next, uint32_t *next_count,
                           uint32_t *tri_count, const Node2 *nodes, const Box *node_boxes, const Box *boxes, const uint32_t *sorted,
                           const Tri *tris, Tri *leaf_tris, BvhNode *out, int32_t n) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cur_count; i += gridDim.x * blockDim.x) {
        int32_t kids[RPTR_BVH_WIDTH]; // >= 0: binary inner node, < 0: ~(position in the sorted triangle array)
        int nk = 0;
        const int32_t root = cur[i];
        if (n == 1) {
            kids[nk++] = ~0;
        } else {
            kids[nk++] = nodes[root].left;
            kids[nk++] = nodes[root].right;
            while (nk < RPTR_BVH_WIDTH) {
                int pick = -1;
                float best = -1.0f;
                for (int k = 0; k < nk; ++k) {
                    if (kids[k] < 0) continue;
                    const float a = half_area(node_boxes[kids[k]]);
                    if (a > best) { best = a; pick = k; }
                }
                if (pick < 0) break;
                const int32_t open = kids[pick];
                kids[pick] = nodes[open].left;
                kids[nk++] = nodes[open].right;
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

#define CUB_OK(call)                                         \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) { err = cudaGetErrorString(e_); goto fail; } \
    } while (0)

} // namespace

bool build_bvh_device(const std::vector<Tri> &tris, float extent, const float *cmin, const float *cmax, cudaStream_t stream, int num_sms,
                      DeviceBvh &out, std::string &error) {
    const int32_t n = (int32_t)tris.size();
    out = DeviceBvh();
    const char *err = nullptr;
    Tri *d_tris = nullptr, *d_leaf = nullptr;
    Box *d_boxes = nullptr, *d_nboxes = nullptr;
    uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_vals = nullptr, *d_sorted = nullptr, *d_arrivals = nullptr, *d_count = nullptr, *d_tri_count = nullptr;
    Node2 *d_nodes2 = nullptr;
    int32_t *d_leaf_parent = nullptr, *d_q[2] = {nullptr, nullptr};
    BvhNode *d_out = nullptr;
    float4 *d_top = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    const int grid = num_sms * 8;
    const size_t max_nodes = (size_t)(n > 1 ? n : 1); // a wide node has >= 2 children, so there are < n of them
    if (n == 0) return true;
    CUB_OK(cudaMalloc(&d_tris, sizeof(Tri) * n));
    CUB_OK(cudaMalloc(&d_leaf, sizeof(Tri) * n));
    CUB_OK(cudaMalloc(&d_boxes, sizeof(Box) * n));
    CUB_OK(cudaMalloc(&d_nboxes, sizeof(Box) * n));
    CUB_OK(cudaMalloc(&d_keys, 8 * (size_t)n));
    CUB_OK(cudaMalloc(&d_keys2, 8 * (size_t)n));
    CUB_OK(cudaMalloc(&d_vals, 4 * (size_t)n));
    CUB_OK(cudaMalloc(&d_sorted, 4 * (size_t)n));
    CUB_OK(cudaMalloc(&d_arrivals, 4 * (size_t)n));
    CUB_OK(cudaMalloc(&d_count, 4));
    CUB_OK(cudaMalloc(&d_tri_count, 4));
    CUB_OK(cudaMemsetAsync(d_tri_count, 0, 4, stream));
    CUB_OK(cudaMalloc(&d_nodes2, sizeof(Node2) * n));
    CUB_OK(cudaMalloc(&d_leaf_parent, 4 * (size_t)n));
    CUB_OK(cudaMalloc(&d_q[0], 4 * (size_t)n));
    CUB_OK(cudaMalloc(&d_q[1], 4 * (size_t)n));
    CUB_OK(cudaMalloc(&d_out, sizeof(BvhNode) * max_nodes));
    CUB_OK(cudaMemcpyAsync(d_tris, tris.data(), sizeof(Tri) * n, cudaMemcpyHostToDevice, stream));
    {
        const float abs_pad = 7.62939453125e-06f * extent;
        float3 mn = make_float3(cmin[0], cmin[1], cmin[2]);
        float3 sc = make_float3(2097152.0f / fmaxf(cmax[0] - cmin[0], 1e-30f), 2097152.0f / fmaxf(cmax[1] - cmin[1], 1e-30f),
                                2097152.0f / fmaxf(cmax[2] - cmin[2], 1e-30f));
        k_morton<<<grid, 256, 0, stream>>>(d_tris, n, abs_pad, mn, sc, d_boxes, d_keys, d_vals);
    }
    CUB_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_sorted, n, 0, 63, stream));
    CUB_OK(cudaMalloc(&d_tmp, tmp_bytes));
    CUB_OK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_sorted, n, 0, 63, stream));
    CUB_OK(cudaMemsetAsync(d_arrivals, 0, 4 * (size_t)n, stream));
    CUB_OK(cudaMemsetAsync(d_nodes2, 0xff, sizeof(Node2) * n, stream)); // parent = -1 everywhere (root keeps it)
    CUB_OK(cudaMemsetAsync(d_leaf_parent, 0xff, 4 * (size_t)n, stream));
    if (n > 1) {
        k_radix_tree<<<grid, 256, 0, stream>>>(d_keys2, n, d_nodes2, d_leaf_parent);
        k_refit<<<grid, 256, 0, stream>>>(d_boxes, d_sorted, n, d_nodes2, d_leaf_parent, d_nboxes, d_arrivals);
    }
    {
        // breadth-first collapse, one launch per level
        int32_t root = 0, cur_count = 1, level_base = 0, depth = 0;
        CUB_OK(cudaMemcpyAsync(d_q[0], &root, 4, cudaMemcpyHostToDevice, stream));
        int cur = 0;
        while (cur_count > 0) {
            if (++depth > RPTR_MAX_BVH_DEPTH) { err = "device LBVH deeper than the traversal stack allows"; goto fail; }
            CUB_OK(cudaMemsetAsync(d_count, 0, 4, stream));
            const int32_t next_base = level_base + cur_count;
            if ((size_t)next_base > max_nodes) { err = "device LBVH node overflow"; goto fail; }
            k_collapse<<<grid, 128, 0, stream>>>(d_q[cur], cur_count, level_base, next_base, d_q[cur ^ 1], d_count, d_tri_count, d_nodes2, d_nboxes,
                                                 d_boxes, d_sorted, d_tris, d_leaf, d_out, n);
            uint32_t next_count = 0;
            CUB_OK(cudaMemcpyAsync(&next_count, d_count, 4, cudaMemcpyDeviceToHost, stream));
            CUB_OK(cudaStreamSynchronize(stream));
            level_base = next_base;
            cur_count = (int32_t)next_count;
            cur ^= 1;
        }
        out.n_nodes = level_base;
        out.depth = depth;
    }
    out.top_k = out.n_nodes < RPTR_TOP_NODES_MAX ? out.n_nodes : RPTR_TOP_NODES_MAX;
    CUB_OK(cudaMalloc(&d_top, sizeof(float4) * RPTR_NODE_WORDS * RPTR_TOP_NODES_MAX));
    CUB_OK(cudaMemsetAsync(d_top, 0, sizeof(float4) * RPTR_NODE_WORDS * RPTR_TOP_NODES_MAX, stream));
    if (out.top_k > 0) k_top_planes<<<64, 256, 0, stream>>>(d_out, out.top_k, d_top);
    CUB_OK(cudaStreamSynchronize(stream));
    CUB_OK(cudaGetLastError());
    out.nodes = d_out;
    out.tris = d_leaf;
    out.top = d_top;
    out.n_tris = n;
    d_out = nullptr;
    d_leaf = nullptr;
    d_top = nullptr;
fail:
    cudaFree(d_tris); cudaFree(d_leaf); cudaFree(d_boxes); cudaFree(d_nboxes); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_vals);
    cudaFree(d_sorted); cudaFree(d_arrivals); cudaFree(d_count); cudaFree(d_tri_count); cudaFree(d_nodes2); cudaFree(d_leaf_parent); cudaFree(d_q[0]); cudaFree(d_q[1]);
    cudaFree(d_out); cudaFree(d_top); cudaFree(d_tmp);
    if (err) {
        error = err;
        return false;
    }
    return true;
}

} // namespace rp
