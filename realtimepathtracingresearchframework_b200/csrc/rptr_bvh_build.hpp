// rptr_bvh_build.hpp -- device-side binned-SAH builder (rptr_bvh_build.cu)
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "rptr_bvh.cuh"

namespace rp {

struct DeviceBvh { // device allocations owned by the caller after a successful build (cudaFree)
    BvhNode *nodes = nullptr;
    Tri *tris = nullptr;     // leaf order
    float4 *top = nullptr;   // word planes of the first top_k nodes (4 * RPTR_TOP_NODES_MAX words, BvhDev::top_planes)
    int32_t n_nodes = 0, n_tris = 0, top_k = 0, depth = 0;
};

// tris: flattened world-space triangles (Tri::id = index); extent = largest |coordinate|; cmin/cmax = bounds of the box centres.
bool build_bvh_device(const std::vector<Tri> &tris, float extent, const float *cmin, const float *cmax, cudaStream_t stream, int num_sms,
                      DeviceBvh &out, std::string &error);

} // namespace rp
