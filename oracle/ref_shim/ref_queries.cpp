// oracle/ref_shim/ref_queries.cpp -- TEST INFRASTRUCTURE.
// The pieces of the reference that decide how a ray query is path traced (RenderBackend::render_ray_queries), executed as C++:
//   * vulkan/setup_pixel_assignment.glsl (included whole, compute flavour: WORKGROUP_SIZE_X defined): the swizzled
//     gl_GlobalInvocationID that seeds the samplers and gl_GlobalInvocationIndex = the query id;
//   * accumulate_query of vulkan/accumulate.glsl (cut out by the Makefile into oracle/_ref/gen/accumulate_query.inc);
//   * the dispatch size of record_frame (vulkan/render_vulkan.cpp: "dispatch ray queries into a virtual screen square", cut out
//     into gen/query_dispatch.inc) and the workgroup count of ComputeRenderPipelineVulkan::dispatch_rays
//     (vulkan/render_pipeline_vulkan.cpp, gen/dispatch_rays.inc).
// This file supplies the GLSL built-ins and the buffers those pieces read.
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace refq {
using namespace glm;
typedef unsigned int uint;
#include "rendering/language.hpp"

struct BuiltinInvocationID { uint x, y, z; uvec2 xy; };
static BuiltinInvocationID gl_GlobalInvocationID; // the built-in the macro of the same name reads
static uint gl_LocalInvocationIndex;
static uvec3 gl_WorkGroupID, gl_NumWorkGroups, gl_WorkGroupSize;
static void set_builtin_invocation_id(uint x, uint y) { // before the macro of the same name exists
    gl_GlobalInvocationID.x = x; gl_GlobalInvocationID.y = y; gl_GlobalInvocationID.z = 0u;
    gl_GlobalInvocationID.xy = uvec2(x, y);
}
#define WORKGROUP_SIZE_X 32 // vulkan/CMakeLists.txt:53-55 (WORKGROUP_SIZE 32 x 16)
#include "vulkan/setup_pixel_assignment.glsl"

static vec4 ray_results[1];
inline vec4 operator/(vec4 a, uint b) { return a / float(b); } // GLSL converts the uint operand to float implicitly (spec 4.1.10)
#include "gen/accumulate_query.inc"

struct Dims { int x, y; };
static void query_dispatch(int num_rayqueries, int *out) {
    Dims dispatch_dim{0, 0};
    {
#include "gen/query_dispatch.inc"
    }
    out[0] = dispatch_dim.x; out[1] = dispatch_dim.y;
}
struct DispatchStandIn {
    glm::uvec3 workgroup_size{32u, 16u, 1u};
    glm::uvec3 dims;
    void run(int width, int height, int batch_spp) {
#include "gen/dispatch_rays.inc"
        dims = dispatch_dim;
    }
};
} // namespace refq

extern "C" {
// invocation (local index l of workgroup (wx, wy) of a dispatch of nwx x nwy workgroups) -> out = swizzled id x, y, invocation index
void ref_query_invocation(uint32_t wx, uint32_t wy, uint32_t nwx, uint32_t nwy, uint32_t l, uint32_t *out) {
    using namespace refq;
    refq::gl_WorkGroupSize = glm::uvec3(32u, 16u, 1u);
    refq::gl_WorkGroupID = glm::uvec3(wx, wy, 0u);
    refq::gl_NumWorkGroups = glm::uvec3(nwx, nwy, 1u);
    refq::gl_LocalInvocationIndex = l;
    refq::set_builtin_invocation_id(wx * 32u + l % 32u, wy * 16u + l / 32u);
    const glm::uvec3 id = gl_GlobalInvocationID; // the macro: swizzled inside the workgroup
    const uint32_t index = gl_GlobalInvocationIndex;
    out[0] = id.x; out[1] = id.y; out[2] = index;
}
// accumulate_query(0, new_result, sample_index) on ray_results[0] = result (in/out)
void ref_accumulate_query(float *result, const float *new_result, uint32_t sample_index) {
    refq::ray_results[0] = glm::vec4(result[0], result[1], result[2], result[3]);
    refq::accumulate_query(0u, glm::vec4(new_result[0], new_result[1], new_result[2], new_result[3]), sample_index);
    std::memcpy(result, &refq::ray_results[0], 16);
}
// out = dispatch_dim.x, dispatch_dim.y (invocations), workgroups x, y, z for `batch_spp` layers
void ref_query_dispatch(int32_t num_queries, int32_t batch_spp, int32_t *out) {
    refq::query_dispatch(num_queries, out);
    refq::DispatchStandIn d;
    d.run(out[0], out[1], batch_spp);
    out[2] = (int32_t)d.dims.x; out[3] = (int32_t)d.dims.y; out[4] = (int32_t)d.dims.z;
}
} // extern "C"
