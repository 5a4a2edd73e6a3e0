// oracle/ref_shim/ref_rq.cpp -- TEST INFRASTRUCTURE.
// The body of vulkan/rt_intersect.comp:main (RaytraceBackend::trace_ray, RQ_CLOSEST) executed as C++ over a scripted ray query:
// the Makefile cuts the statements from the fetch of the query to the store of the result out of the file where it lies
// (oracle/_ref/gen/rq_body.inc, deleted after the compile); this file answers the GL_EXT_ray_query calls with "no hit" or one
// committed triangle hit and supplies the two buffers.  What runs from the reference: t_min, the skip of queries with
// mode < 0, the traversal flags and the packing of the result.
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "../../include/rptr_types.h"

namespace refrq {
using namespace glm;
typedef unsigned int uint;
#include "rendering/language.hpp"
#define RAY_EPSILON 0.000005f // vulkan/gpu_params.glsl:27-29

struct RenderRayQuery { vec3 origin; int mode_or_data; vec3 dir; float t_max; }; // librender/render_params.glsl.h:165-170
static RenderRayQuery ray_queries[1];
static vec4 ray_results[1];
static struct { int hit; vec2 bary; int custom_index, geometry_index, prim; } g_script;
static struct { vec3 o, d; float tmin, tmax; uint flags; int inits; } g_init;
struct rayQueryEXT { int state; };
static int scene;
enum { gl_RayFlagsOpaqueEXT = 1, gl_RayQueryCommittedIntersectionTriangleEXT = 1 };
inline void rayQueryInitializeEXT(rayQueryEXT &q, int, uint flags, uint, vec3 o, float tmin, vec3 d, float tmax) {
    g_init.o = o; g_init.d = d; g_init.tmin = tmin; g_init.tmax = tmax; g_init.flags = flags; ++g_init.inits;
    q.state = 0;
}
inline bool rayQueryProceedEXT(rayQueryEXT &) { return false; } // all geometry opaque: the traversal commits by itself
inline void rayQueryConfirmIntersectionEXT(rayQueryEXT &) {}
inline uint rayQueryGetIntersectionTypeEXT(rayQueryEXT &, bool) { return g_script.hit ? 1u : 0u; }
inline vec2 rayQueryGetIntersectionBarycentricsEXT(rayQueryEXT &, bool) { return g_script.bary; }
inline int rayQueryGetIntersectionInstanceCustomIndexEXT(rayQueryEXT &, bool) { return g_script.custom_index; }
inline int rayQueryGetIntersectionGeometryIndexEXT(rayQueryEXT &, bool) { return g_script.geometry_index; }
inline int rayQueryGetIntersectionPrimitiveIndexEXT(rayQueryEXT &, bool) { return g_script.prim; }
inline vec2 intBitsToFloat(ivec2 v) { vec2 r; std::memcpy(&r.x, &v.x, 4); std::memcpy(&r.y, &v.y, 4); return r; }

static void query_body(uint query_id) {
#include "gen/rq_body.inc"
}
} // namespace refrq

extern "C" {
// q = one RenderRayQuery (8 floats); result is read-modify-write (a skipped query leaves it alone).
// hit script: hit (0/1), bary(2), instance custom index, geometry index, primitive index.
// info: [0] ray queries started, [1] t_min, [2] t_max, [3] traversal flags
void ref_ray_query(const float *q, int32_t hit, const float *bary, int32_t custom_index, int32_t geometry_index, int32_t prim, float *result, float *info) {
    using namespace refrq;
    ray_queries[0].origin = glm::vec3(q[0], q[1], q[2]);
    std::memcpy(&ray_queries[0].mode_or_data, &q[3], 4);
    ray_queries[0].dir = glm::vec3(q[4], q[5], q[6]);
    ray_queries[0].t_max = q[7];
    ray_results[0] = glm::vec4(result[0], result[1], result[2], result[3]);
    g_script.hit = hit; g_script.bary = glm::vec2(bary[0], bary[1]);
    g_script.custom_index = custom_index; g_script.geometry_index = geometry_index; g_script.prim = prim;
    g_init.inits = 0; g_init.tmin = 0.0f; g_init.tmax = 0.0f; g_init.flags = 0;
    query_body(0u);
    std::memcpy(result, &ray_results[0], 16);
    info[0] = (float)g_init.inits; info[1] = g_init.tmin; info[2] = g_init.tmax; info[3] = (float)g_init.flags;
}
} // extern "C"
