// oracle/ref_shim/ref_shade.cpp -- TEST INFRASTRUCTURE.
// The reference's OWN per-vertex shading function, #included from where it lies under REF and executed as C++:
// rendering/mc/shade_base_material.glsl:14-96 (material unpack, emitter MIS, AOV channels, path-length cut, next-event
// estimation, glossy-only cut, BSDF sampling, bounce counting) with everything it pulls in -- shading_interface.glsl,
// rt/material_textures.glsl, mc/nee.glsl, the glTF BSDF, the LCG pointset of rendering/defaults.glsl -- assembled the way
// vulkan/pt_megakernel.glsl:22-109 does.  Supplied here: the uniform blocks it reads (render_params, scene_params), the light
// buffer, a texture unit for 1 x 1 textures and raytrace_test_visibility(), which records the shadow query and says "visible".
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "../../include/rptr_types.h"

namespace refshade {
using namespace glm;
typedef unsigned int uint;
struct Texel1x1 { vec4 v; };
static const Texel1x1 *g_textures = nullptr;
inline vec4 textureLod(const Texel1x1 &t, vec2, float) { return t.v; }
inline uint32_t floatBitsToUint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
#define SCENE_GET_TEXTURE(index) g_textures[index]
#define PREMULTIPLIED_BASE_COLOR_ALPHA

#include "rendering/language.hpp"
#include "rendering/pointsets/lcg_rng.glsl"
// GLSL evaluates the arguments of `vec4(RANDOM_FLOAT2(rng, DIM_POSITION_X), RANDOM_FLOAT2(rng, DIM_LIGHT_SEL_1))`
// (shade_base_material.glsl:61) left to right; g++ evaluates function arguments right to left.  The pointset macros are
// meant to be supplied by the includer (rendering/defaults.glsl:23-38 only fills in what is missing), so this one restores
// the GLSL order: in shader order every dimension-2 pair (light position / lobe) is drawn before its dimension-0 partner
// (light selection / direction); a dimension-0 request that arrives first draws its partner's numbers first and parks them.
#define RANDOM_STATE LCGRand
#define RANDOM_FLOAT1(state, dim) lcg_randomf(state)
static bool g_expect_dim0 = false, g_parked = false;
static vec2 g_parked_pair;
inline vec2 glsl_ordered_random_float2(LCGRand &state, int dim) {
    vec2 r;
    if (dim == 2) {
        if (g_parked) { g_parked = false; return g_parked_pair; }
        r.x = lcg_randomf(state); r.y = lcg_randomf(state);
        g_expect_dim0 = true;
        return r;
    }
    if (!g_expect_dim0) { // evaluated before its dimension-2 partner
        g_parked_pair.x = lcg_randomf(state); g_parked_pair.y = lcg_randomf(state);
        g_parked = true;
    }
    g_expect_dim0 = false;
    r.x = lcg_randomf(state); r.y = lcg_randomf(state);
    return r;
}
#define RANDOM_FLOAT2(state, dim) glsl_ordered_random_float2(state, dim)
#include "rendering/defaults.glsl"
#include "rendering/util.glsl"
#include "rendering/bsdfs/base_material.h.glsl"
#include "rendering/bsdfs/hit_point.glsl"
#include "rendering/lights/tri.glsl"

struct SceneParamsStandIn { vec3 sun_dir; float sun_cos_angle; vec4 sun_radiance; };
struct RenderParamsStandIn { int max_path_depth; int glossy_only_mode; };
static SceneParamsStandIn scene_params;
static RenderParamsStandIn render_params;
static const TriLightData *g_lights = nullptr;
static int g_num_lights = 0;
static int g_bin_size = 16;
#define SCENE_GET_LIGHT_SOURCE(light_id) decode_tri_light(g_lights[light_id])
#define SCENE_GET_LIGHT_SOURCE_COUNT() int(g_num_lights)
#define BINNED_LIGHTS_BIN_MAX_SIZE 16
#define BINNED_LIGHTS_BIN_SIZE int(g_bin_size)
#define SCENE_GET_BINNED_LIGHTS_BIN_COUNT() ((g_num_lights + (g_bin_size - 1)) / g_bin_size)
#define GLOSSY_MODE_ROUGHNESS_THRESHOLD 0.1f

static vec3 g_query_from, g_query_dir;
static float g_query_dist;
static int g_queries;

#include "rendering/rt/materials.glsl"
#include "rendering/bsdfs/gltf_bsdf.glsl"
static int (*g_visibility_cb)(void *, const float *, const float *, float) = nullptr; // set by the whole-path driver (ref_path.cpp)
static void *g_visibility_user = nullptr;
inline bool raytrace_test_visibility(const vec3 from, const vec3 dir, float dist) {
    g_query_from = from; g_query_dir = dir; g_query_dist = dist; ++g_queries;
    if (g_visibility_cb) {
        const float f[3] = {from.x, from.y, from.z}, d[3] = {dir.x, dir.y, dir.z};
        return g_visibility_cb(g_visibility_user, f, d, dist) != 0;
    }
    return true;
}
#include "rendering/rt/material_textures.glsl" // vulkan/pt_megakernel.glsl:106-109: textures, nee, then the shading function
#include "rendering/mc/nee.glsl"
#include "rendering/mc/shade_base_material.glsl"
} // namespace refshade

extern "C" {

// the whole-path driver answers the shadow queries of sample_direct_light itself; null = record the query and say "visible"
void ref_shade_set_visibility(int (*cb)(void *, const float *, const float *, float), void *user) {
    refshade::g_visibility_cb = cb;
    refshade::g_visibility_user = user;
}

// in : material, state (bounce, output_channel, prev_bounce_pdf), illum[3], throughput[3], approx solid angle of the hit triangle,
//      w_o[3], interaction (p, gn, n, v_x, v_y: 15 floats), LCG state, render params (max_path_depth, glossy_only_mode),
//      sun block (sun_dir[3], sun_cos_angle, sun_radiance[4]), binned lights + bin size
// out: [0] result, [1] bounce, [2] prev_bounce_pdf, [3..5] illum, [6..8] throughput, [9..11] w_i, [12] aux.mis_pdf,
//      [13] bits(LCG state after), [14] number of shadow queries, [15..17] query dir, [18] query dist
void ref_shade_base_material(const rptr_base_material *p, int bounce, int output_channel, float prev_bounce_pdf, const float *illum,
                             const float *throughput, float approx_sa, const float *wo, const float *ia, uint32_t rng_state, int max_path_depth,
                             int glossy_only_mode, const float *sun_dir, float sun_cos_angle, const float *sun_radiance,
                             const rptr_tri_light_data *lights, int n_lights, int bin_size, float *out) {
    using namespace refshade;
    scene_params.sun_dir = glm::vec3(sun_dir[0], sun_dir[1], sun_dir[2]);
    scene_params.sun_cos_angle = sun_cos_angle;
    scene_params.sun_radiance = glm::vec4(sun_radiance[0], sun_radiance[1], sun_radiance[2], sun_radiance[3]);
    render_params.max_path_depth = max_path_depth;
    render_params.glossy_only_mode = glossy_only_mode;
    g_lights = reinterpret_cast<const TriLightData *>(lights);
    g_num_lights = n_lights;
    g_bin_size = bin_size;
    g_queries = 0;
    g_expect_dim0 = g_parked = false;
    BaseMaterial bm;
    std::memcpy(&bm, p, sizeof(bm));
    ShadingSampleState st;
    st.bounce = bounce; st.output_channel = output_channel; st.prev_bounce_pdf = prev_bounce_pdf;
    glm::vec3 il(illum[0], illum[1], illum[2]), thr(throughput[0], throughput[1], throughput[2]);
    HitPoint lookup{glm::vec3(ia[0], ia[1], ia[2]), glm::vec2(0.0f), glm::mat2(0.0f), glm::vec3(0.0f, 0.0f, 1.0f)};
    NEESampledArea area;
    area.approx_solid_angle = approx_sa;
    area.type = 0;
    InteractionPoint hit;
    hit.p = glm::vec3(ia[0], ia[1], ia[2]);
    hit.gn = glm::vec3(ia[3], ia[4], ia[5]);
    hit.n = glm::vec3(ia[6], ia[7], ia[8]);
    hit.v_x = glm::vec3(ia[9], ia[10], ia[11]);
    hit.v_y = glm::vec3(ia[12], ia[13], ia[14]);
    hit.primitiveId = 0; hit.instanceId = 0;
    LCGRand rng;
    rng.state = rng_state;
    glm::vec3 wi(0.0f);
    ShadingQueryAux aux;
    aux.sampling_pdf = 0.0f; aux.mis_pdf = 0.0f;
    const int result = shade_base_material(st, il, thr, 0, bm, lookup, area, glm::vec3(wo[0], wo[1], wo[2]), hit, rng, wi, aux);
    std::memset(out, 0, 19 * sizeof(float));
    out[0] = (float)result; out[1] = (float)st.bounce; out[2] = st.prev_bounce_pdf;
    out[3] = il.x; out[4] = il.y; out[5] = il.z; out[6] = thr.x; out[7] = thr.y; out[8] = thr.z;
    out[9] = wi.x; out[10] = wi.y; out[11] = wi.z; out[12] = aux.mis_pdf;
    std::memcpy(&out[13], &rng.state, 4);
    out[14] = (float)g_queries;
    out[15] = g_query_dir.x; out[16] = g_query_dir.y; out[17] = g_query_dir.z; out[18] = g_query_dist;
}

} // extern "C"
