// oracle/ref_shim/ref_nee.cpp -- TEST INFRASTRUCTURE.
// The reference's OWN next-event estimation, #included from where it lies under REF and executed as C++:
// rendering/mc/nee.glsl (sample_direct_light) with rendering/mc/{lights_sun,lights_linear,nee_interface}.glsl,
// rendering/lights/{sun,tri}.glsl and the glTF BSDF registered as MATERIAL_TYPE, exactly as vulkan/pt_megakernel.glsl:22-109
// assembles them.  Supplied here: the scene_params uniform block (sun + light count), the binned light buffer, and
// raytrace_test_visibility(), which records the shadow-ray query and reports "visible" (visibility is the trace stage's job).
// Also executed here: compute_sky_illum(), the miss shading of vulkan/pt_megakernel.glsl (the one self-contained function of
// that file; the rest needs rayQueryEXT).  The Makefile cuts exactly that function out of the file where it lies into the
// git-ignored build directory oracle/_ref/gen/ for the duration of the compile and deletes it afterwards.
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "../../include/rptr_types.h"

namespace refnee {
using namespace glm;
typedef unsigned int uint;
#include "rendering/language.hpp"
#include "rendering/defaults.glsl"
#include "rendering/util.glsl"
#include "rendering/bsdfs/base_material.h.glsl"
#include "rendering/bsdfs/hit_point.glsl"
#include "rendering/lights/tri.glsl"
#include "rendering/lights/sky_model_arhosek/sky_model.glsl"

struct SceneParamsStandIn { // the members of SceneParams (vulkan/gpu_params.glsl:120-131) that nee.glsl / compute_sky_illum read
    vec3 sun_dir;
    float sun_cos_angle;
    vec4 sun_radiance;
    SkyModelParams sky_params;
};
static SceneParamsStandIn scene_params;
static const TriLightData *g_lights = nullptr;
static int g_num_lights = 0;
static int g_bin_size = 16;
#define SCENE_GET_LIGHT_SOURCE(light_id) decode_tri_light(g_lights[light_id])
#define SCENE_GET_LIGHT_SOURCE_COUNT() int(g_num_lights)
#define BINNED_LIGHTS_BIN_MAX_SIZE 16
#define BINNED_LIGHTS_BIN_SIZE int(g_bin_size)
#define SCENE_GET_BINNED_LIGHTS_BIN_COUNT() ((g_num_lights + (g_bin_size - 1)) / g_bin_size)

static vec3 g_query_from, g_query_dir;
static float g_query_dist;
static int g_queries;

namespace notr {
#include "rendering/bsdfs/gltf_bsdf.glsl"
inline bool raytrace_test_visibility(const vec3 from, const vec3 dir, float dist);
#include "rendering/mc/nee.glsl"
inline bool raytrace_test_visibility(const vec3 from, const vec3 dir, float dist) {
    g_query_from = from; g_query_dir = dir; g_query_dist = dist; ++g_queries;
    return true;
}
#include "gen/compute_sky_illum.inc"
}
} // namespace refnee

extern "C" {

// sample_direct_light(mat, hit, w_o, dir_sample, sel_sample, aux) for a constants-only material.
// in: material, hit point p / geometric normal gn / shading normal n / tangent frame v_x, v_y, w_o, 4 uniforms (dir.xy, sel.xy),
//     sun block (sun_dir[3], sun_cos_angle, sun_radiance[4] with w = p_sun), binned lights + bin size.
// out[0..2] contribution (throughput excluded), [3..5] light_dir, [6] light_dist, [7] mis_pdf, [8] number of shadow queries,
// [9..11] query origin, [12..14] query dir, [15] query dist
void ref_sample_direct_light(const rptr_base_material *p, const float *hp, const float *gn, const float *n, const float *vx, const float *vy,
                             const float *wo, const float *u4, const float *sun_dir, float sun_cos_angle, const float *sun_radiance,
                             const rptr_tri_light_data *lights, int n_lights, int bin_size, float *out) {
    using namespace refnee;
    scene_params.sun_dir = glm::vec3(sun_dir[0], sun_dir[1], sun_dir[2]);
    scene_params.sun_cos_angle = sun_cos_angle;
    scene_params.sun_radiance = glm::vec4(sun_radiance[0], sun_radiance[1], sun_radiance[2], sun_radiance[3]);
    g_lights = reinterpret_cast<const TriLightData *>(lights);
    g_num_lights = n_lights;
    g_bin_size = bin_size;
    g_queries = 0;
    notr::GLTFMaterial m;
    std::memset(&m, 0, sizeof(m));
    m.base_color = glm::vec3(p->base_color[0], p->base_color[1], p->base_color[2]);
    m.metallic = p->metallic; m.specular = p->specular; m.roughness = p->roughness; m.ior = p->ior; m.flags = p->flags;
    if (p->emission_intensity != 0.0f) m.base_color = glm::vec3(0.0f);
    InteractionPoint hit;
    hit.p = glm::vec3(hp[0], hp[1], hp[2]);
    hit.gn = glm::vec3(gn[0], gn[1], gn[2]);
    hit.n = glm::vec3(n[0], n[1], n[2]);
    hit.v_x = glm::vec3(vx[0], vx[1], vx[2]);
    hit.v_y = glm::vec3(vy[0], vy[1], vy[2]);
    hit.primitiveId = 0; hit.instanceId = 0;
    notr::NEEQueryAux aux;
    aux.light_dir = glm::vec3(0.0f); aux.light_dist = 0.0f; aux.mis_pdf = 0.0f;
    glm::vec3 r = notr::sample_direct_light(m, hit, glm::vec3(wo[0], wo[1], wo[2]), glm::vec2(u4[0], u4[1]), glm::vec2(u4[2], u4[3]), aux);
    std::memset(out, 0, 16 * sizeof(float));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
    out[3] = aux.light_dir.x; out[4] = aux.light_dir.y; out[5] = aux.light_dir.z;
    out[6] = aux.light_dist; out[7] = aux.mis_pdf; out[8] = (float)g_queries;
    out[9] = g_query_from.x; out[10] = g_query_from.y; out[11] = g_query_from.z;
    out[12] = g_query_dir.x; out[13] = g_query_dir.y; out[14] = g_query_dir.z; out[15] = g_query_dist;
}

// compute_sky_illum(ray_origin, ray_dir, prev_bsdf_pdf) (vulkan/pt_megakernel.glsl:113-149) with the fitted sky block of sp;
// sp->sun_radiance[3] = p_sun as the shader sees it (after the light-count rule of vulkan/render_sky.cpp:67-70).
void ref_compute_sky_illum(const rptr_scene_params *sp, const float *ray_origin, const float *ray_dir, float prev_bsdf_pdf, float *out) {
    using namespace refnee;
    scene_params.sun_dir = glm::vec3(sp->sun_dir[0], sp->sun_dir[1], sp->sun_dir[2]);
    scene_params.sun_cos_angle = sp->sun_cos_angle;
    scene_params.sun_radiance = glm::vec4(sp->sun_radiance[0], sp->sun_radiance[1], sp->sun_radiance[2], sp->sun_radiance[3]);
    for (int i = 0; i < 9; ++i)
        scene_params.sky_params.configs[i] = glm::vec4(sp->sky_configs[i][0], sp->sky_configs[i][1], sp->sky_configs[i][2], sp->sky_configs[i][3]);
    scene_params.sky_params.radiances = glm::vec4(sp->sky_radiances[0], sp->sky_radiances[1], sp->sky_radiances[2], sp->sky_radiances[3]);
    glm::vec3 r = notr::compute_sky_illum(glm::vec3(ray_origin[0], ray_origin[1], ray_origin[2]), glm::vec3(ray_dir[0], ray_dir[1], ray_dir[2]), prev_bsdf_pdf);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

} // extern "C"
