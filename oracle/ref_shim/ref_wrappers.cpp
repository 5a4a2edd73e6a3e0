// oracle/ref_shim/ref_wrappers.cpp -- TEST INFRASTRUCTURE.
// Thin extern "C" wrappers around the reference's OWN sources, #included from where they lie under REF
// (/root/reference): the dual-language GLSL/C++ shading headers (the set rendering/tests/compile.cpp compiles),
// rendering/rt/hit.glsl + librender/dequantize.glsl, librender/quantize.h, librender/lights.cpp and the Hosek-Wilkie
// sky fit.  Nothing here restates reference arithmetic except the ~20 lines of glue marked "glue".
// Output: oracle/_ref/libref.so (git-ignored, travels to the GPU box with the snapshot).
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/rptr_types.h"

// ---- shading headers, as in rendering/tests/compile.cpp:4-41 -----------------------------------------------------
namespace ref {
using namespace glm;
#include "rendering/language.hpp"
#include "rendering/pointsets/lcg_rng.glsl"
#include "rendering/util.glsl"
#include "rendering/bsdfs/base_material.h.glsl"

#define NO_MATERIAL_REGISTRATION
namespace notr {
#include "rendering/bsdfs/gltf_bsdf.glsl"
}
#undef GLTF_BSDF_GLSL
#undef GLTF_COMPONENT_COUNT
#define GLTF_SUPPORT_TRANSMISSION
#define GLTF_SUPPORT_TRANSMISSION_ROUGHNESS
namespace tr {
#include "rendering/bsdfs/gltf_bsdf.glsl"
}
#undef GLTF_SUPPORT_TRANSMISSION
#undef GLTF_SUPPORT_TRANSMISSION_ROUGHNESS

#include "rendering/lights/tri.glsl"

static const TriLightData *g_lights = nullptr;
static int g_num_lights = 0;
static int g_bin_size = 16;
#define SCENE_GET_LIGHT_SOURCE(light_id) decode_tri_light(g_lights[light_id])
#define SCENE_GET_LIGHT_SOURCE_COUNT() int(g_num_lights)
#define BINNED_LIGHTS_BIN_MAX_SIZE 16
#define BINNED_LIGHTS_BIN_SIZE int(g_bin_size)
#define SCENE_GET_BINNED_LIGHTS_BIN_COUNT() ((g_num_lights + (g_bin_size - 1)) / g_bin_size)
namespace binned {
#include "rendering/mc/lights_linear.glsl"
}

// hit attributes with the megakernel's feature switches (vulkan/gpu_params.glsl:7-9)
#define DEFAULT_GEOMETRY_BUFFER_TYPES
#define QUANTIZED_POSITIONS
#define QUANTIZED_NORMALS_AND_UVS
#define NEED_MESH_ID_FOR_VISUALIZATION 0
#include "rendering/rt/hit.glsl"

#include "rendering/color/color_matching.h"
#include "rendering/color/color_matching.glsl"
} // namespace ref

namespace refq {
#include "librender/quantize.h"
}

#include "rendering/lights/sky_model_arhosek/sky_model.h"

static inline glm::vec3 V(const float *p) { return glm::vec3(p[0], p[1], p[2]); }
static inline void S(float *o, glm::vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

// glue: constants-only unpack_material (rendering/rt/material_textures.glsl:95-135) + load_material
template <class M> static M make_mat(const rptr_base_material *p, bool transmission) {
    M m;
    std::memset(&m, 0, sizeof(m));
    m.base_color = V(p->base_color);
    m.specular = p->specular;
    m.roughness = p->roughness;
    m.metallic = p->metallic;
    m.ior = p->ior;
    if (p->emission_intensity != 0.0f) m.base_color = glm::vec3(0.0f);
    m.flags = p->flags;
    (void)transmission;
    return m;
}
static ref::tr::GLTFMaterial make_mat_tr(const rptr_base_material *p) {
    ref::tr::GLTFMaterial m = make_mat<ref::tr::GLTFMaterial>(p, true);
    m.specular_transmission = p->specular_transmission;
    m.transmission_color = glm::vec3(0.0f);
    if (m.specular_transmission > 0.0f) {
        if (!(m.ior > 1.0f)) m.specular_transmission = 0.0f;
        else {
            m.transmission_color = m.base_color;
            m.transmission_roughness = m.roughness;
            m.roughness = std::sqrt(p->clearcoat_gloss);
        }
    }
    return m;
}

extern "C" {

uint32_t ref_lcg_seed(uint32_t index, uint32_t frame, uint32_t linear) { return ref::get_lcg_rng(index, frame, linear).state; }
float ref_lcg_randomf(uint32_t *state) {
    ref::LCGRand r;
    r.state = *state;
    float f = ref::lcg_randomf(r);
    *state = r.state;
    return f;
}
void ref_ortho_basis(const float *n, float *vx, float *vy) {
    glm::vec3 a, b;
    ref::ortho_basis(a, b, V(n));
    S(vx, a);
    S(vy, b);
}
float ref_fast_positive_atan(float y) { return ref::fast_positive_atan(y); }

void ref_gltf_bsdf(const rptr_base_material *p, const float *n, const float *wo, const float *wi, int tr, float *out) {
    glm::vec3 r = tr ? ref::tr::gltf_bsdf(make_mat_tr(p), V(n), V(wo), V(wi), glm::vec3(0), glm::vec3(0))
                     : ref::notr::gltf_bsdf(make_mat<ref::notr::GLTFMaterial>(p, false), V(n), V(wo), V(wi), glm::vec3(0), glm::vec3(0));
    S(out, r);
}
float ref_gltf_wpdf(const rptr_base_material *p, const float *n, const float *wo, const float *wi, int tr) {
    return tr ? ref::tr::gltf_wpdf(make_mat_tr(p), V(n), V(wo), V(wi), glm::vec3(0), glm::vec3(0))
              : ref::notr::gltf_wpdf(make_mat<ref::notr::GLTFMaterial>(p, false), V(n), V(wo), V(wi), glm::vec3(0), glm::vec3(0));
}
void ref_gltf_sample(const rptr_base_material *p, const float *n, const float *wo, const float *vx, const float *vy,
                     const float *rng_sample, const float *fresnel_sample, int tr, float *out) {
    glm::vec3 wi(0.0f);
    float pdf = 0.0f, mis = 0.0f;
    glm::vec2 rs(rng_sample[0], rng_sample[1]), fs(fresnel_sample[0], fresnel_sample[1]);
    glm::vec3 w = tr ? ref::tr::sample_gltf_brdf(make_mat_tr(p), V(n), V(wo), wi, pdf, mis, rs, fs, V(vx), V(vy))
                     : ref::notr::sample_gltf_brdf(make_mat<ref::notr::GLTFMaterial>(p, false), V(n), V(wo), wi, pdf, mis, rs, fs, V(vx), V(vy));
    S(out, w);
    S(out + 3, wi);
    out[6] = pdf;
    out[7] = mis;
}
void ref_triangle_solid_angle(const float *v0, const float *v1, const float *v2, float *out) {
    glm::vec3 prm;
    out[0] = ref::triangle_solid_angle(V(v0), V(v1), V(v2), prm);
    S(out + 1, prm);
}
void ref_sample_solid_angle_polygon(const float *v0, const float *v1, const float *v2, const float *rnd, float *out) {
    glm::vec3 prm;
    float omega = ref::triangle_solid_angle(V(v0), V(v1), V(v2), prm);
    S(out, ref::sample_solid_angle_polygon(V(v0), V(v1), V(v2), omega, prm, glm::vec2(rnd[0], rnd[1])));
}
void ref_sample_tri_lights(const rptr_tri_light_data *lights, int32_t n_lights, int32_t bin_size, const float *hit_p,
                           const float *hit_n, const float *dir_sample, const float *sel_sample, float *out) {
    static_assert(sizeof(rptr_tri_light_data) == sizeof(ref::TriLightData), "TriLightData layout");
    ref::g_lights = reinterpret_cast<const ref::TriLightData *>(lights);
    ref::g_num_lights = n_lights;
    ref::g_bin_size = bin_size;
    glm::vec3 ld(0.0f);
    float dist = 0.0f, pdf = 0.0f, mis = 0.0f;
    glm::vec3 L = ref::binned::sample_tri_lights(V(hit_p), V(hit_n), glm::vec2(dir_sample[0], dir_sample[1]),
                                                 glm::vec2(sel_sample[0], sel_sample[1]), ld, dist, pdf, mis);
    S(out, L);
    S(out + 3, ld);
    out[6] = dist;
    out[7] = pdf;
    out[8] = mis;
}
void ref_dequantize_position(uint64_t q, const float *scale, const float *offset, float *out) {
    using namespace glm;
    vec3 scaling = V(scale), off = V(offset);
    S(out, DEQUANTIZE_POSITION(q, scaling, off));
}
void ref_dequantize_normal(uint32_t w, float *out) { S(out, ref::dequantize_normal(w)); }
void ref_dequantize_uv(uint32_t w, float *out) {
    glm::vec2 uv = ref::dequantize_uv(w);
    out[0] = uv.x;
    out[1] = uv.y;
}
uint64_t ref_quantize_position(const float *p, const float *extent, const float *base) { return refq::quantize_position(V(p), V(extent), V(base)); }
uint32_t ref_quantize_normal(const float *n) { return refq::quantize_normal(V(n)); }
uint32_t ref_quantize_uv(const float *uv) { return refq::quantize_uv(glm::vec2(uv[0], uv[1]), glm::vec3(0.0f)); }
void ref_dequantization_params(const float *base, const float *extent, float *scale, float *offset) {
    S(scale, refq::dequantization_scaling(V(extent)));
    S(offset, refq::dequantization_offset(V(base), V(extent)));
}

// calc_hit_attributes (rendering/rt/hit.glsl:162-203 -> :58-128) on one unrolled triangle.
// w2o = 9 floats, row-major rows of world_to_object's 3x3; material_id/tri_mat as in RenderMeshParams.
void ref_hit_attributes(const uint64_t *qverts3, const uint64_t *qnuv3, const float *scale, const float *offset, int has_normals,
                        int has_uvs, const float *w2o, int material_id, const uint32_t *id_4pack, uint32_t prim, float t, float u,
                        float v, float *out) {
    using namespace ref;
    QuantizedVertexBuffer vb{const_cast<uint64_t *>(qverts3)};
    QuantizedNormalUVBuffer nb{const_cast<uint64_t *>(qnuv3)};
    MaterialIDBuffer mb{const_cast<uint32_t *>(id_4pack)};
    glm::uvec3 idx(0u, 1u, 2u);
    glm::mat3 verts = calc_hit_vertices(vb, V(scale), V(offset), idx);
    // transpose(mat3(world_to_object)): columns are the rows of world_to_object
    glm::mat3 n2w(glm::vec3(w2o[0], w2o[1], w2o[2]), glm::vec3(w2o[3], w2o[4], w2o[5]), glm::vec3(w2o[6], w2o[7], w2o[8]));
    RTHit h = calc_hit_attributes(t, prim, glm::vec2(u, v), verts, idx, n2w, nb, has_normals != 0, has_uvs != 0, material_id, mb);
    S(out, h.normal);
    out[3] = h.dist;
    S(out + 4, h.geo_normal);
    out[7] = float(h.material_id);
    S(out + 8, h.tangent);
    out[11] = h.bitangent_l;
    out[12] = h.uv.x;
    out[13] = h.uv.y;
}

// glue: update_sky_light (vulkan/render_sky.cpp:25-72) over the reference's sky_model.cpp and colour tables.
// sun_radiance.w is returned BEFORE the light-count rule (:67-70): 1 if the sun is up, else 0.
void ref_sky_fit(const rptr_scene_config *cfg, rptr_scene_params *out) {
    std::memset(out, 0, sizeof(*out));
    glm::vec3 sun_dir = glm::normalize(V(cfg->sun_dir));
    ArHosekSkyModelState state;
    arhosek_rgb_skymodelstate_alloc_init(cfg->turbidity, glm::dot(V(cfg->albedo), glm::vec3(0.3333f)), sun_dir.y, &state);
    S(out->sun_dir, sun_dir);
    out->sun_cos_angle = std::cos(glm::radians(0.53f) / 2.0f);
    for (int i = 0; i < 9; ++i) {
        out->sky_configs[i][0] = float(state.configs[0][i]);
        out->sky_configs[i][1] = float(state.configs[1][i]);
        out->sky_configs[i][2] = float(state.configs[2][i]);
        out->sky_configs[i][3] = 0.0f;
    }
    out->sky_radiances[0] = float(state.radiances[0]);
    out->sky_radiances[1] = float(state.radiances[1]);
    out->sky_radiances[2] = float(state.radiances[2]);
    ArHosekSkyModelState sunState;
    arhosekskymodelstate_alloc_init(state.elevation, state.turbidity, state.albedo, &sunState);
    glm::vec3 xyz(0.0f);
    int numSamples = 0;
    float last_wavelength = CM_CIE_MIN;
    for (int i = 0; i < CM_CIE_SAMPLES; ++i) {
        float wavelength = float(i) * float(CM_CIE_MAX - CM_CIE_MIN) / float(CM_CIE_SAMPLES - 1) + float(CM_CIE_MIN);
        if (wavelength > 720.0f) break;
        float radiance = arhosekskymodel_solar_radiance(&sunState, sun_dir.y, 0.0, wavelength);
        radiance -= arhosekskymodel_radiance(&sunState, sun_dir.y, 0.0, wavelength);
        {
            using namespace ref;
            xyz += glm::vec3(CM_TABLE_X[i], CM_TABLE_Y[i], CM_TABLE_Z[i]) * radiance;
        }
        ++numSamples;
        last_wavelength = wavelength;
    }
    xyz *= float(last_wavelength - CM_CIE_MIN) / float(numSamples);
    if (sun_dir.y > 0.0f && glm::all(glm::greaterThanEqual(xyz, glm::vec3(0.0f)))) {
        glm::vec3 rgb = 0.01f * ref::xyz_to_srgb(xyz);
        S(out->sun_radiance, rgb);
        out->sun_radiance[3] = 1.0f;
    }
    out->normal_z_scale = 1.0f / cfg->bump_scale; // vulkan/render_vulkan.cpp:2954-2959
}

} // extern "C"
