// oracle/ref_shim/ref_materials.cpp -- TEST INFRASTRUCTURE.
// The reference's OWN material decode, #included from where it lies under REF and executed as C++:
// rendering/rt/materials.glsl + rendering/bsdfs/gltf_bsdf.glsl (load_material) + rendering/rt/material_textures.glsl
// (textured_color_param / textured_scalar_param / unpack_material / get_material_alpha), in the non-unrolled standard-texture
// configuration, with and without GLTF_SUPPORT_TRANSMISSION[_ROUGHNESS] and with PREMULTIPLIED_BASE_COLOR_ALPHA
// (vulkan/gpu_params.glsl:12).  The only thing supplied here is the texture unit: SCENE_GET_TEXTURE(id) + textureLod() return
// the single texel of a 1 x 1 texture handed in by the caller as four floats (what a sampler returns for any uv / LOD).
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "../../include/rptr_types.h"

namespace refmat {
using namespace glm;
typedef unsigned int uint;
struct Texel1x1 { vec4 v; };
static const Texel1x1 *g_textures = nullptr;
inline vec4 textureLod(const Texel1x1 &t, vec2, float) { return t.v; }
inline uint32_t floatBitsToUint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
#define SCENE_GET_TEXTURE(index) g_textures[index]
#define PREMULTIPLIED_BASE_COLOR_ALPHA
#define NO_MATERIAL_REGISTRATION

#include "rendering/language.hpp"
#include "rendering/util.glsl"
#include "rendering/bsdfs/base_material.h.glsl"

#include "rendering/defaults.glsl"
namespace notr {
#include "rendering/rt/materials.glsl"
#include "rendering/bsdfs/gltf_bsdf.glsl"
#define MATERIAL_TYPE GLTFMaterial
#include "rendering/rt/material_textures.glsl"
}
#undef MATERIAL_DECODE_HEADER
#undef MATERIAL_TEXTURE_DECODE_HEADER
#undef HIT_POINT_H_GLSL
#undef GLTF_BSDF_GLSL
#undef GLTF_COMPONENT_COUNT
#undef MATERIAL_TYPE
#undef ENABLE_MATERIAL_DECODE
#undef EmitterParams
#undef textured_scalar_standard_param
#undef textured_color_standard_param
#undef get_standard_texture_sampler
#undef get_base_material_alpha
#undef NO_TEXTURE_GRAD
#define GLTF_SUPPORT_TRANSMISSION
#define GLTF_SUPPORT_TRANSMISSION_ROUGHNESS
namespace tr {
#include "rendering/rt/materials.glsl"
#include "rendering/bsdfs/gltf_bsdf.glsl"
#define MATERIAL_TYPE GLTFMaterial
#include "rendering/rt/material_textures.glsl"
}
// third configuration: USE_MIPMAPPING as librender/render_params.glsl.h:8 defines it for the megakernel, i.e. the texture reads are
// textureGrad(SCENE_GET_TEXTURE(id), hit.uv, hit.duvdxy[0], hit.duvdxy[1]) (material_textures.glsl:37-63).  The sampler object only
// carries its index; textureGrad hands (index, uv, dPdx, dPdy) to a caller-supplied texture unit.
#undef MATERIAL_DECODE_HEADER
#undef MATERIAL_TEXTURE_DECODE_HEADER
#undef HIT_POINT_H_GLSL
#undef GLTF_BSDF_GLSL
#undef GLTF_COMPONENT_COUNT
#undef MATERIAL_TYPE
#undef ENABLE_MATERIAL_DECODE
#undef EmitterParams
#undef textured_scalar_standard_param
#undef textured_color_standard_param
#undef get_standard_texture_sampler
#undef get_base_material_alpha
#undef NO_TEXTURE_GRAD
#undef SCENE_GET_TEXTURE
#define USE_MIPMAPPING
struct SamplerIndex { uint32_t id; };
typedef void (*texture_unit_fn)(uint32_t id, const float *uv, const float *dpdx, const float *dpdy, float *rgba);
static texture_unit_fn g_texture_unit = nullptr;
inline vec4 textureGrad(SamplerIndex s, vec2 uv, vec2 dpdx, vec2 dpdy) {
    float rgba[4] = {0, 0, 0, 0};
    const float a[2] = {uv.x, uv.y}, b[2] = {dpdx.x, dpdx.y}, c[2] = {dpdy.x, dpdy.y};
    g_texture_unit(s.id, a, b, c, rgba);
    return vec4(rgba[0], rgba[1], rgba[2], rgba[3]);
}
inline vec4 textureLod(SamplerIndex, vec2, float) { return vec4(0.0f); } // not reached with USE_MIPMAPPING
#define SCENE_GET_TEXTURE(index) SamplerIndex{(uint32_t)(index)}
namespace grad {
#include "rendering/rt/materials.glsl"
#include "rendering/bsdfs/gltf_bsdf.glsl"
#define MATERIAL_TYPE GLTFMaterial
#include "rendering/rt/material_textures.glsl"
}
} // namespace refmat

extern "C" {
// unpack_material + get_material_alpha in the USE_MIPMAPPING + transmission configuration with image textures: every texture read goes
// to `unit` with the hit's uv and duvdxy (column-major 2 x 2: d(uv)/dx, d(uv)/dy).  Output layout as ref_unpack_material.
void ref_unpack_material_grad(const rptr_base_material *p, const float *uv, const float *duvdxy, refmat::texture_unit_fn unit, float *out) {
    using namespace refmat;
    g_texture_unit = unit;
    BaseMaterial bm;
    std::memcpy(&bm, p, sizeof(bm));
    std::memset(out, 0, 17 * sizeof(float));
    grad::GLTFMaterial m;
    std::memset(&m, 0, sizeof(m));
    grad::EmitterInteraction e;
    grad::HitPoint hit{glm::vec3(0.0f), glm::vec2(uv[0], uv[1]), glm::mat2(duvdxy[0], duvdxy[1], duvdxy[2], duvdxy[3]), glm::vec3(0.0f, 0.0f, 1.0f)};
    out[15] = grad::unpack_material(m, e, 0u, bm, hit);
    out[16] = grad::get_material_alpha(0u, bm, hit);
    out[0] = m.base_color.x; out[1] = m.base_color.y; out[2] = m.base_color.z;
    out[3] = m.metallic; out[4] = m.specular; out[5] = m.roughness; out[6] = m.ior;
    out[7] = m.specular_transmission; out[8] = m.transmission_roughness;
    out[9] = m.transmission_color.x; out[10] = m.transmission_color.y; out[11] = m.transmission_color.z;
    out[12] = e.radiance.x; out[13] = e.radiance.y; out[14] = e.radiance.z;
}


// unpack_material(mat, emitter, material_id, params, hit) + get_material_alpha(...) for one BaseMaterial whose parameters may
// carry texture handles into `texels` (n_textures x 4 floats, the value textureLod returns).
// out[0..2] base_color, [3] metallic, [4] specular, [5] roughness, [6] ior, [7] specular_transmission,
// [8] transmission_roughness, [9..11] transmission_color, [12..14] emitter radiance, [15] alpha returned by unpack_material,
// [16] get_material_alpha
void ref_unpack_material(const rptr_base_material *p, const float *texels, int n_textures, int transmission, float *out) {
    using namespace refmat;
    static Texel1x1 store[64];
    for (int i = 0; i < n_textures && i < 64; ++i) store[i].v = glm::vec4(texels[4 * i], texels[4 * i + 1], texels[4 * i + 2], texels[4 * i + 3]);
    g_textures = store;
    BaseMaterial bm;
    static_assert(sizeof(BaseMaterial) == sizeof(rptr_base_material), "BaseMaterial layout");
    std::memcpy(&bm, p, sizeof(bm));
    std::memset(out, 0, 17 * sizeof(float));
    if (transmission) {
        tr::GLTFMaterial m;
        std::memset(&m, 0, sizeof(m));
        tr::EmitterInteraction e;
        tr::HitPoint hit{glm::vec3(0.0f), glm::vec2(0.25f, 0.75f), glm::mat2(0.0f), glm::vec3(0.0f, 0.0f, 1.0f)};
        out[15] = tr::unpack_material(m, e, 0u, bm, hit);
        out[16] = tr::get_material_alpha(0u, bm, hit);
        out[0] = m.base_color.x; out[1] = m.base_color.y; out[2] = m.base_color.z;
        out[3] = m.metallic; out[4] = m.specular; out[5] = m.roughness; out[6] = m.ior;
        out[7] = m.specular_transmission; out[8] = m.transmission_roughness;
        out[9] = m.transmission_color.x; out[10] = m.transmission_color.y; out[11] = m.transmission_color.z;
        out[12] = e.radiance.x; out[13] = e.radiance.y; out[14] = e.radiance.z;
    } else {
        notr::GLTFMaterial m;
        std::memset(&m, 0, sizeof(m));
        notr::EmitterInteraction e;
        notr::HitPoint hit{glm::vec3(0.0f), glm::vec2(0.25f, 0.75f), glm::mat2(0.0f), glm::vec3(0.0f, 0.0f, 1.0f)};
        out[15] = notr::unpack_material(m, e, 0u, bm, hit);
        out[16] = notr::get_material_alpha(0u, bm, hit);
        out[0] = m.base_color.x; out[1] = m.base_color.y; out[2] = m.base_color.z;
        out[3] = m.metallic; out[4] = m.specular; out[5] = m.roughness; out[6] = m.ior;
        out[12] = e.radiance.x; out[13] = e.radiance.y; out[14] = e.radiance.z;
    }
}

} // extern "C"
