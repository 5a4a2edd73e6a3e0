// oracle/ref_shim/ref_sky.cpp -- TEST INFRASTRUCTURE.
// extern "C" wrappers around two more of the reference's own GLSL sources that compile as C++ with the shim, #included from
// where they lie under REF: rendering/lights/sky_model_arhosek/sky_model.glsl (skymodel_radiance, the Hosek-Wilkie RGB
// evaluation the miss shader calls) and rendering/lights/sun.glsl (sample_sun_dir / sample_sun_dir_pdf of the sun NEE).
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "../../include/rptr_types.h"

namespace refsky {
using namespace glm;
#include "rendering/language.hpp"
#include "rendering/util.glsl"
#include "rendering/lights/sky_model_arhosek/sky_model.glsl"
#include "rendering/lights/sun.glsl"
#include "rendering/postprocess/tonemapping_utils.glsl"
} // namespace refsky

extern "C" {

// skymodel_radiance(SkyModelParams{configs, radiances}, sun_dir, view_dir) with the fitted block of rptr_scene_params
void ref_skymodel_radiance(const rptr_scene_params *sp, const float *sun_dir, const float *view_dir, float *out) {
    refsky::SkyModelParams p;
    for (int i = 0; i < 9; ++i) p.configs[i] = glm::vec4(sp->sky_configs[i][0], sp->sky_configs[i][1], sp->sky_configs[i][2], sp->sky_configs[i][3]);
    p.radiances = glm::vec4(sp->sky_radiances[0], sp->sky_radiances[1], sp->sky_radiances[2], sp->sky_radiances[3]);
    glm::vec3 r = refsky::skymodel_radiance(p, glm::vec3(sun_dir[0], sun_dir[1], sun_dir[2]), glm::vec3(view_dir[0], view_dir[1], view_dir[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// sample_sun_dir(sun_dir, cos_radius, sample) -> out[0..2], sample_sun_dir_pdf -> out[3]
void ref_sample_sun_dir(const float *sun_dir, float cos_radius, const float *u2, float *out) {
    const glm::vec3 s(sun_dir[0], sun_dir[1], sun_dir[2]);
    glm::vec3 d = refsky::sample_sun_dir(s, cos_radius, glm::vec2(u2[0], u2[1]));
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
    out[3] = refsky::sample_sun_dir_pdf(s, cos_radius, d);
}

// tonemap(mode, rgb) of rendering/postprocess/tonemapping_utils.glsl:16-32 followed by linear_to_srgb (rendering/util.glsl:25-28):
// the display chain of process_samples.comp:148-149, 188.  out = tonemapped rgb (3), then its sRGB encoding (3)
void ref_tonemap_srgb(int32_t mode, const float *rgb, float *out) {
    glm::vec3 c = refsky::tonemap(mode, glm::vec3(rgb[0], rgb[1], rgb[2]));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
    out[3] = refsky::linear_to_srgb(c.x); out[4] = refsky::linear_to_srgb(c.y); out[5] = refsky::linear_to_srgb(c.z);
}

// srgb_to_linear (rendering/util.glsl:39-42): the reference's own statement of the sRGB transfer function, the cross-check for
// the texel decode the backend performs in place of the texture unit
float ref_srgb_to_linear(float x) { return refsky::srgb_to_linear(x); }

} // extern "C"
