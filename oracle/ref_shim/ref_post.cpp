// oracle/ref_shim/ref_post.cpp -- TEST INFRASTRUCTURE.
// The temporal passes of the reference's ENABLE_REALTIME_RESOLVE build executed as C++ from the shader sources where they lie:
//   * rendering/postprocess/reprojection.glsl: the option block (:17-24), the three test-tonemap helpers and the whole of
//     reproject_and_accumulate (:44-367), cut out by the Makefile into oracle/_ref/gen/reproject.inc.  The only edit is
//     mechanical, done by sed at build time: the GLSL swizzle `v.xyz` becomes `vec3(v)` and the one swizzled store
//     `accum_color.xyz = E;` becomes `accum_color = vec4(E, accum_color.w);` (C++ has no swizzles), and the implicit ivec2 -> vec2
//     conversion of `vec2 anchor_point = ivec2(..)` is written out;
//   * vulkan/processing/process_taa.comp: lanczosWeight, lanczos and main() (:28-112) into gen/process_taa.inc, unedited.
// This file supplies what the shaders get from their environment: the images behind the REPROJECTION_* macros / bindings,
// imageLoad / imageStore / texelFetch / textureLod, a few GLSL built-ins the shared glm shim does not have, and the push constants.
// Driver-defined behaviour is given the same reading as in oracle/post_oracle.h (out-of-range loads read zero, LINEAR +
// CLAMP_TO_EDGE sampling in exact fp32, rgba8 stores round to nearest); exp / sin / sqrt are libm's here, so the oracle is
// pinned to these outputs within a tolerance, not bit for bit.
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace refp {
using namespace glm;
typedef unsigned int uint;
#include "rendering/language.hpp"
#ifndef M_PI_F_SHIM
#undef M_PI
#define M_PI 3.14159265358979323846f // rendering/defaults.glsl:8-10
#endif

static float half_to_float(uint16_t h) {
    const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t u;
    if (e == 0) { float f = std::ldexp((float)m, -24); return s ? -f : f; }
    if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &u, 4); return f;
}
struct Image {
    int w = 0, h = 0, kind = 0; // 0: RGBA32F, 1: RGBA16F, 2: RGBA8
    const void *p = nullptr;
    void *out = nullptr;        // imageStore target (same format)
    vec4 at(ivec2 c) const {
        if (c.x < 0 || c.y < 0 || c.x >= w || c.y >= h) return vec4(0.0f);
        const size_t i = 4 * ((size_t)c.y * w + c.x);
        if (kind == 0) { const float *f = (const float *)p; return vec4(f[i], f[i + 1], f[i + 2], f[i + 3]); }
        if (kind == 1) { const uint16_t *q = (const uint16_t *)p; return vec4(half_to_float(q[i]), half_to_float(q[i + 1]), half_to_float(q[i + 2]), half_to_float(q[i + 3])); }
        const uint8_t *b = (const uint8_t *)p;
        return vec4((float)b[i] / 255.0f, (float)b[i + 1] / 255.0f, (float)b[i + 2] / 255.0f, (float)b[i + 3] / 255.0f);
    }
};
static vec4 imageLoad(const Image &im, ivec2 c) { return im.at(c); }
static vec4 texelFetch(const Image &im, ivec2 c, int) { return im.at(c); }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static vec4 textureLod(const Image &im, vec2 uv, float) { // screen_sampler: LINEAR, CLAMP_TO_EDGE (vulkan/render_vulkan.cpp:417-427)
    const float x = uv.x * (float)im.w - 0.5f, y = uv.y * (float)im.h - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y), a = x - fx, b = y - fy;
    const int i0 = clampi((int)fx, 0, im.w - 1), i1 = clampi((int)fx + 1, 0, im.w - 1), j0 = clampi((int)fy, 0, im.h - 1), j1 = clampi((int)fy + 1, 0, im.h - 1);
    const vec4 t00 = im.at(ivec2(i0, j0)), t10 = im.at(ivec2(i1, j0)), t01 = im.at(ivec2(i0, j1)), t11 = im.at(ivec2(i1, j1));
    const vec4 top = t00 * (1.0f - a) + t10 * a, bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
}
static void imageStore(Image &im, ivec2 c, vec4 v) {
    if (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h || !im.out) return;
    const size_t i = 4 * ((size_t)c.y * im.w + c.x);
    if (im.kind == 0) { float *f = (float *)im.out; f[i] = v.x; f[i + 1] = v.y; f[i + 2] = v.z; f[i + 3] = v.w; }
    else if (im.kind == 2) {
        uint8_t *b = (uint8_t *)im.out;
        for (int k = 0; k < 4; ++k) b[i + k] = (uint8_t)(std::fmin(std::fmax(v[k], 0.0f), 1.0f) * 255.0f + 0.5f);
    }
}
// GLSL built-ins and implicit conversions the shaders use beyond the shared shim
using glm::clamp; using glm::max; using glm::min; using glm::abs; using glm::ceil; using glm::floor; using glm::all; using glm::sqrt; using glm::exp;
static float log2(float x) { return std::log2(x); }
static float exp2(float x) { return std::exp2(x); }
static vec4 log2(vec4 v) { return vec4(log2(v.x), log2(v.y), log2(v.z), log2(v.w)); }
static vec4 exp2(vec4 v) { return vec4(exp2(v.x), exp2(v.y), exp2(v.z), exp2(v.w)); }
static float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
static float smoothstep(float e0, float e1, float x) { const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
static float length2(vec2 v) { return dot(v, v); } // rendering/util.glsl:103-105
static vec2 clamp(vec2 v, vec2 lo, vec2 hi) { return vec2(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y)); }
static vec4 clamp(vec4 v, vec4 lo, vec4 hi) { return vec4(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z), clamp(v.w, lo.w, hi.w)); }
static vec2 ceil(vec2 v) { return vec2(std::ceil(v.x), std::ceil(v.y)); }
static vec4 max(vec4 a, float b) { return vec4(max(a.x, b), max(a.y, b), max(a.z, b), max(a.w, b)); }
static vec4 max(vec4 a, vec4 b) { return vec4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
static vec4 min(vec4 a, vec4 b) { return vec4(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)); }
static ivec2 abs(ivec2 v) { return ivec2(v.x < 0 ? -v.x : v.x, v.y < 0 ? -v.y : v.y); }
static bvec2 lessThanEqual(ivec2 a, ivec2 b) { return bvec2(a.x <= b.x, a.y <= b.y); }
static bool all(bvec2 v) { return v.x && v.y; }
static bool operator==(ivec2 a, ivec2 b) { return a.x == b.x && a.y == b.y; }
static bool operator==(ivec2 a, vec2 b) { return (float)a.x == b.x && (float)a.y == b.y; } // GLSL converts the ivec2 operand
static vec2 operator*(vec2 a, ivec2 b) { return vec2(a.x * (float)b.x, a.y * (float)b.y); }   // likewise
static vec2 operator*(int a, vec2 b) { return vec2((float)a * b.x, (float)a * b.y); }
static vec2 operator/(vec2 a, int b) { return vec2(a.x / (float)b, a.y / (float)b); }
static ivec2 operator*(ivec2 a, int b) { return ivec2(a.x * b, a.y * b); }
static ivec2 operator*(int a, ivec2 b) { return ivec2(a * b.x, a * b.y); }
static ivec2 operator/(ivec2 a, int b) { return ivec2(a.x / b, a.y / b); }

// ---- reprojection.glsl ---------------------------------------------------------------------------------------------------
static Image g_motion, g_history, g_nd_history, g_accum, g_nd;
#define REPROJECTION_MOTION_JITTER_BUFFER g_motion
#define REPROJECTION_ACCUM_HISTORY g_history
#define REPROJECTION_NORMAL_DEPTH_HISTORY g_nd_history
#define REPROJECTION_ACCUM_TARGET g_accum
#define REPROJECTION_ACCUM_NORMAL_DEPTH_TARGET g_nd
#define accum_buffer g_accum
#include "gen/reproject.inc"
#undef accum_buffer

// ---- process_taa.comp ----------------------------------------------------------------------------------------------------
static Image framebuffer, history_framebuffer, aov_motion_jitter_buffer;
static ivec2 fb_dims;
static int render_upscale_factor;
static struct { uvec2 xy; } gl_GlobalInvocationID;
#define ENABLE_AOV_BUFFERS
#include "gen/process_taa.inc"
} // namespace refp

extern "C" {
void ref_reproject_accumulate(int32_t w, int32_t h, const float *accum_cur, const float *history, const uint16_t *nd_history, const uint16_t *nd,
                              const uint16_t *mj, float min_sample_weight, int32_t batch, float *stored, float *shown) {
    using namespace refp;
    std::memcpy(stored, accum_cur, sizeof(float) * 4 * (size_t)w * h); // the pass reads and writes the accumulator in place
    g_motion = Image{w, h, 1, mj, nullptr};
    g_history = Image{w, h, 0, history, nullptr};
    g_nd_history = Image{w, h, 1, nd_history, nullptr};
    g_nd = Image{w, h, 1, nd, nullptr};
    // reads of the accumulator see this frame's samples (the shader's neighbourhood reads race with its own stores; their results
    // are unused in the shipped configuration)
    g_accum = Image{w, h, 0, accum_cur, stored};
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t i = 4 * ((size_t)y * w + x);
            const glm::vec4 r = reproject_and_accumulate(glm::vec4(accum_cur[i], accum_cur[i + 1], accum_cur[i + 2], accum_cur[i + 3]), glm::ivec2(x, y),
                                                         glm::ivec2(w, h), min_sample_weight, 1 /* sample_base_index > 0 */, batch, 0.0f, 0.0f);
            shown[i] = r.x; shown[i + 1] = r.y; shown[i + 2] = r.z; shown[i + 3] = r.w;
        }
}
void ref_process_taa(int32_t w, int32_t h, int32_t upscale, int32_t rw, int32_t rh, const uint8_t *current, const uint8_t *history, const uint16_t *mj,
                     uint8_t *out) {
    using namespace refp;
    std::memcpy(out, current, 4 * (size_t)w * h);
    framebuffer = Image{w, h, 2, current, out}; // every read sees the target as process_samples left it
    history_framebuffer = Image{w, h, 2, history, nullptr};
    aov_motion_jitter_buffer = Image{rw, rh, 1, mj, nullptr};
    fb_dims = glm::ivec2(w, h);
    render_upscale_factor = upscale;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            gl_GlobalInvocationID.xy = glm::uvec2((unsigned)x, (unsigned)y);
            refp::main();
        }
}
} // extern "C"
