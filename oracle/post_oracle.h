// oracle/post_oracle.h -- TEST INFRASTRUCTURE (CPU oracle).  Never included by the product.
//
// The temporal passes of the reference's ENABLE_REALTIME_RESOLVE build, restated statement by statement -- including the parts
// whose results nothing reads in the shipped configuration (reprojection.glsl:17-24), so that the product's pruned version
// (csrc/rptr_post.cuh) is checked against the WHOLE control flow:
//   reproject_and_accumulate   rendering/postprocess/reprojection.glsl:44-367
//   process_taa main()         vulkan/processing/process_taa.comp:28-112
// Conventions for what the driver decides (same list as in rptr_post.cuh): out-of-range image loads and texel fetches read
// zero; the history sampler (LINEAR, CLAMP_TO_EDGE: render_vulkan.cpp:417-427) blends in exact fp32; float -> int truncates
// with NaN -> 0; rgba8 stores round to nearest, loads give v / 255; exp / sin are the RPTR-FP kernels of fp32.h.
// The shader reads `accum_buffer` around the pixel while other invocations overwrite it (the mean / variance block, :228-243);
// nothing uses that block's result in this configuration, so it is the one piece left out here as well.
#pragma once
#include <cstddef>

#include "fp32.h"

namespace post {
using namespace fp;

struct IV2 { int x, y; };
static inline float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h >> 15) << 31;
    const int exponent = (h >> 10) & 0x1f;
    const uint32_t mant = h & 0x3ffu;
    if (exponent == 0) return (sign ? -1.0f : 1.0f) * ldexpf((float)mant, -24);
    if (exponent == 31) return u2f(sign | 0x7f800000u | (mant << 13));
    return u2f(sign | (uint32_t)(exponent - 15 + 127) << 23 | (mant << 13));
}
static inline int to_int(float x) {
    if (x != x) return 0;
    if (x >= 2147483520.0f) return 2147483520;
    if (x <= -2147483520.0f) return -2147483520;
    return (int)x;
}
struct Image4f { const float *p; int w, h; };
struct Image4h { const uint16_t *p; int w, h; };
struct Image4b { const uint8_t *p; int w, h; };
static inline V4 fetch(Image4f im, IV2 c) {
    if (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) return V4{0, 0, 0, 0};
    const float *t = im.p + 4 * ((size_t)c.y * im.w + c.x);
    return V4{t[0], t[1], t[2], t[3]};
}
static inline V4 fetch(Image4h im, IV2 c) {
    if (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) return V4{0, 0, 0, 0};
    const uint16_t *t = im.p + 4 * ((size_t)c.y * im.w + c.x);
    return V4{half_to_float(t[0]), half_to_float(t[1]), half_to_float(t[2]), half_to_float(t[3])};
}
static inline V4 fetch(Image4b im, IV2 c) {
    if (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) return V4{0, 0, 0, 0};
    const uint8_t *t = im.p + 4 * ((size_t)c.y * im.w + c.x);
    return V4{(float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f, (float)t[3] / 255.0f};
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline V4 texture_linear(Image4f im, V2 uv) { // textureLod(sampler2D, uv, 0), LINEAR + CLAMP_TO_EDGE
    const float x = uv.x * (float)im.w - 0.5f, y = uv.y * (float)im.h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    const int i0 = clampi(to_int(fx), 0, im.w - 1), i1 = clampi(to_int(fx) + 1, 0, im.w - 1);
    const int j0 = clampi(to_int(fy), 0, im.h - 1), j1 = clampi(to_int(fy) + 1, 0, im.h - 1);
    const V4 t00 = fetch(im, IV2{i0, j0}), t10 = fetch(im, IV2{i1, j0}), t01 = fetch(im, IV2{i0, j1}), t11 = fetch(im, IV2{i1, j1});
    V4 r;
    r.x = mix(mix(t00.x, t10.x, a), mix(t01.x, t11.x, a), b);
    r.y = mix(mix(t00.y, t10.y, a), mix(t01.y, t11.y, a), b);
    r.z = mix(mix(t00.z, t10.z, a), mix(t01.z, t11.z, a), b);
    r.w = mix(mix(t00.w, t10.w, a), mix(t01.w, t11.w, a), b);
    return r;
}
static inline float smoothstep(float e0, float e1, float x) {
    const float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
static inline int iabs(int v) { return v < 0 ? -v : v; }

struct ReprojectIn {
    Image4f history;      // REPROJECTION_ACCUM_HISTORY
    Image4h nd_history;   // REPROJECTION_NORMAL_DEPTH_HISTORY
    Image4h nd;           // REPROJECTION_ACCUM_NORMAL_DEPTH_TARGET (read only without REPROJECTION_ACCUM_GBUFFER)
    Image4h mj;           // REPROJECTION_MOTION_JITTER_BUFFER
};

// returns the function's return value; *store = the imageStore to REPROJECTION_ACCUM_TARGET
static inline V4 reproject_and_accumulate(const ReprojectIn &in, V4 accum_color, IV2 fb_pixel, IV2 fb_dims, float min_sample_weight,
                                          int sample_batch_size, V4 *store) {
    V2 motions[3][3], edge_motions[3][3];
    for (int ny = -1; ny <= 1; ++ny)
        for (int nx = -1; nx <= 1; ++nx) {
            const V4 t = fetch(in.mj, IV2{fb_pixel.x + nx, fb_pixel.y + ny});
            motions[nx + 1][ny + 1] = V2{t.x, t.y};
            edge_motions[nx + 1][ny + 1] = motions[nx + 1][ny + 1];
        }
    // REPROJECTION_ACCUM_BOUNDARY_SEARCH
    for (int oy = -2; oy <= 2; ++oy)
        for (int ox = -2; ox <= 2; ++ox) {
            const V4 t = fetch(in.mj, IV2{fb_pixel.x + ox, fb_pixel.y + oy});
            const V2 m{t.x, t.y};
            const float ml = dot(m, m);
            for (int ny = -1; ny <= 1; ++ny)
                for (int nx = -1; nx <= 1; ++nx)
                    if (iabs(ox - nx) <= 1 && iabs(oy - ny) <= 1) {
                        const V2 cm = edge_motions[nx + 1][ny + 1];
                        const float cml = dot(cm, cm);
                        if (ml > cml) edge_motions[nx + 1][ny + 1] = m;
                    }
        }
    for (int ny = -1; ny <= 1; ++ny)
        for (int nx = -1; nx <= 1; ++nx) {
            const V2 starting_point{((float)(fb_pixel.x + nx) + 0.5f) / (float)fb_dims.x, ((float)(fb_pixel.y + ny) + 0.5f) / (float)fb_dims.y};
            V2 reconstruction_point{starting_point.x + 0.5f * motions[nx + 1][ny + 1].x, starting_point.y + 0.5f * motions[nx + 1][ny + 1].y};
            // vec2 anchor_point = ivec2(starting_point + 0.5f * edge_motions[..]): truncated in normalised coordinates, as written
            const V2 anchor_point{(float)to_int(starting_point.x + 0.5f * edge_motions[nx + 1][ny + 1].x),
                                  (float)to_int(starting_point.y + 0.5f * edge_motions[nx + 1][ny + 1].y)};
            const V2 anchor_min{floorf(anchor_point.x) - 0.5f, floorf(anchor_point.y) - 0.5f};
            const V2 anchor_max{floorf(anchor_point.x) + 1.5f, floorf(anchor_point.y) + 1.5f};
            reconstruction_point.x = clampf(reconstruction_point.x, anchor_min.x, anchor_max.x);
            reconstruction_point.y = clampf(reconstruction_point.y, anchor_min.y, anchor_max.y);
            motions[nx + 1][ny + 1] = V2{2.0f * (reconstruction_point.x - starting_point.x), 2.0f * (reconstruction_point.y - starting_point.y)};
        }
    const V2 starting_point{((float)fb_pixel.x + 0.5f) / (float)fb_dims.x, ((float)fb_pixel.y + 0.5f) / (float)fb_dims.y};
    const V2 reconstruction_point{starting_point.x + 0.5f * motions[1][1].x, starting_point.y + 0.5f * motions[1][1].y};
    const V2 motion_px{(float)fb_dims.x * 0.5f * motions[1][1].x, (float)fb_dims.y * 0.5f * motions[1][1].y};
    float motion_rate = fmaxf(fabsf(motion_px.x), fabsf(motion_px.y)) / min_sample_weight;
    motion_rate *= 0.5f; // read by REPROJECTION_ACCUM_BACKGROUND / BILATERAL_TEST only
    (void)motion_rate;

    V4 history_color{0, 0, 0, 0};
    const V4 test_result{1.0f, 0.0f, 1.0f, -1.0f};
    float new_sample_weight = 1.0f;
    float old_sample_weight = 0.0f;
    if (reconstruction_point.x >= 0.0f && reconstruction_point.y >= 0.0f && reconstruction_point.x < 1.0f && reconstruction_point.y < 1.0f) {
        history_color = texture_linear(in.history, reconstruction_point);
        old_sample_weight = 1.0f - history_color.w;
        if (old_sample_weight > 0.0f) new_sample_weight = old_sample_weight / (1.0f + old_sample_weight * (float)sample_batch_size);
    }
    new_sample_weight = fmaxf(new_sample_weight, min_sample_weight);
    if (accum_color.w > 1.0f) new_sample_weight = 0.95f;

    const V4 current_normal_depth = fetch(in.nd, fb_pixel);
    if (new_sample_weight < 1.0f) { // REPROJECTION_ACCUM_BILATERAL
        const IV2 reconstruction_pixel{to_int(reconstruction_point.x * (float)fb_dims.x), to_int(reconstruction_point.y * (float)fb_dims.y)};
        // REPROJECTION_ACCUM_FIT_GEOMETRY_DISTRIBUTION
        V3 avg_normal{0, 0, 0};
        float avg_depth = 0.0f, sq_depth = 0.0f, min_depth = 2.e32f, max_depth = 0.0f;
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const V4 recons_normal_depth = fetch(in.nd, IV2{fb_pixel.x + ox, fb_pixel.y + oy});
                avg_normal = avg_normal + V3{recons_normal_depth.x, recons_normal_depth.y, recons_normal_depth.z};
                const float rel_depth = recons_normal_depth.w / current_normal_depth.w;
                avg_depth += rel_depth;
                sq_depth += rel_depth * rel_depth;
                min_depth = fminf(min_depth, rel_depth);
                max_depth = fmaxf(max_depth, rel_depth);
            }
        avg_normal = avg_normal / 9.0f;
        avg_depth /= 9.0f;
        sq_depth /= 9.0f;
        const float normal_sigma = fmaxf(1.0f - length(avg_normal), 0.0f);
        const float depth_sigma = sqrtf(fmaxf(sq_depth - avg_depth * avg_depth, 0.0f));

        float bilateral_weight = 0.0f, mix_weight = 0.0f, max_weight = 0.0f;
        V4 mix_history_color{0, 0, 0, 0};
        float old_sample_weight_b = 0.0f; // the block's own `old_sample_weight`, shadowing the outer one
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const IV2 at{reconstruction_pixel.x + ox, reconstruction_pixel.y + oy};
                const V4 neighbor_history_color = fetch(in.history, at);
                const float neighbor_old_sample_weight = 1.0f - neighbor_history_color.w;
                const V4 recons_normal_depth = fetch(in.nd_history, at);
                const float angle = dot(V3{recons_normal_depth.x, recons_normal_depth.y, recons_normal_depth.z},
                                        V3{current_normal_depth.x, current_normal_depth.y, current_normal_depth.z});
                const float rcp_depth_delta = fabsf(recons_normal_depth.w / current_normal_depth.w - 1.0f);
                float weight = smoothstep(-0.66f, 1.0f, angle + normal_sigma) *
                               fminf(fmaxf(0.0f, 1.0f - fminf(10.0f, 1.0f / depth_sigma) * rcp_depth_delta), 1.0f);
                max_weight = fmaxf(max_weight, weight);
                bilateral_weight += weight; // the log-space sums next to it (:264-266) feed sigma_ldr, which only BILATERAL_TEST reads
                const V2 d{((float)at.x + 0.5f) - reconstruction_point.x * (float)fb_dims.x, ((float)at.y + 0.5f) - reconstruction_point.y * (float)fb_dims.y};
                const float filterWeight = exp_f(-3.0f * dot(d, d));
                weight *= filterWeight;
                if (neighbor_old_sample_weight > 0.0f) {
                    mix_weight += weight;
                    mix_history_color.x += weight * neighbor_history_color.x;
                    mix_history_color.y += weight * neighbor_history_color.y;
                    mix_history_color.z += weight * neighbor_history_color.z;
                    mix_history_color.w += weight * neighbor_history_color.w;
                    old_sample_weight_b += weight * weight * neighbor_old_sample_weight;
                }
            }
        (void)bilateral_weight;
        (void)max_weight;
        if (mix_weight > 0.0f) {
            mix_history_color.x /= mix_weight; mix_history_color.y /= mix_weight; mix_history_color.z /= mix_weight; mix_history_color.w /= mix_weight;
            old_sample_weight_b /= mix_weight * mix_weight;
            (void)old_sample_weight_b;
            // REPROJECTION_ACCUM_BILATERAL_PROJECTION
            const V3 line{history_color.x - accum_color.x, history_color.y - accum_color.y, history_color.z - accum_color.z};
            const V3 to_mix{mix_history_color.x - accum_color.x, mix_history_color.y - accum_color.y, mix_history_color.z - accum_color.z};
            const float t = dot(to_mix, line) / dot(line, line);
            new_sample_weight = fmaxf(new_sample_weight, 1.0f - fmaxf(t, 0.0f));
        } else {
            new_sample_weight = 1.0f;
        }
    }
    new_sample_weight = fmaxf(new_sample_weight, min_sample_weight);

    history_color.x = history_color.x + (accum_color.x - history_color.x) * new_sample_weight;
    history_color.y = history_color.y + (accum_color.y - history_color.y) * new_sample_weight;
    history_color.z = history_color.z + (accum_color.z - history_color.z) * new_sample_weight;
    history_color.w = history_color.w + (accum_color.w - history_color.w) * new_sample_weight;
    history_color.w = 1.0f - new_sample_weight;

    const V3 mixed = mix(V3{accum_color.x, accum_color.y, accum_color.z}, V3{history_color.x, history_color.y, history_color.z}, 1.0f);
    accum_color.x = mixed.x; accum_color.y = mixed.y; accum_color.z = mixed.z;
    *store = V4{accum_color.x, accum_color.y, accum_color.z, history_color.w};
    if (test_result.w >= 0.0f) return test_result;
    return accum_color;
}

// ---- process_taa.comp ---------------------------------------------------------------------------------------------------------
static inline float sin_any(float x) {
    float s, c;
    sincos_pos(fabsf(x), s, c);
    return x < 0.0f ? -s : s;
}
static inline float lanczosWeight(float x, float r) {
    if (x == 0.0f) return 1.0f;
    return r * sin_any(x * PI_F) * sin_any((x / r) * PI_F) / (PI_F * PI_F * x * x);
}
static inline V4 lanczos(Image4b history_framebuffer, V2 coord, int r, IV2 fb_dims, int render_upscale_factor) {
    const V2 point{coord.x * (float)fb_dims.x - 0.5f, coord.y * (float)fb_dims.y - 0.5f};
    const V2 cpoint{ceilf(point.x), ceilf(point.y)};
    V4 accum{0, 0, 0, 0};
    float total = 0.0f;
    for (int oy = -r; oy < r; ++oy)
        for (int ox = -r; ox < r; ++ox) {
            const V2 npoint{(float)(render_upscale_factor * ox) + cpoint.x, (float)(render_upscale_factor * oy) + cpoint.y};
            const float weight = lanczosWeight((npoint.x - point.x) / (float)render_upscale_factor, (float)r) *
                                 lanczosWeight((npoint.y - point.y) / (float)render_upscale_factor, (float)r);
            const V4 t = fetch(history_framebuffer, IV2{to_int(npoint.x), to_int(npoint.y)});
            accum.x += weight * t.x; accum.y += weight * t.y; accum.z += weight * t.z; accum.w += weight * t.w;
            total += weight;
        }
    return V4{accum.x / total, accum.y / total, accum.z / total, accum.w / total};
}
static inline uint8_t to_unorm8(float x) { return (uint8_t)(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f + 0.5f); }

// framebuffer = the target as process_samples wrote it (every read sees that state), out = the pixel the invocation stores
static inline void process_taa_main(Image4b framebuffer, Image4b history_framebuffer, Image4h aov_motion_jitter_buffer, IV2 fb_pixel, IV2 fb_dims,
                                    int render_upscale_factor, uint8_t *out) {
    V4 accum_color = fetch(framebuffer, fb_pixel);
    const IV2 mpx{fb_pixel.x / render_upscale_factor, fb_pixel.y / render_upscale_factor};
    V4 t = fetch(aov_motion_jitter_buffer, mpx);
    V2 motion{t.x, t.y};
    float motion_len = dot(motion, motion);
    for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
            const V4 tm = fetch(aov_motion_jitter_buffer, mpx); // the shader does not add the offset
            const V2 m{tm.x, tm.y};
            const float ml = dot(m, m);
            if (ml > motion_len) { motion = m; motion_len = ml; }
        }
    const V2 starting_point{((float)fb_pixel.x + 0.5f) / (float)fb_dims.x, ((float)fb_pixel.y + 0.5f) / (float)fb_dims.y};
    const V2 reconstruction_point{starting_point.x + 0.5f * motion.x, starting_point.y + 0.5f * motion.y};
    V4 history_color{0, 0, 0, 0};
    float new_sample_weight = 1.0f;
    if (reconstruction_point.x >= 0.0f && reconstruction_point.y >= 0.0f && reconstruction_point.x <= 1.0f && reconstruction_point.y <= 1.0f) {
        history_color = lanczos(history_framebuffer, reconstruction_point, 5, fb_dims, render_upscale_factor);
        new_sample_weight = 0.15f;
    }
    if (new_sample_weight < 1.0f) {
        float trim[4] = {0, 0, 0, 0}, max2[4] = {0, 0, 0, 0};
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const V4 val = fetch(framebuffer, IV2{fb_pixel.x + ox * render_upscale_factor, fb_pixel.y + oy * render_upscale_factor});
                const float v[4] = {val.x, val.y, val.z, val.w};
                for (int c = 0; c < 4; ++c) { trim[c] += v[c]; max2[c] += v[c] * v[c]; }
            }
        float a[4] = {accum_color.x, accum_color.y, accum_color.z, accum_color.w};
        const float hc[4] = {history_color.x, history_color.y, history_color.z, history_color.w};
        for (int c = 0; c < 4; ++c) {
            trim[c] /= 9.0f;
            max2[c] /= 9.0f;
            max2[c] = sqrtf(max2[c]);
            const float stddev = 9.0f / 8.0f * (max2[c] - trim[c]);
            const float trim_low = fmaxf(0.0f, trim[c] - stddev);
            const float trim_high = fmaxf(trim[c] + 3.0f * stddev, a[c] + stddev);
            a[c] = hc[c] + (a[c] - hc[c]) * new_sample_weight;
            a[c] = fminf(fmaxf(a[c], trim_low), trim_high);
        }
        accum_color = V4{a[0], a[1], a[2], a[3]};
    }
    out[0] = to_unorm8(accum_color.x); out[1] = to_unorm8(accum_color.y); out[2] = to_unorm8(accum_color.z); out[3] = to_unorm8(accum_color.w);
}

} // namespace post
