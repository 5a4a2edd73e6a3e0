// rptr_post.cuh -- the temporal part of the post-accumulate chain: what the reference builds with
// ENABLE_REALTIME_RESOLVE (CMakeLists.txt:98, OFF by default; option "realtime_resolve" here):
//   reproject_and_accumulate   rendering/postprocess/reprojection.glsl:44-367, called from vulkan/process_samples.comp:106-113 when
//                              reprojection_mode == REPROJECTION_MODE_ACCUMULATE and the frame is not the first since a reset
//   process_taa                vulkan/processing/process_taa.comp:28-112 (Lanczos-resampled LDR history, variance clamp)
// Per-pixel functions, __host__ __device__ like the shading code, so that tests/hostsim runs the very same statements on the CPU.
//
// Restated for the configuration the reference ships (reprojection.glsl:17-24): BOUNDARY_SEARCH, BILATERAL,
// BILATERAL_PROJECTION and FIT_GEOMETRY_DISTRIBUTION defined; CONFLICT_RESOLUTION, BILATERAL_TEST, BACKGROUND, ACCUM_GBUFFER
// and TEST_BILATERAL_ACCUM_GUESS not.  Statements whose results nothing reads in that configuration are left out (the motion of
// the eight neighbours after the boundary search, motion_rate, the mean / variance of the 3x3 accumulator neighbourhood, the
// log-space statistics of the bilateral history): what remains reads only images no invocation of the pass writes, so -- unlike
// the shader as a whole -- it has no race.  What the driver leaves open is fixed as follows (RPTR-FP):
//   * image loads / texel fetches outside the image return zero;
//   * textureLod(history, p, 0) through screen_sampler (render_vulkan.cpp:417-427: LINEAR, CLAMP_TO_EDGE) is the exact fp32
//     bilinear blend of the four nearest texel centres, mix(mix(t00, t10, ax), mix(t01, t11, ax), ay);
//   * float -> int conversions truncate, NaN converts to 0; min / max return the operand that is not NaN;
//   * exp, sin are the RPTR-FP kernels of rptr_math.cuh; dot(vec2) = fma(a.y, b.y, a.x * b.x);
//   * rgba8 stores round to nearest (x * 255 + 0.5, truncated), rgba8 loads return v / 255.
#pragma once
#include "rptr_math.cuh"

namespace rp {

RPTR_HD float half_bits_to_float(uint32_t h) {
    const uint32_t s = (h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u;
    if (e == 0u) {
        const float f = (float)m * 5.9604644775390625e-08f; // subnormal: m * 2^-24, exact
        return s ? -f : f;
    }
    if (e == 31u) return u2f(s | 0x7f800000u | (m << 13));
    return u2f(s | ((e + 112u) << 23) | (m << 13));
}
RPTR_HD float4 half4_bits_to_float4(ushort4 h) {
    return f4(half_bits_to_float(h.x), half_bits_to_float(h.y), half_bits_to_float(h.z), half_bits_to_float(h.w));
}
RPTR_HD int trunc_to_int(float x) { // NaN -> 0, saturating: the conversions of the shader are never asked for more
    if (!(x == x)) return 0;
    if (x >= 2147483520.0f) return 2147483520;
    if (x <= -2147483520.0f) return -2147483520;
    return (int)x;
}

struct ResolveImages {
    int32_t w, h;
    const float4 *history;     // accumulator of the previous frame (rgb, 1 - its sample weight)
    const ushort4 *nd_history; // normal + depth AOV of the previous frame
    const ushort4 *nd;         // normal + depth AOV of this frame
    const ushort4 *mj;         // motion + jitter AOV of this frame
};

RPTR_HD float4 load_f4(const float4 *img, int32_t w, int32_t h, int x, int y) {
    if (x < 0 || y < 0 || x >= w || y >= h) return f4(0.0f, 0.0f, 0.0f, 0.0f);
    return img[(size_t)y * (size_t)w + (size_t)x];
}
RPTR_HD float4 load_h4(const ushort4 *img, int32_t w, int32_t h, int x, int y) {
    if (x < 0 || y < 0 || x >= w || y >= h) return f4(0.0f, 0.0f, 0.0f, 0.0f);
    return half4_bits_to_float4(img[(size_t)y * (size_t)w + (size_t)x]);
}
RPTR_HD float4 sample_linear_clamp(const float4 *img, int32_t w, int32_t h, float u, float v) {
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float ax = x - x0f, ay = y - y0f;
    const int xi = trunc_to_int(x0f), yi = trunc_to_int(y0f);
    const int x0 = xi < 0 ? 0 : (xi > w - 1 ? w - 1 : xi), x1 = xi + 1 < 0 ? 0 : (xi + 1 > w - 1 ? w - 1 : xi + 1);
    const int y0 = yi < 0 ? 0 : (yi > h - 1 ? h - 1 : yi), y1 = yi + 1 < 0 ? 0 : (yi + 1 > h - 1 ? h - 1 : yi + 1);
    const float4 t00 = img[(size_t)y0 * w + x0], t10 = img[(size_t)y0 * w + x1], t01 = img[(size_t)y1 * w + x0], t11 = img[(size_t)y1 * w + x1];
    return f4(mixf(mixf(t00.x, t10.x, ax), mixf(t01.x, t11.x, ax), ay), mixf(mixf(t00.y, t10.y, ax), mixf(t01.y, t11.y, ax), ay),
              mixf(mixf(t00.z, t10.z, ax), mixf(t01.z, t11.z, ax), ay), mixf(mixf(t00.w, t10.w, ax), mixf(t01.w, t11.w, ax), ay));
}
RPTR_HD float smoothstep_f(float e0, float e1, float x) {
    const float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
RPTR_HD float dot2(float ax, float ay, float bx, float by) { return fmaf(ay, by, ax * bx); }

// reprojection.glsl:44-367.  accum_color = what the megakernel left in the accumulator for this pixel (this frame's sample);
// *stored = what the pass writes back (rgb, 1 - sample weight); the return value goes on to the display chain (rgb, alpha of the sample).
RPTR_HD float4 reproject_and_accumulate(const ResolveImages &im, float4 accum_color, int px, int py, float min_sample_weight, int sample_batch_size,
                                        float4 *stored) {
    const float fw = (float)im.w, fh = (float)im.h;
    // :51-71 boundary search, for the centre pixel: the largest motion of the 3x3 neighbourhood (first one wins ties)
    const float4 m0 = load_h4(im.mj, im.w, im.h, px, py);
    float mx = m0.x, my = m0.y, ex = m0.x, ey = m0.y;
    for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
            const float4 m = load_h4(im.mj, im.w, im.h, px + ox, py + oy);
            if (dot2(m.x, m.y, m.x, m.y) > dot2(ex, ey, ex, ey)) { ex = m.x; ey = m.y; }
        }
    const float spx = ((float)px + 0.5f) / fw, spy = ((float)py + 0.5f) / fh;
    {   // :72-84 clip the motion target to the anchor box of the edge motion (the anchor is truncated in NORMALISED coordinates, as written)
        float rpx = spx + 0.5f * mx, rpy = spy + 0.5f * my;
        const float apx = (float)trunc_to_int(spx + 0.5f * ex), apy = (float)trunc_to_int(spy + 0.5f * ey);
        rpx = clampf(rpx, floorf(apx) - 0.5f, floorf(apx) + 1.5f);
        rpy = clampf(rpy, floorf(apy) - 0.5f, floorf(apy) + 1.5f);
        mx = 2.0f * (rpx - spx);
        my = 2.0f * (rpy - spy);
    }
    const float rpx = spx + 0.5f * mx, rpy = spy + 0.5f * my; // :87
    float4 history_color = f4(0.0f, 0.0f, 0.0f, 0.0f);
    float new_sample_weight = 1.0f;
    if (rpx >= 0.0f && rpy >= 0.0f && rpx < 1.0f && rpy < 1.0f) { // :100-155
        history_color = sample_linear_clamp(im.history, im.w, im.h, rpx, rpy);
        const float old_sample_weight = 1.0f - history_color.w;
        if (old_sample_weight > 0.0f) new_sample_weight = old_sample_weight / (1.0f + old_sample_weight * (float)sample_batch_size);
    }
    new_sample_weight = fmaxf(new_sample_weight, min_sample_weight);
    if (accum_color.w > 1.0f) new_sample_weight = 0.95f; // :159-160 non-accumulation object types
    const float4 cur_nd = load_h4(im.nd, im.w, im.h, px, py);
    if (new_sample_weight < 1.0f) { // :165-335 normal / depth based history invalidation
        const int rx = trunc_to_int(rpx * fw), ry = trunc_to_int(rpy * fh);
        float3 avg_normal = f3(0.0f);
        float avg_depth = 0.0f, sq_depth = 0.0f;
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const float4 r = load_h4(im.nd, im.w, im.h, px + ox, py + oy);
                avg_normal = avg_normal + f3(r.x, r.y, r.z);
                const float rel_depth = r.w / cur_nd.w;
                avg_depth += rel_depth;
                sq_depth += rel_depth * rel_depth;
            }
        avg_normal = avg_normal / 9.0f;
        avg_depth /= 9.0f;
        sq_depth /= 9.0f;
        const float normal_sigma = fmaxf(1.0f - length(avg_normal), 0.0f);
        const float depth_sigma = sqrtf(fmaxf(sq_depth - avg_depth * avg_depth, 0.0f));
        float mix_weight = 0.0f;
        float4 mix_history = f4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const float4 nh = load_f4(im.history, im.w, im.h, rx + ox, ry + oy);
                const float neighbor_old_sample_weight = 1.0f - nh.w;
                const float4 rnd = load_h4(im.nd_history, im.w, im.h, rx + ox, ry + oy);
                const float angle = dot(f3(rnd.x, rnd.y, rnd.z), f3(cur_nd.x, cur_nd.y, cur_nd.z));
                const float rcp_depth_delta = fabsf(rnd.w / cur_nd.w - 1.0f);
                float weight = smoothstep_f(-0.66f, 1.0f, angle + normal_sigma) *
                               fminf(fmaxf(0.0f, 1.0f - fminf(10.0f, 1.0f / depth_sigma) * rcp_depth_delta), 1.0f);
                const float dx = ((float)(rx + ox) + 0.5f) - rpx * fw, dy = ((float)(ry + oy) + 0.5f) - rpy * fh;
                weight *= exp_f(-3.0f * dot2(dx, dy, dx, dy));
                if (neighbor_old_sample_weight > 0.0f) {
                    mix_weight += weight;
                    mix_history.x += weight * nh.x; mix_history.y += weight * nh.y; mix_history.z += weight * nh.z; mix_history.w += weight * nh.w;
                }
            }
        if (mix_weight > 0.0f) { // :287-325, REPROJECTION_ACCUM_BILATERAL_PROJECTION
            const float3 mh = f3(mix_history.x / mix_weight, mix_history.y / mix_weight, mix_history.z / mix_weight);
            const float3 a = f3(accum_color.x, accum_color.y, accum_color.z);
            const float3 line = f3(history_color.x, history_color.y, history_color.z) - a;
            const float t = dot(mh - a, line) / dot(line, line);
            new_sample_weight = fmaxf(new_sample_weight, 1.0f - fmaxf(t, 0.0f));
        } else
            new_sample_weight = 1.0f;
    }
    new_sample_weight = fmaxf(new_sample_weight, min_sample_weight);
    history_color.x = history_color.x + (accum_color.x - history_color.x) * new_sample_weight; // :340-341
    history_color.y = history_color.y + (accum_color.y - history_color.y) * new_sample_weight;
    history_color.z = history_color.z + (accum_color.z - history_color.z) * new_sample_weight;
    history_color.w = 1.0f - new_sample_weight;
    const float3 out = mix3(f3(accum_color.x, accum_color.y, accum_color.z), f3(history_color.x, history_color.y, history_color.z), 1.0f); // :343
    *stored = f4(out.x, out.y, out.z, history_color.w);
    return f4(out.x, out.y, out.z, accum_color.w);
}

// ---- TAA on the LDR target (process_taa.comp) -----------------------------------------------------------------------------
struct TaaImages {
    int32_t w, h;         // LDR target = render size x upscale
    int32_t upscale;
    int32_t rw, rh;       // render size (motion image)
    const uchar4 *current; // this frame's LDR image as process_samples wrote it (a snapshot: the pass writes a different buffer)
    const uchar4 *history; // the previous frame's LDR target after its own TAA pass
    const ushort4 *mj;
};
RPTR_HD float4 load_u8(const uchar4 *img, int32_t w, int32_t h, int x, int y) {
    if (x < 0 || y < 0 || x >= w || y >= h) return f4(0.0f, 0.0f, 0.0f, 0.0f);
    const uchar4 c = img[(size_t)y * (size_t)w + (size_t)x];
    return f4((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f, (float)c.w / 255.0f);
}
RPTR_HD float sin_signed(float x) {
    float s, c;
    sincos_pos(fabsf(x), s, c);
    return x < 0.0f ? -s : s;
}
RPTR_HD float lanczos_weight(float x, float r) { // :28-31
    if (x == 0.0f) return 1.0f;
    return r * sin_signed(x * RPTR_PI) * sin_signed((x / r) * RPTR_PI) / (RPTR_PI * RPTR_PI * x * x);
}
RPTR_HD unsigned char unorm8(float x) { return (unsigned char)(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f + 0.5f); }

RPTR_HD uchar4 process_taa_pixel(const TaaImages &im, int px, int py) {
    float4 accum = load_u8(im.current, im.w, im.h, px, py);
    // :61-71 the neighbourhood loop of the shader re-reads the same texel nine times: the motion is the pixel's own
    const float4 mj = load_h4(im.mj, im.rw, im.rh, px / im.upscale, py / im.upscale);
    const float fw = (float)im.w, fh = (float)im.h;
    const float rpx = ((float)px + 0.5f) / fw + 0.5f * mj.x, rpy = ((float)py + 0.5f) / fh + 0.5f * mj.y;
    float4 history = f4(0.0f, 0.0f, 0.0f, 0.0f);
    float new_sample_weight = 1.0f;
    if (rpx >= 0.0f && rpy >= 0.0f && rpx <= 1.0f && rpy <= 1.0f) { // :78-83, lanczos(reconstruction_point, 5): :35-52
        const float ptx = rpx * fw - 0.5f, pty = rpy * fh - 0.5f;
        const float cx = ceilf(ptx), cy = ceilf(pty);
        float total = 0.0f;
        for (int oy = -5; oy < 5; ++oy)
            for (int ox = -5; ox < 5; ++ox) {
                const float nx = (float)(im.upscale * ox) + cx, ny = (float)(im.upscale * oy) + cy;
                const float weight = lanczos_weight((nx - ptx) / (float)im.upscale, 5.0f) * lanczos_weight((ny - pty) / (float)im.upscale, 5.0f);
                const float4 t = load_u8(im.history, im.w, im.h, trunc_to_int(nx), trunc_to_int(ny));
                history.x += weight * t.x; history.y += weight * t.y; history.z += weight * t.z; history.w += weight * t.w;
                total += weight;
            }
        history.x /= total; history.y /= total; history.z /= total; history.w /= total;
        new_sample_weight = 0.15f;
    }
    if (new_sample_weight < 1.0f) { // :86-106 variance clamp against the 3x3 neighbourhood (stride = upscale factor)
        float4 trim = f4(0.0f, 0.0f, 0.0f, 0.0f), max2 = f4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const float4 v = load_u8(im.current, im.w, im.h, px + ox * im.upscale, py + oy * im.upscale);
                trim.x += v.x; trim.y += v.y; trim.z += v.z; trim.w += v.w;
                max2.x += v.x * v.x; max2.y += v.y * v.y; max2.z += v.z * v.z; max2.w += v.w * v.w;
            }
        float a[4] = {accum.x, accum.y, accum.z, accum.w};
        const float tr[4] = {trim.x, trim.y, trim.z, trim.w}, m2[4] = {max2.x, max2.y, max2.z, max2.w};
        const float hc[4] = {history.x, history.y, history.z, history.w};
        for (int c = 0; c < 4; ++c) {
            const float mean = tr[c] / 9.0f;
            const float rms = sqrtf(m2[c] / 9.0f);
            const float stddev = 9.0f / 8.0f * (rms - mean);
            const float low = fmaxf(0.0f, mean - stddev);
            const float high = fmaxf(mean + 3.0f * stddev, a[c] + stddev);
            const float blended = hc[c] + (a[c] - hc[c]) * new_sample_weight;
            a[c] = fminf(fmaxf(blended, low), high);
        }
        accum = f4(a[0], a[1], a[2], a[3]);
    }
    uchar4 o;
    o.x = unorm8(accum.x); o.y = unorm8(accum.y); o.z = unorm8(accum.z); o.w = unorm8(accum.w);
    return o;
}

} // namespace rp
