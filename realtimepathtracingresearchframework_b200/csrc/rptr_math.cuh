// rptr_math.cuh -- fp32 arithmetic contract ("RPTR-FP") for the CUDA wavefront.
//
// The reference's GLSL leaves operation order, FMA contraction and the accuracy of sin/cos/exp/acos/pow to the Vulkan
// driver.  This backend fixes ONE choice so that a frame is a pure function of (scene, camera, counters):
//   * IEEE-754 binary32, round-to-nearest-even everywhere; this file is compiled with -fmad=false (device) and
//     -ffp-contract=off (host), so the only fused operations are the fmaf() calls written below;
//   * dot / cross / matrix-vector products are fixed fma chains;
//   * transcendental functions are polynomial kernels made of + - * / fma sqrt only.
// All functions are __host__ __device__ so the host-side unit tests (tests/hostsim) can execute the exact code the
// kernels run.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector_types.h>

#if defined(__CUDACC__)
#define RPTR_HD __host__ __device__ __forceinline__
#else
#define RPTR_HD inline
#endif

namespace rp {

RPTR_HD float3 f3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
RPTR_HD float3 f3(float s) { return f3(s, s, s); }
RPTR_HD float2 f2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
RPTR_HD float4 f4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
RPTR_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
RPTR_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
RPTR_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
RPTR_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
RPTR_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
RPTR_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
RPTR_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
RPTR_HD bool is_zero(float3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

RPTR_HD float dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
RPTR_HD float3 cross(float3 a, float3 b) {
    return f3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
RPTR_HD float length(float3 a) { return sqrtf(dot(a, a)); }
RPTR_HD float3 normalize(float3 a) {
    float inv = 1.0f / sqrtf(dot(a, a));
    return a * inv;
}
RPTR_HD float mixf(float x, float y, float a) { return fmaf(y, a, x * (1.0f - a)); }
RPTR_HD float3 mix3(float3 x, float3 y, float a) { return f3(mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)); }
// mat3(c0,c1,c2) * v
RPTR_HD float3 mat_mul(float3 c0, float3 c1, float3 c2, float3 v) {
    return f3(fmaf(c2.x, v.z, fmaf(c1.x, v.y, c0.x * v.x)), fmaf(c2.y, v.z, fmaf(c1.y, v.y, c0.y * v.x)),
              fmaf(c2.z, v.z, fmaf(c1.z, v.y, c0.z * v.x)));
}
RPTR_HD float3 reflect3(float3 i, float3 n) { return i - n * (2.0f * dot(n, i)); }
RPTR_HD float3 refract3(float3 i, float3 n, float eta) {
    float d = dot(n, i);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return f3(0.0f);
    return i * eta - n * (eta * d + sqrtf(k));
}
RPTR_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
RPTR_HD float luminance(float3 c) { return fmaf(0.0722f, c.z, fmaf(0.7152f, c.y, 0.2126f * c.x)); }
RPTR_HD float3 abs3(float3 a) { return f3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
RPTR_HD float3 max0(float3 a) { return f3(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f), fmaxf(a.z, 0.0f)); }
RPTR_HD float pow2(float x) { return x * x; }

RPTR_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RPTR_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

#define RPTR_PI 3.14159265358979323846f
#define RPTR_INV_PI 0.318309886183790671538f
#define RPTR_TWO_PI 6.28318530717958647692f

// sin/cos for x >= 0 (2*pi*u and half solid angles): Cody-Waite by pi/2, minimax kernels on [-pi/4, pi/4]
RPTR_HD void sincos_pos(float x, float &s, float &c) {
    int k = (int)(x * 0.636619772367581343f + 0.5f);
    float fk = (float)k;
    float r = fmaf(-fk, 1.57079637050628662109375f, x);
    r = fmaf(-fk, -4.37113900018624283e-8f, r);
    float r2 = r * r;
    float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(r2, ps, -1.6666654611e-1f);
    float sn = fmaf(r * r2, ps, r);
    float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(r2, pc, 4.166664568298827e-2f);
    float cs = fmaf(r2 * r2, pc, fmaf(-0.5f, r2, 1.0f));
    int q = k & 3;
    float so = (q & 1) ? cs : sn;
    float co = (q & 1) ? sn : cs;
    s = (q & 2) ? -so : so;
    c = (q == 1 || q == 2) ? -co : co;
}

RPTR_HD float exp_f(float x) {
    if (!(x > -87.0f)) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float fk = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(-fk, 0.693359375f, x);
    r = fmaf(-fk, -2.12194440e-4f, r);
    float p = fmaf(r, 1.9875691500e-4f, 1.3981999507e-3f);
    p = fmaf(r, p, 8.3334519073e-3f);
    p = fmaf(r, p, 4.1665795894e-2f);
    p = fmaf(r, p, 1.6666665459e-1f);
    p = fmaf(r, p, 5.0000001201e-1f);
    float e = fmaf(r * r, p, r) + 1.0f;
    int k = (int)fk;
    return e * u2f((uint32_t)(k + 127) << 23);
}

RPTR_HD float asin_kernel(float z) {
    float p = fmaf(z, 4.2163199048e-2f, 2.4181311049e-2f);
    p = fmaf(z, p, 4.5470025998e-2f);
    p = fmaf(z, p, 7.4953002686e-2f);
    p = fmaf(z, p, 1.6666752422e-1f);
    return p;
}
RPTR_HD float acos_f(float x) {
    float a = fabsf(x);
    if (a > 0.5f) {
        float z = 0.5f * (1.0f - a);
        float s = sqrtf(z);
        float r = 2.0f * fmaf(s * z, asin_kernel(z), s);
        return x > 0.0f ? r : RPTR_PI - r;
    }
    float z = x * x;
    float as = fmaf(x * z, asin_kernel(z), x);
    return 1.57079637050628662109375f - as;
}

} // namespace rp
