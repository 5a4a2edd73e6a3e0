// oracle/fp32.h -- TEST INFRASTRUCTURE (CPU oracle). Never included by the product.
//
// The fp32 arithmetic contract ("RPTR-FP") as scalar host code.  The GLSL reference leaves
// operation order, FMA contraction and transcendental accuracy to the driver; a path tracer is chaotic in those
// choices (hard part 1), so oracle and CUDA kernels both implement ONE written-down choice:
//   * every operation is IEEE-754 binary32 round-to-nearest-even; no contraction except where fmaf() is written
//     (build with -ffp-contract=off; the CUDA side builds with -fmad=false);
//   * dot/cross/matrix products are fixed fma chains (below);
//   * sin/cos/exp/acos are the polynomial kernels below (built only from +,-,*,/,fma,sqrt), not libm.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace fp {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 v3(float s) { return V3{s, s, s}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

// dot(a,b) := fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))
static inline float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline float dot(V2 a, V2 b) { return fmaf(a.y, b.y, a.x * b.x); }
// cross(a,b).x := fma(a.y,b.z, -(a.z*b.y)) (cyclic)
static inline V3 cross(V3 a, V3 b) {
    return V3{fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
static inline float length(V3 a) { return sqrtf(dot(a, a)); }
// normalize(v) := v * (1 / sqrt(dot(v,v)))
static inline V3 normalize(V3 a) {
    float inv = 1.0f / sqrtf(dot(a, a));
    return a * inv;
}
// mix(x,y,a) := fma(y, a, x*(1-a))
static inline float mix(float x, float y, float a) { return fmaf(y, a, x * (1.0f - a)); }
static inline V3 mix(V3 x, V3 y, float a) { return V3{mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)}; }
// mat3(c0,c1,c2) * v := fma(c2, v.z, fma(c1, v.y, c0*v.x)) per component
static inline V3 mat_mul(V3 c0, V3 c1, V3 c2, V3 v) {
    return V3{fmaf(c2.x, v.z, fmaf(c1.x, v.y, c0.x * v.x)), fmaf(c2.y, v.z, fmaf(c1.y, v.y, c0.y * v.x)),
              fmaf(c2.z, v.z, fmaf(c1.z, v.y, c0.z * v.x))};
}
// reflect(I,N) := I - N * (2*dot(N,I))
static inline V3 reflect(V3 i, V3 n) { return i - n * (2.0f * dot(n, i)); }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float luminance(V3 c) { return fmaf(0.0722f, c.z, fmaf(0.7152f, c.y, 0.2126f * c.x)); }
static inline float max3(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static const float PI_F = 3.14159265358979323846f;
static const float INV_PI_F = 0.318309886183790671538f;
static const float TWO_PI_F = 6.28318530717958647692f;

// sincos for x >= 0 (arguments on the path are 2*pi*u or half solid angles, all in [0, ~6.3]).
// Cody-Waite reduction by pi/2 with two fma steps, cephes sinf/cosf minimax kernels on [-pi/4, pi/4].
static inline void sincos_pos(float x, float &s, float &c) {
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
    switch (k & 3) {
    case 0: s = sn; c = cs; break;
    case 1: s = cs; c = -sn; break;
    case 2: s = -sn; c = -cs; break;
    default: s = -cs; c = sn; break;
    }
}

// exp(x): k = round(x*log2(e)), two-step Cody-Waite by ln2, cephes expf polynomial, exact 2^k scaling.
static inline float exp_f(float x) {
    if (!(x > -87.0f)) return 0.0f; // also maps NaN to 0
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

// acos(x) for x in [-1,1]: cephes asinf kernel.
static inline float asin_kernel(float z) {
    float p = fmaf(z, 4.2163199048e-2f, 2.4181311049e-2f);
    p = fmaf(z, p, 4.5470025998e-2f);
    p = fmaf(z, p, 7.4953002686e-2f);
    p = fmaf(z, p, 1.6666752422e-1f);
    return p;
}
static inline float acos_f(float x) {
    float a = fabsf(x);
    if (a > 0.5f) {
        float z = 0.5f * (1.0f - a);
        float s = sqrtf(z);
        float r = 2.0f * fmaf(s * z, asin_kernel(z), s);
        return x > 0.0f ? r : PI_F - r;
    }
    float z = x * x;
    float as = fmaf(x * z, asin_kernel(z), x);
    return 1.57079637050628662109375f - as;
}

} // namespace fp
