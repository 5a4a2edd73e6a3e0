// oracle/ref_shim/glm/glm.hpp -- TEST INFRASTRUCTURE.
// A small glm-compatible subset (our own code; glm is not installed in the build image) that is just large enough to
// compile the reference's dual-language shading headers as C++ straight from /root/reference, exactly the way
// rendering/tests/compile.cpp does.  Arithmetic is plain C++ float expressions evaluated the way glm 0.9.9 writes
// them (no explicit fma), so results pin the oracle to within a few ulp, not bit-for-bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace glm {

typedef unsigned int uint;

template <class T> struct tvec2 {
    union { T x, r, s; };
    union { T y, g, t; };
    tvec2() : x(0), y(0) {}
    tvec2(T a) : x(a), y(a) {}
    tvec2(T a, T b) : x(a), y(b) {}
    template <class U> explicit tvec2(const tvec2<U> &o) : x(T(o.x)), y(T(o.y)) {}
    T &operator[](int i) { return i == 0 ? x : y; }
    const T &operator[](int i) const { return i == 0 ? x : y; }
};
template <class T> struct tvec4;
template <class T> struct tvec3 {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    tvec3() : x(0), y(0), z(0) {}
    tvec3(T a) : x(a), y(a), z(a) {}
    tvec3(T a, T b_, T c) : x(a), y(b_), z(c) {}
    tvec3(const tvec2<T> &v, T c) : x(v.x), y(v.y), z(c) {}
    template <class U> explicit tvec3(const tvec3<U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
    explicit tvec3(const tvec4<T> &o);
    T &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> struct tvec4 {
    // .xy is the one GLSL swizzle the Z-order Sobol sampler needs (rendering/pointsets/sobol.glsl:174)
    union { struct { union { T x, r; }; union { T y, g; }; }; tvec2<T> xy; };
    union { struct { union { T z, b; }; union { T w, a; }; }; tvec2<T> zw; }; // .zw: rendering/mc/shade_base_material.glsl:64
    tvec4() : x(0), y(0), z(0), w(0) {}
    tvec4(T s) : x(s), y(s), z(s), w(s) {}
    tvec4(T a_, T b_, T c, T d) : x(a_), y(b_), z(c), w(d) {}
    tvec4(const tvec3<T> &v, T d) : x(v.x), y(v.y), z(v.z), w(d) {}
    tvec4(const tvec2<T> &u, const tvec2<T> &v) : x(u.x), y(u.y), z(v.x), w(v.y) {}
    tvec4(T a_, T b_, const tvec2<T> &v) : x(a_), y(b_), z(v.x), w(v.y) {} // vulkan/rt_intersect.comp:56
    template <class U> explicit tvec4(const tvec4<U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}
    T &operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    const T &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};
template <class T> tvec3<T>::tvec3(const tvec4<T> &o) : x(o.x), y(o.y), z(o.z) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;
typedef tvec3<int> ivec3;
typedef tvec4<int> ivec4;
typedef tvec2<uint> uvec2;
typedef tvec3<uint> uvec3;
typedef tvec4<uint> uvec4;
typedef tvec3<bool> bvec3;

#define GLM_SHIM_OP(op) \
    template <class T> inline tvec2<T> operator op(tvec2<T> a, tvec2<T> b) { return tvec2<T>(a.x op b.x, a.y op b.y); } \
    template <class T> inline tvec2<T> operator op(tvec2<T> a, T b) { return tvec2<T>(a.x op b, a.y op b); } \
    template <class T> inline tvec2<T> operator op(T a, tvec2<T> b) { return tvec2<T>(a op b.x, a op b.y); } \
    template <class T> inline tvec3<T> operator op(tvec3<T> a, tvec3<T> b) { return tvec3<T>(a.x op b.x, a.y op b.y, a.z op b.z); } \
    template <class T> inline tvec3<T> operator op(tvec3<T> a, T b) { return tvec3<T>(a.x op b, a.y op b, a.z op b); } \
    template <class T> inline tvec3<T> operator op(T a, tvec3<T> b) { return tvec3<T>(a op b.x, a op b.y, a op b.z); } \
    template <class T> inline tvec4<T> operator op(tvec4<T> a, tvec4<T> b) { return tvec4<T>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    template <class T> inline tvec4<T> operator op(tvec4<T> a, T b) { return tvec4<T>(a.x op b, a.y op b, a.z op b, a.w op b); } \
    template <class T> inline tvec4<T> operator op(T a, tvec4<T> b) { return tvec4<T>(a op b.x, a op b.y, a op b.z, a op b.w); } \
    template <class T, class U> inline tvec2<T> &operator op##=(tvec2<T> &a, U b) { a = a op b; return a; } \
    template <class T, class U> inline tvec3<T> &operator op##=(tvec3<T> &a, U b) { a = a op b; return a; } \
    template <class T, class U> inline tvec4<T> &operator op##=(tvec4<T> &a, U b) { a = a op b; return a; }
GLM_SHIM_OP(+)
GLM_SHIM_OP(-)
GLM_SHIM_OP(*)
GLM_SHIM_OP(/)
#undef GLM_SHIM_OP
// float scalars against double literals, e.g. vec3 * 0.5
inline vec3 operator*(vec3 a, double b) { return a * float(b); }
inline vec3 operator*(double a, vec3 b) { return float(a) * b; }
inline vec3 operator/(vec3 a, double b) { return a / float(b); }
inline vec2 operator*(vec2 a, double b) { return a * float(b); }
inline vec3 operator+(double a, vec3 b) { return float(a) + b; }
inline vec3 operator*(vec3 a, int b) { return a * float(b); }

template <class T> inline tvec2<T> operator-(tvec2<T> a) { return tvec2<T>(-a.x, -a.y); }
template <class T> inline tvec3<T> operator-(tvec3<T> a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <class T> inline tvec4<T> operator-(tvec4<T> a) { return tvec4<T>(-a.x, -a.y, -a.z, -a.w); }
template <class T> inline bool operator==(tvec3<T> a, tvec3<T> b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <class T> inline bool operator!=(tvec3<T> a, tvec3<T> b) { return !(a == b); }
inline uvec3 operator&(uvec3 a, uvec3 b) { return uvec3(a.x & b.x, a.y & b.y, a.z & b.z); }
inline uvec2 operator&(uvec2 a, uvec2 b) { return uvec2(a.x & b.x, a.y & b.y); }
// integer built-ins used by rendering/pointsets/{sobol,sample_order}.glsl
typedef tvec2<bool> bvec2;
inline int bitCount(uint v) { return __builtin_popcount(v); }
inline int findMSB(uint v) { return v ? 31 - __builtin_clz(v) : -1; }
inline ivec2 findMSB(uvec2 v) { return ivec2(findMSB(v.x), findMSB(v.y)); }
inline bvec2 notEqual(uvec2 a, uvec2 b) { return bvec2(a.x != b.x, a.y != b.y); }
inline uvec2 operator<<(uvec2 a, ivec2 b) { return uvec2(a.x << b.x, a.y << b.y); }
inline uvec2 operator<<(uvec2 a, uvec2 b) { return uvec2(a.x << b.x, a.y << b.y); }
inline uvec4 operator>>(uvec4 a, uint b) { return uvec4(a.x >> b, a.y >> b, a.z >> b, a.w >> b); }

// --- scalar built-ins -------------------------------------------------------------------------------------------
using std::abs; using std::sqrt; using std::sin; using std::cos; using std::tan; using std::exp; using std::log;
using std::acos; using std::asin; using std::atan; using std::floor; using std::ceil; using std::isinf; using std::isnan;
inline float pow(float a, float b) { return std::pow(a, b); }
inline float pow(float a, int b) { return std::pow(a, float(b)); }
inline float pow(float a, double b) { return std::pow(a, float(b)); }
inline float fma(float a, float b, float c) { return std::fma(a, b, c); }
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline float min(float a, double b) { return min(a, float(b)); }
inline float max(float a, double b) { return max(a, float(b)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float ldexp(float x, int e) { return std::ldexp(x, e); }
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }

// --- vector built-ins -------------------------------------------------------------------------------------------
#define GLM_SHIM_MAP1(fn) \
    inline vec2 fn(vec2 a) { return vec2(fn(a.x), fn(a.y)); } \
    inline vec3 fn(vec3 a) { return vec3(fn(a.x), fn(a.y), fn(a.z)); } \
    inline vec4 fn(vec4 a) { return vec4(fn(a.x), fn(a.y), fn(a.z), fn(a.w)); }
GLM_SHIM_MAP1(abs) GLM_SHIM_MAP1(sqrt) GLM_SHIM_MAP1(exp) GLM_SHIM_MAP1(log) GLM_SHIM_MAP1(sign) GLM_SHIM_MAP1(floor)
#undef GLM_SHIM_MAP1
inline vec3 pow(vec3 a, vec3 b) { return vec3(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)); }
inline vec2 min(vec2 a, vec2 b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2 max(vec2 a, vec2 b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec2 max(vec2 a, float b) { return vec2(max(a.x, b), max(a.y, b)); }
inline vec3 min(vec3 a, vec3 b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(vec3 a, vec3 b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline uvec3 min(uvec3 a, uvec3 b) { return uvec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline ivec2 clamp(ivec2 v, ivec2 lo, ivec2 hi) { return ivec2(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y)); }
inline vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(vec3 x, vec3 y, vec3 a) { return x * (vec3(1.0f) - a) + y * a; }
inline vec2 fma(vec2 a, vec2 b, vec2 c) { return vec2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
inline vec3 fma(vec3 a, vec3 b, vec3 c) { return vec3(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y), fma(a.z, b.z, c.z)); }
inline vec3 ldexp(vec3 a, ivec3 e) { return vec3(ldexp(a.x, e.x), ldexp(a.y, e.y), ldexp(a.z, e.z)); }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(vec2 a) { return sqrt(dot(a, a)); }
inline float length(vec3 a) { return sqrt(dot(a, a)); }
inline float length(float a) { return abs(a); }
inline vec2 normalize(vec2 a) { return a * inversesqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a * inversesqrt(dot(a, a)); }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 refract(vec3 I, vec3 N, float eta) {
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return vec3(0.0f);
    return eta * I - (eta * d + sqrt(k)) * N;
}
inline bvec3 greaterThanEqual(vec3 a, vec3 b) { return bvec3(a.x >= b.x, a.y >= b.y, a.z >= b.z); }
inline bool all(bvec3 v) { return v.x && v.y && v.z; }

// --- matrices (column-major like glm) ---------------------------------------------------------------------------
struct mat2 {
    vec2 c[2];
    mat2() { c[0] = vec2(1, 0); c[1] = vec2(0, 1); }
    explicit mat2(float d) { c[0] = vec2(d, 0); c[1] = vec2(0, d); }
    mat2(vec2 a, vec2 b) { c[0] = a; c[1] = b; }
    mat2(float a, float b, float cc, float d) { c[0] = vec2(a, b); c[1] = vec2(cc, d); }
    vec2 &operator[](int i) { return c[i]; }
    const vec2 &operator[](int i) const { return c[i]; }
};
typedef mat2 mat2x2;
inline float determinant(const mat2 &m) { return m[0][0] * m[1][1] - m[1][0] * m[0][1]; }
inline mat2 operator*(const mat2 &m, float s) { return mat2(m[0] * s, m[1] * s); }
inline vec2 operator*(const mat2 &m, vec2 v) { return m[0] * v.x + m[1] * v.y; }
struct mat3x2 {
    vec2 c[3];
    mat3x2() {}
    mat3x2(vec2 a, vec2 b, vec2 d) { c[0] = a; c[1] = b; c[2] = d; }
    vec2 &operator[](int i) { return c[i]; }
    const vec2 &operator[](int i) const { return c[i]; }
};
inline vec2 operator*(const mat3x2 &m, vec3 v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z; }
struct mat4;
struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
    mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
    mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
        c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2);
    }
    explicit mat3(const mat4 &m);
    vec3 &operator[](int i) { return c[i]; }
    const vec3 &operator[](int i) const { return c[i]; }
};
inline vec3 operator*(const mat3 &m, vec3 v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z; }
inline mat3 transpose(const mat3 &m) {
    return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z));
}
struct mat4 {
    vec4 c[4];
    mat4() { c[0] = vec4(1, 0, 0, 0); c[1] = vec4(0, 1, 0, 0); c[2] = vec4(0, 0, 1, 0); c[3] = vec4(0, 0, 0, 1); }
    explicit mat4(float d) { c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d); }
    mat4(vec4 a, vec4 b, vec4 d, vec4 e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4 &m) { c[0] = vec3(m[0]); c[1] = vec3(m[1]); c[2] = vec3(m[2]); }
inline vec4 operator*(const mat4 &m, vec4 v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
typedef mat4 mat4x4;

} // namespace glm
