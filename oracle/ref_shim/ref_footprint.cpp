// oracle/ref_shim/ref_footprint.cpp -- TEST INFRASTRUCTURE.
// rendering/rt/footprint.glsl (the ray-footprint algebra behind the megakernel's texture level of detail, USE_MIPMAPPING) executed
// as C++: the functions dpdxy_to_footprint, transform_footprint, reflect_footprint and footprint_to_dpdxy (:10-61) are cut out of the
// file by the Makefile into oracle/_ref/gen/footprint.inc, unedited; this file supplies the GLSL matrix types they use.
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>

namespace reffp {
using namespace glm;
#include "rendering/language.hpp"
#include "rendering/util.glsl" // ortho_basis

// GLSL mat2x3: 2 columns of 3 rows
struct mat2x3 {
    vec3 c[2];
    mat2x3() {}
    mat2x3(vec3 a, vec3 b) { c[0] = a; c[1] = b; }
    vec3 &operator[](int i) { return c[i]; }
    const vec3 &operator[](int i) const { return c[i]; }
};
struct mat3x2t { vec3 r[2]; }; // transpose(mat2x3): 3 columns of 2 rows, kept as the two rows
static mat3x2t transpose(const mat2x3 &m) { mat3x2t t; t.r[0] = m[0]; t.r[1] = m[1]; return t; }
static mat2 transpose(const mat2 &m) { return mat2(vec2(m[0][0], m[1][0]), vec2(m[0][1], m[1][1])); }
static mat2 operator*(const mat3x2t &a, const mat2x3 &b) { // (2 x 3) * (3 x 2)
    return mat2(vec2(dot(a.r[0], b[0]), dot(a.r[1], b[0])), vec2(dot(a.r[0], b[1]), dot(a.r[1], b[1])));
}
static mat2x3 operator*(const mat3 &a, const mat2x3 &b) { return mat2x3(a * b[0], a * b[1]); }
static vec3 operator*(const mat2x3 &a, vec2 v) { return a[0] * v.x + a[1] * v.y; }
static mat2 operator*(const mat2 &a, const mat2 &b) { return mat2(a * b[0], a * b[1]); }
static mat3 outerProduct(vec3 c, vec3 r) { return mat3(c * r.x, c * r.y, c * r.z); }
static mat3 operator*(float s, const mat3 &m) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
static mat3 operator-(const mat3 &a, const mat3 &b) { return mat3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
#include "gen/footprint.inc"
} // namespace reffp

extern "C" void ref_footprint_op(int32_t op, const float *in, float *out) {
    using namespace reffp;
    if (op == 0) {
        const glm::mat2 F = dpdxy_to_footprint(glm::vec3(in[0], in[1], in[2]), glm::vec3(in[3], in[4], in[5]), glm::vec3(in[6], in[7], in[8]));
        out[0] = F[0][0]; out[1] = F[0][1]; out[2] = F[1][0]; out[3] = F[1][1];
    } else if (op == 1) {
        const glm::mat2 F = reflect_footprint(glm::vec3(in[0], in[1], in[2]), glm::vec3(in[3], in[4], in[5]), glm::mat2(in[6], in[7], in[8], in[9]));
        out[0] = F[0][0]; out[1] = F[0][1]; out[2] = F[1][0]; out[3] = F[1][1];
    } else {
        glm::vec3 dx, dy;
        footprint_to_dpdxy(dx, dy, glm::vec3(in[0], in[1], in[2]), glm::mat2(in[3], in[4], in[5], in[6]));
        out[0] = dx.x; out[1] = dx.y; out[2] = dx.z; out[3] = dy.x; out[4] = dy.y; out[5] = dy.z;
    }
}
