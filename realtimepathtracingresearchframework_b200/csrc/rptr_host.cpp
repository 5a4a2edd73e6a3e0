// rptr_host.cpp -- host-side scene ingestion: instance flattening, per-geometry parameter blocks, emitter
// collection + bin equalisation, and the SAH BVH build that stands in for the driver's BLAS/TLAS build.
#include "rptr_host.hpp"

#include <algorithm>
#include <chrono>
#include <functional>
#include <cmath>
#include <cstdio>
#include <array>
#include <atomic>
#include <stdexcept>
#include <thread>
#include <string>
#include <cstdlib>

namespace rp {

// ---- small host helpers ------------------------------------------------------------------------------------------------
static inline float3 xfm_point(const float *m, float3 v) {
    return f3(fmaf(m[0], v.x, fmaf(m[1], v.y, fmaf(m[2], v.z, m[3]))), fmaf(m[4], v.x, fmaf(m[5], v.y, fmaf(m[6], v.z, m[7]))),
              fmaf(m[8], v.x, fmaf(m[9], v.y, fmaf(m[10], v.z, m[11]))));
}

// rows of inverse(mat3(m)) = cross products of the columns / determinant
static void inverse_rows(const float *m, float *rows9) {
    float3 c0 = f3(m[0], m[4], m[8]), c1 = f3(m[1], m[5], m[9]), c2 = f3(m[2], m[6], m[10]);
    float3 r0 = cross(c1, c2), r1 = cross(c2, c0), r2 = cross(c0, c1);
    float det = dot(c0, r0);
    if (!(det != 0.0f)) throw std::runtime_error("instance transform is singular");
    float inv = 1.0f / det;
    r0 = r0 * inv; r1 = r1 * inv; r2 = r2 * inv;
    rows9[0] = r0.x; rows9[1] = r0.y; rows9[2] = r0.z;
    rows9[3] = r1.x; rows9[4] = r1.y; rows9[5] = r1.z;
    rows9[6] = r2.x; rows9[7] = r2.y; rows9[8] = r2.z;
}

void view_params(const rptr_camera_params &cam, int w, int h, float *du_, float *dv_, float *tl_) {
    float3 dir = ld3(cam.dir), up = ld3(cam.up);
    float py = 2.0f * tanf(0.5f * cam.fovy * 0.01745329251994329576923690768489f);
    float aspect = (float)w / (float)h;
    float px = py * aspect;
    float3 du = normalize(cross(dir, up)) * px;
    float3 dv = -normalize(cross(du, dir)) * py;
    float3 tl = dir - du * 0.5f - dv * 0.5f;
    du_[0] = du.x; du_[1] = du.y; du_[2] = du.z;
    dv_[0] = dv.x; dv_[1] = dv.y; dv_[2] = dv.z;
    tl_[0] = tl.x; tl_[1] = tl.y; tl_[2] = tl.z;
}

// ---- view_params.VP: vulkan/render_vulkan.cpp:2926-2930 ----------------------------------------------------------------------
// GLToVulkan * glm::infinitePerspective(radians(fovy), aspect, 0.5f) * inverse(mat4(mat4x3(cross(dir, up), up, -dir, cam_pos))).
// glm 0.9.9.8 is a build-time download of the reference (ext/CMakeLists.txt:18-21), not in its tree: its published algorithms
// (column-major storage; operator* accumulating A[k] * B[j][k] left to right; the cofactor inverse of
// detail::compute_inverse<4, 4>; infinitePerspectiveRH) are restated here with their operation order.
namespace {
inline float &at(float *m, int col, int row) { return m[4 * col + row]; }
inline float at(const float *m, int col, int row) { return m[4 * col + row]; }
void mat_mul(const float *a, const float *b, float *out) {
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            at(out, j, r) = ((at(a, 0, r) * at(b, j, 0) + at(a, 1, r) * at(b, j, 1)) + at(a, 2, r) * at(b, j, 2)) + at(a, 3, r) * at(b, j, 3);
}
void mat_inverse(const float *m, float *out) {
    // 2 x 2 sub-determinants, named by the row / column pairs they span
    const float c00 = at(m, 2, 2) * at(m, 3, 3) - at(m, 3, 2) * at(m, 2, 3), c02 = at(m, 1, 2) * at(m, 3, 3) - at(m, 3, 2) * at(m, 1, 3),
                c03 = at(m, 1, 2) * at(m, 2, 3) - at(m, 2, 2) * at(m, 1, 3), c04 = at(m, 2, 1) * at(m, 3, 3) - at(m, 3, 1) * at(m, 2, 3),
                c06 = at(m, 1, 1) * at(m, 3, 3) - at(m, 3, 1) * at(m, 1, 3), c07 = at(m, 1, 1) * at(m, 2, 3) - at(m, 2, 1) * at(m, 1, 3),
                c08 = at(m, 2, 1) * at(m, 3, 2) - at(m, 3, 1) * at(m, 2, 2), c10 = at(m, 1, 1) * at(m, 3, 2) - at(m, 3, 1) * at(m, 1, 2),
                c11 = at(m, 1, 1) * at(m, 2, 2) - at(m, 2, 1) * at(m, 1, 2), c12 = at(m, 2, 0) * at(m, 3, 3) - at(m, 3, 0) * at(m, 2, 3),
                c14 = at(m, 1, 0) * at(m, 3, 3) - at(m, 3, 0) * at(m, 1, 3), c15 = at(m, 1, 0) * at(m, 2, 3) - at(m, 2, 0) * at(m, 1, 3),
                c16 = at(m, 2, 0) * at(m, 3, 2) - at(m, 3, 0) * at(m, 2, 2), c18 = at(m, 1, 0) * at(m, 3, 2) - at(m, 3, 0) * at(m, 1, 2),
                c19 = at(m, 1, 0) * at(m, 2, 2) - at(m, 2, 0) * at(m, 1, 2), c20 = at(m, 2, 0) * at(m, 3, 1) - at(m, 3, 0) * at(m, 2, 1),
                c22 = at(m, 1, 0) * at(m, 3, 1) - at(m, 3, 0) * at(m, 1, 1), c23 = at(m, 1, 0) * at(m, 2, 1) - at(m, 2, 0) * at(m, 1, 1);
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3_[4] = {c12, c12, c14, c15}, f4_[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    float adj[16];
    for (int k = 0; k < 4; ++k) {
        const int col = k == 0 ? 1 : 0; // (m[1][r], m[0][r], m[0][r], m[0][r])
        const float v0 = at(m, col, 0), v1 = at(m, col, 1), v2 = at(m, col, 2), v3 = at(m, col, 3);
        const float sa = (k & 1) ? -1.0f : 1.0f, sb = -sa;
        at(adj, 0, k) = ((v1 * f0[k] - v2 * f1[k]) + v3 * f2[k]) * sa;
        at(adj, 1, k) = ((v0 * f0[k] - v2 * f3_[k]) + v3 * f4_[k]) * sb;
        at(adj, 2, k) = ((v0 * f1[k] - v1 * f3_[k]) + v3 * f5[k]) * sa;
        at(adj, 3, k) = ((v0 * f2[k] - v1 * f4_[k]) + v2 * f5[k]) * sb;
    }
    const float d0 = at(m, 0, 0) * at(adj, 0, 0), d1 = at(m, 0, 1) * at(adj, 1, 0), d2 = at(m, 0, 2) * at(adj, 2, 0), d3 = at(m, 0, 3) * at(adj, 3, 0);
    const float inv_det = 1.0f / ((d0 + d1) + (d2 + d3));
    for (int i = 0; i < 16; ++i) out[i] = adj[i] * inv_det;
}
} // namespace

void view_projection(const rptr_camera_params &cam, int w, int h, float *vp) {
    const float3 dir = ld3(cam.dir), up = ld3(cam.up), side = cross(dir, up);
    const float view[16] = {side.x, side.y, side.z, 0.0f, up.x, up.y, up.z, 0.0f, -dir.x, -dir.y, -dir.z, 0.0f, cam.pos[0], cam.pos[1], cam.pos[2], 1.0f};
    float inv_view[16];
    mat_inverse(view, inv_view);
    // infinitePerspectiveRH(fovy, aspect, zNear)
    const float fovy = cam.fovy * 0.01745329251994329576923690768489f, aspect = (float)w / (float)h, z_near = 0.5f;
    const float range = tanf(fovy / 2.0f) * z_near;
    const float left = -range * aspect, right = range * aspect, bottom = -range, top = range;
    float proj[16] = {0.0f};
    at(proj, 0, 0) = (2.0f * z_near) / (right - left);
    at(proj, 1, 1) = (2.0f * z_near) / (top - bottom);
    at(proj, 2, 2) = -1.0f;
    at(proj, 2, 3) = -1.0f;
    at(proj, 3, 2) = -2.0f * z_near;
    float to_vulkan[16] = {0.0f};
    at(to_vulkan, 0, 0) = 1.0f; at(to_vulkan, 1, 1) = -1.0f; at(to_vulkan, 2, 2) = 0.5f; at(to_vulkan, 3, 3) = 1.0f; at(to_vulkan, 3, 2) = 0.5f;
    float clip[16];
    mat_mul(to_vulkan, proj, clip);
    mat_mul(clip, inv_view, vp);
}

// ---- emitters: librender/lights.cpp ---------------------------------------------------------------------------------------
namespace {

struct Emitter { float3 v0, v1, v2, radiance; };

inline float halton2(unsigned index) { // util/compute_util.h:19-33
    index = (index << 16) | (index >> 16);
    index = ((index & 0x00ff00ffu) << 8) | ((index & 0xff00ff00u) >> 8);
    index = ((index & 0x0f0f0f0fu) << 4) | ((index & 0xf0f0f0f0u) >> 4);
    index = ((index & 0x33333333u) << 2) | ((index & 0xccccccccu) >> 2);
    index = ((index & 0x55555555u) << 1) | ((index & 0xaaaaaaaau) >> 1);
    return u2f(0x3f800000u | (index >> 9)) - 1.0f;
}

// lights.cpp:169-203; the host variant uses libm atan (lights.cpp:122-129) and divides by M_2_PI (sic, :195)
std::vector<float> estimate_normalized_radiance(const std::vector<Emitter> &em, float min_dist) {
    std::vector<float> rad(em.size());
    for (size_t i = 0; i < em.size(); ++i) {
        const Emitter &l = em[i];
        float3 n = normalize(cross(l.v1 - l.v0, l.v2 - l.v0));
        if (!(fabsf(length(n) - 1.0f) < 0.05f)) { rad[i] = 0.0f; continue; }
        float3 c = (l.v0 + l.v1 + l.v2) / 3.0f;
        float3 o = n * min_dist;
        float3 prm;
        float tangent = half_tri_solid_angle_tan(normalize(l.v0 - c - o), normalize(l.v1 - c - o), normalize(l.v2 - c - o), prm);
        float sa = 2.0f * (atanf(tangent) + ((tangent < 0.0f) ? (float)M_PI : 0.0f));
        float lum = 0.2126f * l.radiance.x + 0.7152f * l.radiance.y + 0.0722f * l.radiance.z; // util/util.cpp:293-296
        rad[i] = (float)((double)lum * ((double)sa / M_2_PI));
    }
    return rad;
}

struct Slot { float radiance; int source; int splits; };

void halton_shuffle(std::vector<Slot> &slots) { // lights.cpp:246-263
    std::vector<Slot> out(slots.size());
    const int count = (int)slots.size();
    for (int i = 0; i < count; ++i) {
        int src = (int)(unsigned)(halton2((unsigned)i) * (float)(unsigned)count);
        for (;;) {
            if (src >= count) src = 0;
            if (slots[src].source == ~0) ++src;
            else break;
        }
        out[i] = slots[src];
        slots[src].source = ~0;
    }
    slots.swap(out);
}

float bin_equality(const std::vector<Slot> &slots, int bin_size) { // lights.cpp:266-279
    float lo = 2.0e32f, hi = 0.0f;
    for (int i = 0, n = (int)slots.size(); i < n;) {
        float total = 0.0f;
        for (int j = 0; j < bin_size && i < n; ++i, ++j) total += slots[i].radiance;
        lo = std::min(total, lo);
        hi = std::max(total, hi);
    }
    return std::min(lo / hi, 1.0f);
}

// lights.cpp:220-349
void equalize_emitter_bins(std::vector<Emitter> &emitters, std::vector<float> &radiances, int bin_size) {
    if (bin_size <= 1 || radiances.empty()) return;
    const int n0 = (int)radiances.size();
    const int bins0 = (n0 + (bin_size - 1)) / bin_size;
    float avg = 0.0f;
    for (float r : radiances) avg += r;
    avg /= (float)radiances.size();
    std::vector<Slot> slots;
    slots.reserve(2 * radiances.size());
    for (int i = 0; i < n0; ++i) {
        int clones = (int)std::max((unsigned)std::min(radiances[i] / avg, (float)bins0), 1u);
        for (int j = 0; j < clones; ++j) slots.push_back(Slot{radiances[i] / (float)clones, i, clones});
    }
    halton_shuffle(slots);
    float equality = bin_equality(slots, bin_size);
    std::vector<Slot> cdf;
    for (int retry = 0; equality < 0.6f && retry < 2; ++retry) {
        cdf.resize(slots.size());
        Slot acc = slots[0];
        cdf[0] = acc;
        for (size_t i = 1; i < slots.size(); ++i) {
            acc = Slot{acc.radiance + slots[i].radiance, slots[i].source, 1};
            cdf[i] = acc;
        }
        cdf.front().splits = 1;
        const float sum = cdf.back().radiance;
        for (Slot &c : cdf) c.radiance /= sum;
        const int prev = (int)slots.size();
        const int padded = ((prev + (bin_size - 1)) / bin_size + 1) * bin_size;
        unsigned h = 0;
        while ((int)slots.size() < padded) {
            float u = halton2(h++);
            auto it = std::upper_bound(cdf.begin(), cdf.end(), u, [](float bound, const Slot &s) { return bound < s.radiance; });
            if (it == cdf.end()) it = cdf.end() - 1;
            ++it->splits;
            slots.push_back(Slot{it->radiance, (int)(it - cdf.begin()), 0});
        }
        for (int i = prev; i < padded; ++i) {
            Slot &clone = slots[i];
            Slot &orig = slots[clone.source];
            int &counter = cdf[clone.source].splits;
            if (counter > 1) {
                orig.radiance /= (float)counter;
                orig.splits *= counter;
                counter = 1;
            }
            clone.radiance = orig.radiance;
            clone.source = orig.source;
            clone.splits = orig.splits;
        }
        halton_shuffle(slots);
        equality = bin_equality(slots, bin_size);
    }
    std::vector<Emitter> out(slots.size());
    radiances.resize(slots.size());
    for (size_t i = 0; i < slots.size(); ++i) {
        radiances[i] = slots[i].radiance;
        out[i] = emitters[slots[i].source];
        out[i].radiance = out[i].radiance / (float)slots[i].splits;
    }
    emitters.swap(out);
}

} // namespace

// ---- scene flattening ---------------------------------------------------------------------------------------------------
// Raster-TAA screen jitter of update_view_parameters (vulkan/render_vulkan.cpp:2917-2926): halton_23[(frame_offset +
// frame_id) % RASTER_TAA_NUM_SAMPLES] * 2 / dims - 1 / dims.  The table of librender/halton.h holds the 2-3 Halton points
// of index k + 1 as 6-decimal literals, which is how they are rebuilt here (RASTER_TAA_NUM_SAMPLES = 16, CMakeLists.txt:30).
void halton_23(int k, float *out) {
    const int bases[2] = {2, 3};
    for (int c = 0; c < 2; ++c) {
        double r = 0.0, f = 1.0;
        for (int i = k + 1; i > 0; i /= bases[c]) {
            f /= bases[c];
            r += f * (i % bases[c]);
        }
        // the literal "0.dddddd" the table spells out (printf rounding: exact ties to even), parsed to the nearest float
        char lit[32];
        std::snprintf(lit, sizeof(lit), "%.6f", r);
        out[c] = std::strtof(lit, nullptr);
    }
}
void screen_jitter(uint32_t frame_offset, uint32_t frame_id, int w, int h, float *out) {
    float hp[2];
    halton_23((int)((frame_offset + frame_id) % 16u), hp);
    out[0] = hp[0] * 2.0f / (float)w - 1.0f / (float)w;
    out[1] = hp[1] * 2.0f / (float)h - 1.0f / (float)h;
}

// ---- 1x1-texel mode: material parameters that carry texture handles are resolved on the host -----------------
// The reference samples textures at the hit's uv (rendering/rt/material_textures.glsl:37-63); a 1 x 1 texture returns its only
// texel for every uv and every LOD, so the lookup can be done once per material here and the kernels keep reading
// constants.  UNORM8 -> v / 255 in float; colour channels of an sRGB image go through the sRGB transfer function evaluated in
// double and rounded once (the texture unit's own table is not specified bit for bit; this is our statement of it).
namespace {
float srgb8_to_linear(int v) {
    const double c = (double)v / 255.0;
    return (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
}
struct Texel { float c[4]; int a8; };
const rptr_texture_desc &texture_of(const rptr_scene_desc &d, uint32_t handle_or_id) {
    const uint32_t id = RPTR_GET_TEXTURE_ID(handle_or_id);
    if (!d.textures || id >= (uint32_t)d.n_textures) throw std::runtime_error("material refers to texture " + std::to_string(id) + " which the scene does not have");
    const rptr_texture_desc &t = d.textures[id];
    if (t.width < 1 || t.height < 1 || t.width > 32768 || t.height > 32768) throw std::runtime_error("texture " + std::to_string(id) + " has an invalid size");
    if (!t.texels) throw std::runtime_error("texture " + std::to_string(id) + " has no texels");
    if (t.bc_format == 0 && (t.channels < 1 || t.channels > 4)) throw std::runtime_error("texture " + std::to_string(id) + " has a bad channel count");
    if (t.bc_format != 0 && t.bc_format != 1 && t.bc_format != -1 && t.bc_format != 3 && t.bc_format != 5)
        throw std::runtime_error("texture " + std::to_string(id) + ": block compression format " + std::to_string(t.bc_format) + " is not supported (BC1, BC3, BC5 UNORM are)");
    if (t.mip_levels < 0 || t.mip_levels > 16) throw std::runtime_error("texture " + std::to_string(id) + " has an invalid number of mip levels");
    return t;
}
bool is_one_texel(const rptr_texture_desc &t) { return t.width == 1 && t.height == 1; }
Texel fetch_texel(const rptr_scene_desc &d, uint32_t handle) { // the only texel of a 1 x 1 texture
    const rptr_texture_desc &t = texture_of(d, handle);
    HostTexture ht;
    decode_texture(t, ht);
    Texel x;
    for (int k = 0; k < 3; ++k) x.c[k] = t.color_space == RPTR_COLOR_SPACE_SRGB ? srgb8_to_linear(ht.rgba[k]) : (float)ht.rgba[k] / 255.0f;
    x.a8 = ht.rgba[3];
    x.c[3] = alpha8_to_float(x.a8);
    return x;
}
bool has_alpha_channel(const rptr_texture_desc &t) { // can a texel of this image have alpha != 1?
    return t.bc_format == 0 ? t.channels == 4 : (t.bc_format == -1 || t.bc_format == 3);
}
bool is_handle(float v) { return (f2u(v) & RPTR_TEXTURED_PARAM_MASK) != 0; }
// textured_scalar_param (material_textures.glsl:50-63): folded when the texture has one texel, left as a handle otherwise
float resolve_scalar(const rptr_scene_desc &d, float v, bool &kept_handle) {
    if (!is_handle(v)) return v;
    if (!is_one_texel(texture_of(d, f2u(v)))) { kept_handle = true; return v; }
    return fetch_texel(d, f2u(v)).c[RPTR_GET_TEXTURE_CHANNEL(f2u(v))];
}
} // namespace

// ---- texture ingestion: every mip level to RGBA8 ----------------------------------------------------------------------------
// Block formats as the Khronos Data Format Specification defines them (S3TC / RGTC sections); interpolated values are the exact
// rationals of the specification rounded to the nearest 8-bit code, which is where hardware decoders are allowed to differ.
namespace {
void expand565(uint32_t c, int *rgb) {
    const int r = (c >> 11) & 31, g = (c >> 5) & 63, b = c & 31;
    rgb[0] = (r << 3) | (r >> 2); rgb[1] = (g << 2) | (g >> 4); rgb[2] = (b << 3) | (b >> 2);
}
// one BC1 colour block (8 bytes) into a 4 x 4 RGBA tile; mode: 0 = BC1 RGB (no transparency), 1 = BC1 RGBA (index 3 of the three-colour
// mode is transparent black), 2 = the colour half of BC3 (always four colours, alpha untouched)
void decode_bc1_block(const uint8_t *b, int mode, uint8_t tile[16][4]) {
    const uint32_t c0 = b[0] | (b[1] << 8), c1 = b[2] | (b[3] << 8);
    int pal[4][4];
    expand565(c0, pal[0]); expand565(c1, pal[1]);
    pal[0][3] = pal[1][3] = pal[2][3] = pal[3][3] = 255;
    if (c0 > c1 || mode == 2) {
        for (int k = 0; k < 3; ++k) { pal[2][k] = (2 * pal[0][k] + pal[1][k] + 1) / 3; pal[3][k] = (pal[0][k] + 2 * pal[1][k] + 1) / 3; }
    } else {
        for (int k = 0; k < 3; ++k) { pal[2][k] = (pal[0][k] + pal[1][k] + 1) / 2; pal[3][k] = 0; }
        if (mode == 1) pal[3][3] = 0;
    }
    const uint32_t idx = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
    for (int i = 0; i < 16; ++i) {
        const int *p = pal[(idx >> (2 * i)) & 3];
        for (int k = 0; k < 3; ++k) tile[i][k] = (uint8_t)p[k];
        if (mode != 2) tile[i][3] = (uint8_t)p[3];
    }
}
// one BC4 UNORM block (8 bytes: two endpoints, 16 three-bit indices) into channel `ch` of the tile
void decode_bc4_block(const uint8_t *b, int ch, uint8_t tile[16][4]) {
    int pal[8];
    pal[0] = b[0]; pal[1] = b[1];
    if (pal[0] > pal[1]) for (int i = 1; i <= 6; ++i) pal[1 + i] = ((7 - i) * pal[0] + i * pal[1] + 3) / 7;
    else {
        for (int i = 1; i <= 4; ++i) pal[1 + i] = ((5 - i) * pal[0] + i * pal[1] + 2) / 5;
        pal[6] = 0; pal[7] = 255;
    }
    uint64_t idx = 0;
    for (int k = 0; k < 6; ++k) idx |= (uint64_t)b[2 + k] << (8 * k);
    for (int i = 0; i < 16; ++i) tile[i][ch] = (uint8_t)pal[(idx >> (3 * i)) & 7];
}
} // namespace

void decode_texture(const rptr_texture_desc &td, HostTexture &out) {
    if (td.bc_format != 0 && td.bc_format != 1 && td.bc_format != -1 && td.bc_format != 3 && td.bc_format != 5)
        throw std::runtime_error("block compression format " + std::to_string(td.bc_format) + " is not supported (BC1, BC3, BC5 UNORM are)");
    out.width = td.width; out.height = td.height; out.srgb = td.color_space == RPTR_COLOR_SPACE_SRGB;
    out.levels = td.mip_levels > 0 ? td.mip_levels : 1;
    out.rgba.clear();
    const uint8_t *src = td.texels;
    int w = td.width, h = td.height;
    for (int l = 0; l < out.levels; ++l) {
        const size_t base = out.rgba.size();
        out.rgba.resize(base + 4 * (size_t)w * h);
        uint8_t *dst = out.rgba.data() + base;
        if (td.bc_format == 0) {
            for (size_t i = 0; i < (size_t)w * h; ++i)
                for (int k = 0; k < 4; ++k) dst[4 * i + k] = k < td.channels ? src[i * td.channels + k] : (k == 3 ? 255 : 0);
            src += (size_t)w * h * td.channels;
        } else { // levels are padded to whole 4 x 4 blocks (vulkan/resource_utils.cpp:85-99)
            const int bw = (w + 3) / 4, bh = (h + 3) / 4;
            const int block_bytes = (td.bc_format == 1 || td.bc_format == -1) ? 8 : 16;
            for (int by = 0; by < bh; ++by)
                for (int bx = 0; bx < bw; ++bx) {
                    const uint8_t *b = src + ((size_t)by * bw + bx) * block_bytes;
                    uint8_t tile[16][4];
                    for (int i = 0; i < 16; ++i) { tile[i][0] = tile[i][1] = tile[i][2] = 0; tile[i][3] = 255; }
                    if (td.bc_format == 1) decode_bc1_block(b, 0, tile);
                    else if (td.bc_format == -1) decode_bc1_block(b, 1, tile);
                    else if (td.bc_format == 3) { decode_bc4_block(b, 3, tile); decode_bc1_block(b + 8, 2, tile); }
                    else { decode_bc4_block(b, 0, tile); decode_bc4_block(b + 8, 1, tile); } // BC5 UNORM: red, green; blue 0, alpha 1
                    for (int ty = 0; ty < 4; ++ty)
                        for (int tx = 0; tx < 4; ++tx) {
                            const int x = 4 * bx + tx, y = 4 * by + ty;
                            if (x < w && y < h) memcpy(dst + 4 * ((size_t)y * w + x), tile[4 * ty + tx], 4);
                        }
                }
            src += (size_t)bw * bh * block_bytes;
        }
        if (w > 1) w /= 2;
        if (h > 1) h /= 2;
    }
}

// unpack_material's texture reads (material_textures.glsl:95-135, non-unrolled standard-texture semantics of
// rendering/rt/materials.glsl:42-49): parameters that refer to 1 x 1 textures are folded into the BaseMaterial (a 1 x 1 texture
// returns its only texel for every uv and LOD), alpha8 = the texel get_material_alpha() would read; parameters that refer to
// larger textures keep their handles and are sampled at the hit's uv on the device (rptr_shading.cuh: sample_texture).
static void resolve_materials(const rptr_scene_desc &d, HostScene &s) {
    s.materials.assign(d.materials, d.materials + d.n_materials);
    s.material_alpha8.assign(d.n_materials, RPTR_TRI_OPAQUE);
    s.material_alpha_textured.assign(d.n_materials, 0);
    s.normal_texels.assign(4 * (size_t)d.n_materials, 0.0f);
    for (int v = 0; v < 256; ++v) s.srgb_lut[v] = srgb8_to_linear(v);
    s.textures.assign(d.textures ? (size_t)d.n_textures : 0, HostTexture());
    for (size_t t = 0; t < s.textures.size(); ++t) {
        const rptr_texture_desc &td = texture_of(d, (uint32_t)t);
        HostTexture &ht = s.textures[t];
        ht.width = td.width; ht.height = td.height; ht.srgb = td.color_space == RPTR_COLOR_SPACE_SRGB;
        if (is_one_texel(td)) continue;
        decode_texture(td, ht);
    }
    for (int i = 0; i < d.n_materials; ++i) {
        rptr_base_material &m = s.materials[i];
        bool kept = false;
        if (m.normal_map != -1) { // the normal-map texel as textureLod(...).rgb returns it (pt_megakernel.glsl:647)
            if (is_one_texel(texture_of(d, (uint32_t)m.normal_map))) {
                const Texel x = fetch_texel(d, (uint32_t)m.normal_map);
                s.normal_texels[4 * i + 0] = x.c[0]; s.normal_texels[4 * i + 1] = x.c[1]; s.normal_texels[4 * i + 2] = x.c[2];
            } else {
                s.normal_texels[4 * i + 3] = 1.0f; // sampled at the hit's uv
                kept = true;
            }
            s.any_normal_map = true;
        }
        if (is_handle(m.base_color[0])) {
            if (m.emission_intensity != 0.0f) throw std::runtime_error("emissive materials with a textured base colour are not supported by this backend yet");
            if (is_one_texel(texture_of(d, f2u(m.base_color[0])))) {
                const Texel x = fetch_texel(d, f2u(m.base_color[0]));
                const float alpha = x.c[3];
                for (int k = 0; k < 3; ++k) m.base_color[k] = alpha > 0.001f ? x.c[k] / alpha : x.c[k]; // PREMULTIPLIED_BASE_COLOR_ALPHA, :101-104
                s.material_alpha8[i] = x.a8;
            } else {
                kept = true;
                // three- or fewer-channel images have alpha 1 everywhere: nothing to test during traversal
                if (has_alpha_channel(texture_of(d, f2u(m.base_color[0])))) s.material_alpha_textured[i] = 1;
            }
        }
        m.specular = resolve_scalar(d, m.specular, kept);
        m.roughness = resolve_scalar(d, m.roughness, kept);
        m.metallic = resolve_scalar(d, m.metallic, kept);
        m.ior = resolve_scalar(d, m.ior, kept);
        m.specular_transmission = resolve_scalar(d, m.specular_transmission, kept);
        m.clearcoat_gloss = resolve_scalar(d, m.clearcoat_gloss, kept);
        if (m.flags & RPTR_BASE_MATERIAL_NOALPHA) { // never alpha-tested (pt_megakernel.glsl:202)
            s.material_alpha8[i] = RPTR_TRI_OPAQUE;
            s.material_alpha_textured[i] = 0;
        }
        if (kept) s.any_textured = true;
    }
}

void build_host_scene(const rptr_scene_desc &d, const rptr_light_sampling_config &ls, HostScene &s, bool with_bvh) {
    s = HostScene();
    if (d.n_materials <= 0 || !d.materials) throw std::runtime_error("scene has no materials");
    resolve_materials(d, s);
    s.qverts.resize(d.n_geometries);
    s.qnuv.resize(d.n_geometries);
    for (int g = 0; g < d.n_geometries; ++g) {
        const rptr_geometry_desc &gd = d.geometries[g];
        if (gd.n_tris < 0 || (gd.n_tris > 0 && !gd.qverts)) throw std::runtime_error("geometry without vertices");
        s.qverts[g].assign(gd.qverts, gd.qverts + 3 * (size_t)gd.n_tris);
        if (gd.qnormal_uv && (gd.has_normals || gd.has_uvs)) s.qnuv[g].assign(gd.qnormal_uv, gd.qnormal_uv + 3 * (size_t)gd.n_tris);
    }
    s.tri_mat.resize(d.n_pmeshes);
    for (int p = 0; p < d.n_pmeshes; ++p) {
        const rptr_pmesh_desc &pm = d.pmeshes[p];
        if (pm.mesh_id < 0 || pm.mesh_id >= d.n_meshes) throw std::runtime_error("parameterized mesh refers to a missing mesh");
        if (pm.tri_material_ids) s.tri_mat[p].assign(pm.tri_material_ids, pm.tri_material_ids + pm.n_tri_material_ids);
    }
    size_t total = 0;
    for (int i = 0; i < d.n_instances; ++i) {
        const rptr_instance_desc &inst = d.instances[i];
        if (inst.pmesh_id < 0 || inst.pmesh_id >= d.n_pmeshes) throw std::runtime_error("instance refers to a missing parameterized mesh");
        const rptr_mesh_desc &mesh = d.meshes[d.pmeshes[inst.pmesh_id].mesh_id];
        for (int j = 0; j < mesh.n_geometries; ++j) total += (size_t)d.geometries[mesh.first_geometry + j].n_tris;
    }
    if (total > 0x1ffffff0u) throw std::runtime_error("too many triangles after instancing (limit 2^29)");
    {
        size_t n_gi = 0;
        for (int i = 0; i < d.n_instances; ++i) n_gi += (size_t)d.meshes[d.pmeshes[d.instances[i].pmesh_id].mesh_id].n_geometries;
        if (n_gi >= (1u << 23)) throw std::runtime_error("too many (instance, geometry) pairs (limit 2^23)");
    }
    s.tris.reserve(total);

    std::vector<Emitter> emitters;
    std::vector<char> nonemissive(d.n_pmeshes, 0);
    for (int i = 0; i < d.n_instances; ++i) {
        const rptr_instance_desc &inst = d.instances[i];
        const rptr_pmesh_desc &pm = d.pmeshes[inst.pmesh_id];
        const rptr_mesh_desc &mesh = d.meshes[pm.mesh_id];
        const bool per_tri = !s.tri_mat[inst.pmesh_id].empty();
        int64_t prim_offset = 0;
        std::vector<Emitter> found;
        for (int j = 0; j < mesh.n_geometries; ++j) {
            const int gidx = mesh.first_geometry + j;
            const rptr_geometry_desc &gd = d.geometries[gidx];
            HostGeomInst hg;
            memset(&hg, 0, sizeof(hg));
            hg.geometry = gidx;
            hg.pmesh = inst.pmesh_id;
            hg.prim_offset = prim_offset;
            GeomInst &g = hg.g;
            for (int k = 0; k < 3; ++k) { g.scale[k] = gd.quantized_scaling[k]; g.offset[k] = gd.quantized_offset[k]; }
            g.has_normals = gd.has_normals && !s.qnuv[gidx].empty();
            g.has_uvs = gd.has_uvs && !s.qnuv[gidx].empty();
            const int mat_off = pm.n_material_offsets ? pm.material_offsets[j] : 0;
            g.flags = RPTR_GEOMETRY_FLAGS_IMPLICIT_INDICES;
            const uint8_t *tm = per_tri ? s.tri_mat[inst.pmesh_id].data() + prim_offset : nullptr;
            bool no_alpha;
            if (per_tri) { // render_vulkan.cpp:2812-2818
                if (prim_offset + gd.n_tris > (int64_t)s.tri_mat[inst.pmesh_id].size()) throw std::runtime_error("per-triangle material ids too short");
                g.material_id = -1 - mat_off;
                g.flags |= RPTR_GEOMETRY_FLAGS_EXTENDED_SHADER;
                no_alpha = true;
                for (int t = 0; t < gd.n_tris; ++t) {
                    int mid = mat_off + tm[t];
                    if (mid >= d.n_materials) throw std::runtime_error("per-triangle material id out of range");
                    if (!(s.materials[mid].flags & RPTR_BASE_MATERIAL_NOALPHA)) no_alpha = false;
                }
            } else {
                if (mat_off < 0 || mat_off >= d.n_materials) throw std::runtime_error("material offset out of range");
                g.material_id = mat_off;
                no_alpha = (s.materials[mat_off].flags & RPTR_BASE_MATERIAL_NOALPHA) != 0;
                if (s.materials[mat_off].flags & RPTR_BASE_MATERIAL_EXTENDED) g.flags |= RPTR_GEOMETRY_FLAGS_EXTENDED_SHADER;
                if (!(s.materials[mat_off].flags & RPTR_BASE_MATERIAL_ONESIDED)) g.flags |= RPTR_GEOMETRY_FLAGS_THIN;
            }
            if (no_alpha) g.flags |= RPTR_GEOMETRY_FLAGS_NOALPHA;
            else s.any_non_opaque = true;
            g.instance = i;
            memcpy(hg.o2w, inst.transform, sizeof(hg.o2w));
            inverse_rows(hg.o2w, g.w2o);
            const int gi = (int)s.ginst.size();
            const uint64_t *qv = s.qverts[gidx].data();
            for (int t = 0; t < gd.n_tris; ++t) {
                float3 a = xfm_point(hg.o2w, dequantize_position(qv[3 * (size_t)t + 0], g.scale, g.offset));
                float3 b = xfm_point(hg.o2w, dequantize_position(qv[3 * (size_t)t + 1], g.scale, g.offset));
                float3 c = xfm_point(hg.o2w, dequantize_position(qv[3 * (size_t)t + 2], g.scale, g.offset));
                float3 e1 = b - a, e2 = c - a;
                Tri tr;
                tr.v0x = a.x; tr.v0y = a.y; tr.v0z = a.z;
                tr.e1x = e1.x; tr.e1y = e1.y; tr.e1z = e1.z;
                tr.e2x = e2.x; tr.e2y = e2.y; tr.e2z = e2.z;
                tr.id = (int32_t)s.tris.size();
                const int tri_material = per_tri ? mat_off + tm[t] : mat_off;
                const int32_t a8 = no_alpha ? RPTR_TRI_OPAQUE : s.material_alpha8[tri_material];
                const bool alpha_tex = !no_alpha && s.material_alpha_textured[tri_material] != 0;
                if (a8 != RPTR_TRI_OPAQUE || alpha_tex) s.any_alpha_tested = true;
                if (alpha_tex && !(gd.has_uvs && !s.qnuv[gidx].empty())) throw std::runtime_error("a material with an alpha texture is used on a geometry without uvs");
                tr.gi_alpha = pack_gi_alpha(gi, a8, alpha_tex);
                tr.prim = t;
                s.tris.push_back(tr);
                if (!nonemissive[inst.pmesh_id]) { // collect_emitters, lights.cpp:33-73
                    const rptr_base_material &m = s.materials[per_tri ? mat_off + tm[t] : mat_off];
                    if (m.emission_intensity > 0.0f) found.push_back(Emitter{a, b, c, ld3(m.base_color) * m.emission_intensity});
                }
            }
            s.ginst.push_back(hg);
            prim_offset += gd.n_tris;
        }
        if (!nonemissive[inst.pmesh_id]) { // lights.cpp:17-30: each instance's emitters are prepended
            if (!found.empty()) emitters.insert(emitters.begin(), found.begin(), found.end());
            else nonemissive[inst.pmesh_id] = 1;
        }
    }
    if (d.binned_lights && d.n_binned_lights > 0) {
        s.lights.assign(d.binned_lights, d.binned_lights + d.n_binned_lights);
    } else if (!emitters.empty()) { // update_light_sampling, lights.cpp:75-90
        std::vector<float> rad = estimate_normalized_radiance(emitters, ls.min_perceived_receiver_dist);
        if (ls.min_radiance > 0.0f) { // trim_dim_emitters, lights.cpp:205-218
            size_t n = 0;
            for (size_t i = 0; i < emitters.size(); ++i)
                if (rad[i] >= ls.min_radiance) { emitters[n] = emitters[i]; rad[n] = rad[i]; ++n; }
            emitters.resize(n);
            rad.resize(n);
        }
        equalize_emitter_bins(emitters, rad, ls.bin_size);
        s.lights.reserve(emitters.size());
        for (const Emitter &e : emitters) {
            rptr_tri_light_data t;
            t.v0[0] = e.v0.x; t.v0[1] = e.v0.y; t.v0[2] = e.v0.z;
            t.v1[0] = e.v1.x; t.v1[1] = e.v1.y; t.v1[2] = e.v1.z;
            t.v2[0] = e.v2.x; t.v2[1] = e.v2.y; t.v2[2] = e.v2.z;
            t.radiance[0] = e.radiance.x; t.radiance[1] = e.radiance.y; t.radiance[2] = e.radiance.z;
            s.lights.push_back(t);
        }
    }
    if (with_bvh) build_bvh(s);
}

// ---- BVH build: binned-SAH binary tree, collapsed to the 4-wide breadth-first layout of rptr_bvh.cuh ---------------------
namespace {

struct Prim { float lo[3], hi[3], c[3]; int32_t id; };
struct Node2 { // temporary binary node
    float lo[3], hi[3];
    int32_t left, right; // inner: child indices; leaf: left = ~first_prim_slot, right = count
};
struct Builder {
    std::vector<Prim> prims;
    std::vector<Node2> nodes;
    // sub-trees of at most `cutoff` primitives are left as placeholders by the serial top-down pass and built by worker
    // threads afterwards (the primitive ranges are disjoint, the result does not depend on the thread count)
    struct Job { int32_t node; int lo, hi, depth; };
    int sah_depth_limit = 32;
    // cost of one traversal step in triangle tests (tuning: RPTR_SAH_TRAV_COST in the environment, profiles/r01_trace_sweep.md)
    float trav_cost = std::getenv("RPTR_SAH_TRAV_COST") ? (float)std::atof(std::getenv("RPTR_SAH_TRAV_COST")) : 0.6f;
    static constexpr int NB = 16;
    static constexpr int MAX_LEAF = 4;
    static constexpr int SWEEP_MAX = 8;
    // nodes of at least PAR_MIN primitives (the top levels) are processed by all threads: bounds and bins are reduced from
    // per-thread partials (min / max / counts: the result does not depend on the split), and the primitives are partitioned
    // STABLY -- by every path, whatever the thread count -- so that the tree is the same for any number of threads
    static constexpr int PAR_MIN = 131072;
    int par_threads = 1;
    std::vector<Prim> scratch;
    template <class F> void pfor(int T, int lo, int n, F &&body) { // body(t, i0, i1) over T contiguous slices of [lo, lo + n)
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back([&, t]() { body(t, lo + (int)((int64_t)n * t / T), lo + (int)((int64_t)n * (t + 1) / T)); });
        body(0, lo, lo + (int)((int64_t)n / T));
        for (std::thread &x : th) x.join();
    }
    struct Bins { float bl[3][NB][3], bh[3][NB][3]; int cnt[3][NB]; };
    static void bins_clear(Bins &b) {
        for (int ax = 0; ax < 3; ++ax)
            for (int i = 0; i < NB; ++i) {
                b.cnt[ax][i] = 0;
                for (int k = 0; k < 3; ++k) { b.bl[ax][i][k] = 1e30f; b.bh[ax][i][k] = -1e30f; }
            }
    }
    void bins_add(Bins &b, int i0, int i1, const float *cmin, const float *sc) const { // sc[ax] == 0: axis not binned
        for (int i = i0; i < i1; ++i)
            for (int ax = 0; ax < 3; ++ax) {
                if (!(sc[ax] > 0.0f)) continue;
                const int q = std::min(NB - 1, std::max(0, (int)((prims[i].c[ax] - cmin[ax]) * sc[ax])));
                b.cnt[ax][q]++;
                for (int k = 0; k < 3; ++k) {
                    b.bl[ax][q][k] = fminf(b.bl[ax][q][k], prims[i].lo[k]);
                    b.bh[ax][q][k] = fmaxf(b.bh[ax][q][k], prims[i].hi[k]);
                }
            }
    }
    static float half_area(const float *lo, const float *hi) {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    int32_t build(std::vector<Node2> &nodes, int lo, int hi, int depth, std::vector<Job> *jobs, int cutoff) {
        const int32_t idx = (int32_t)nodes.size();
        nodes.push_back(Node2());
        if (jobs && depth > 0 && hi - lo <= cutoff) {
            jobs->push_back(Job{idx, lo, hi, depth});
            return idx;
        }
        const int n = hi - lo;
        const int T = (jobs && n >= PAR_MIN) ? par_threads : 1; // threads working on THIS node
        float blo[3], bhi[3], cmin[3], cmax[3];
        {
            std::vector<std::array<float, 12>> part((size_t)T);
            pfor(T, lo, n, [&](int t, int i0, int i1) {
                std::array<float, 12> &b = part[(size_t)t];
                for (int k = 0; k < 3; ++k) { b[k] = 1e30f; b[3 + k] = -1e30f; b[6 + k] = 1e30f; b[9 + k] = -1e30f; }
                for (int i = i0; i < i1; ++i)
                    for (int k = 0; k < 3; ++k) {
                        b[k] = fminf(b[k], prims[i].lo[k]);
                        b[3 + k] = fmaxf(b[3 + k], prims[i].hi[k]);
                        b[6 + k] = fminf(b[6 + k], prims[i].c[k]);
                        b[9 + k] = fmaxf(b[9 + k], prims[i].c[k]);
                    }
            });
            for (int k = 0; k < 3; ++k) { blo[k] = 1e30f; bhi[k] = -1e30f; cmin[k] = 1e30f; cmax[k] = -1e30f; }
            for (const std::array<float, 12> &b : part)
                for (int k = 0; k < 3; ++k) {
                    blo[k] = fminf(blo[k], b[k]); bhi[k] = fmaxf(bhi[k], b[3 + k]);
                    cmin[k] = fminf(cmin[k], b[6 + k]); cmax[k] = fmaxf(cmax[k], b[9 + k]);
                }
        }
        for (int k = 0; k < 3; ++k) { nodes[idx].lo[k] = blo[k]; nodes[idx].hi[k] = bhi[k]; }
        auto leaf = [&]() {
            nodes[idx].left = ~lo;
            nodes[idx].right = n;
            return idx;
        };
        if (n == 1) return leaf();
        if (n <= SWEEP_MAX && depth < sah_depth_limit) {
            // few primitives: exact sweep SAH over the centroid order of each axis instead of 3 x 16 bins
            int order[3][SWEEP_MAX];
            float best = 1e30f;
            int bax = -1, bsplit = -1;
            for (int ax = 0; ax < 3; ++ax) {
                int *o = order[ax];
                for (int i = 0; i < n; ++i) o[i] = lo + i;
                std::sort(o, o + n, [&](int a, int b) { return prims[a].c[ax] < prims[b].c[ax] || (prims[a].c[ax] == prims[b].c[ax] && prims[a].id < prims[b].id); });
                float ra[SWEEP_MAX];
                float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
                for (int i = n - 1; i > 0; --i) {
                    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], prims[o[i]].lo[k]); mx[k] = fmaxf(mx[k], prims[o[i]].hi[k]); }
                    ra[i] = half_area(mn, mx);
                }
                for (int k = 0; k < 3; ++k) { mn[k] = 1e30f; mx[k] = -1e30f; }
                for (int i = 0; i < n - 1; ++i) {
                    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], prims[o[i]].lo[k]); mx[k] = fmaxf(mx[k], prims[o[i]].hi[k]); }
                    const float cost = half_area(mn, mx) * (float)(i + 1) + ra[i + 1] * (float)(n - i - 1);
                    if (cost < best) { best = cost; bax = ax; bsplit = i + 1; }
                }
            }
            const float parent = half_area(blo, bhi);
            if (n <= MAX_LEAF && (float)n * parent <= trav_cost * parent + best) return leaf();
            Prim tmp[SWEEP_MAX];
            for (int i = 0; i < n; ++i) tmp[i] = prims[order[bax][i]];
            for (int i = 0; i < n; ++i) prims[lo + i] = tmp[i];
            const int32_t l = build(nodes, lo, lo + bsplit, depth + 1, jobs, cutoff);
            const int32_t r = build(nodes, lo + bsplit, hi, depth + 1, jobs, cutoff);
            nodes[idx].left = l;
            nodes[idx].right = r;
            return idx;
        }
        int best_axis = -1, best_bin = -1;
        float best_cost = 1e30f;
        if (depth < sah_depth_limit) {
            float scs[3];
            for (int ax = 0; ax < 3; ++ax) {
                const float ext = cmax[ax] - cmin[ax];
                scs[ax] = ext > 0.0f ? (float)NB / ext : 0.0f;
            }
            std::vector<Bins> parts((size_t)T);
            pfor(T, lo, n, [&](int t, int i0, int i1) {
                bins_clear(parts[(size_t)t]);
                bins_add(parts[(size_t)t], i0, i1, cmin, scs);
            });
            Bins &all = parts[0];
            for (int t = 1; t < T; ++t)
                for (int ax = 0; ax < 3; ++ax)
                    for (int b = 0; b < NB; ++b) {
                        all.cnt[ax][b] += parts[(size_t)t].cnt[ax][b];
                        for (int k = 0; k < 3; ++k) {
                            all.bl[ax][b][k] = fminf(all.bl[ax][b][k], parts[(size_t)t].bl[ax][b][k]);
                            all.bh[ax][b][k] = fmaxf(all.bh[ax][b][k], parts[(size_t)t].bh[ax][b][k]);
                        }
                    }
            for (int ax = 0; ax < 3; ++ax) {
                if (!(scs[ax] > 0.0f)) continue;
                float(*bl)[3] = all.bl[ax];
                float(*bh)[3] = all.bh[ax];
                int *cnt = all.cnt[ax];
                float ra[NB];
                int rc[NB];
                float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
                int c = 0;
                for (int b = NB - 1; b > 0; --b) {
                    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], bl[b][k]); mx[k] = fmaxf(mx[k], bh[b][k]); }
                    c += cnt[b];
                    ra[b] = c ? half_area(mn, mx) : 0.0f;
                    rc[b] = c;
                }
                for (int k = 0; k < 3; ++k) { mn[k] = 1e30f; mx[k] = -1e30f; }
                c = 0;
                for (int b = 0; b < NB - 1; ++b) {
                    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], bl[b][k]); mx[k] = fmaxf(mx[k], bh[b][k]); }
                    c += cnt[b];
                    if (c == 0 || rc[b + 1] == 0) continue;
                    float cost = half_area(mn, mx) * (float)c + ra[b + 1] * (float)rc[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
                }
            }
        }
        int mid;
        if (best_axis < 0) {
            if (n <= MAX_LEAF) return leaf();
            // coincident centroids or past the SAH depth limit: balanced median split along the widest centroid axis
            int ax = 0;
            if (cmax[1] - cmin[1] > cmax[ax] - cmin[ax]) ax = 1;
            if (cmax[2] - cmin[2] > cmax[ax] - cmin[ax]) ax = 2;
            mid = lo + n / 2;
            std::nth_element(prims.begin() + lo, prims.begin() + mid, prims.begin() + hi, [ax](const Prim &a, const Prim &b) { return a.c[ax] < b.c[ax]; });
        } else {
            const float parent = half_area(blo, bhi);
            // SAH termination: one traversal step costs about `trav_cost` triangle tests
            if (n <= MAX_LEAF && (float)n * parent <= trav_cost * parent + best_cost) return leaf();
            const float sc = (float)NB / (cmax[best_axis] - cmin[best_axis]);
            const float cm = cmin[best_axis];
            const int ax = best_axis, bb = best_bin;
            auto left = [&](const Prim &q) { return std::min(NB - 1, std::max(0, (int)((q.c[ax] - cm) * sc))) <= bb; };
            if (n >= PAR_MIN) { // stable, by all threads of this node: counts per slice, then a scatter through the scratch array
                if (scratch.size() < prims.size()) scratch.resize(prims.size());
                std::vector<int> n_left((size_t)T + 1, 0);
                pfor(T, lo, n, [&](int t, int i0, int i1) {
                    int c = 0;
                    for (int i = i0; i < i1; ++i) c += left(prims[i]) ? 1 : 0;
                    n_left[(size_t)t + 1] = c;
                });
                for (int t = 0; t < T; ++t) n_left[(size_t)t + 1] += n_left[(size_t)t];
                const int total_left = n_left[(size_t)T];
                pfor(T, lo, n, [&](int t, int i0, int i1) {
                    int l = lo + n_left[(size_t)t], r = lo + total_left + (i0 - lo) - n_left[(size_t)t];
                    for (int i = i0; i < i1; ++i) {
                        if (left(prims[i])) scratch[(size_t)l++] = prims[i];
                        else scratch[(size_t)r++] = prims[i];
                    }
                });
                pfor(T, lo, n, [&](int, int i0, int i1) { std::copy(scratch.begin() + i0, scratch.begin() + i1, prims.begin() + i0); });
                mid = lo + total_left;
            } else {
                auto it = std::partition(prims.begin() + lo, prims.begin() + hi, left);
                mid = (int)(it - prims.begin());
            }
            if (mid == lo || mid == hi) mid = lo + n / 2;
        }
        const int32_t l = build(nodes, lo, mid, depth + 1, jobs, cutoff);
        const int32_t r = build(nodes, mid, hi, depth + 1, jobs, cutoff);
        nodes[idx].left = l;
        nodes[idx].right = r;
        return idx;
    }
    void build_all() {
        const int n = (int)prims.size();
        nodes.clear();
        nodes.reserve(2 * (size_t)n);
        std::vector<Job> jobs;
        scratch.clear();
        const int n_threads = std::getenv("RPTR_BUILD_THREADS") ? std::max(1, std::atoi(std::getenv("RPTR_BUILD_THREADS")))
                                                                : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        par_threads = n_threads;
        build(nodes, 0, n, 0, n_threads > 1 && n > 65536 ? &jobs : nullptr, std::max(4096, n / (8 * n_threads)));
        par_threads = 1;
        if (jobs.empty()) return;
        if (std::getenv("RPTR_BUILD_VERBOSE")) fprintf(stderr, "bvh: top pass done, %zu jobs, %d threads\n", jobs.size(), n_threads);
        std::vector<std::vector<Node2>> local(jobs.size());
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t j = next.fetch_add(1); j < jobs.size(); j = next.fetch_add(1)) {
                local[j].reserve(2 * (size_t)(jobs[j].hi - jobs[j].lo));
                build(local[j], jobs[j].lo, jobs[j].hi, jobs[j].depth, nullptr, 0);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (std::thread &t : pool) t.join();
        if (std::getenv("RPTR_BUILD_VERBOSE")) fprintf(stderr, "bvh: workers joined\n");
        // splice: local node 0 replaces the placeholder, local node i >= 1 lands at off + i
        for (size_t j = 0; j < jobs.size(); ++j) {
            const int32_t off = (int32_t)nodes.size() - 1;
            auto fix = [&](Node2 nd) {
                if (nd.left >= 0) { nd.left += off; nd.right += off; }
                return nd;
            };
            nodes[jobs[j].node] = fix(local[j][0]);
            for (size_t i = 1; i < local[j].size(); ++i) nodes.push_back(fix(local[j][i]));
        }
    }
};

// collapse the binary tree into 8-wide nodes (rptr_bvh.cuh), emitted breadth-first with the inner children and the triangles
// of every node consecutive and in slot order; returns the depth of the wide tree
int collapse(const Builder &b, HostScene &s) {
    // a child under construction: a binary node (inner, or a leaf of several primitives) or one primitive of the builder's array
    struct Kid { int32_t node2; int32_t prim; };
    struct Item { int32_t node2; int32_t depth; };
    std::vector<Item> queue;
    s.nodes.clear();
    s.leaf_tris.clear();
    s.sah_cost = 0.0f;
    auto is_leaf2 = [&](int32_t i) { return b.nodes[i].left < 0; };
    auto leaf_count = [&](int32_t i) { return b.nodes[i].right; };
    auto leaf_first = [&](int32_t i) { return ~b.nodes[i].left; };
    auto kid_area = [&](const Kid &k) {
        return k.node2 >= 0 ? Builder::half_area(b.nodes[k.node2].lo, b.nodes[k.node2].hi) : Builder::half_area(b.prims[k.prim].lo, b.prims[k.prim].hi);
    };
    // SAH-optimal collapse by dynamic programming over the binary tree (after Ylitie et al. 2017, section 3): cost[n][i - 1] = the
    // cheapest way to represent the sub-tree of binary node n as a forest of at most i roots (i = 1 .. 7), a root being a triangle
    // slot (cost: its box area x c_tri) or a wide node (its box area x c_node + the best distribution of <= 8 roots over its two
    // sub-trees).  Binary nodes are stored parent before children, so one backward sweep sees the children first.
    const float c_node = std::getenv("RPTR_COLLAPSE_NODE_COST") ? (float)std::atof(std::getenv("RPTR_COLLAPSE_NODE_COST")) : 1.0f;
    const float c_tri = std::getenv("RPTR_COLLAPSE_TRI_COST") ? (float)std::atof(std::getenv("RPTR_COLLAPSE_TRI_COST")) : 0.6f;
    const bool greedy = std::getenv("RPTR_COLLAPSE_GREEDY") != nullptr; // A/B: open the largest child until eight are held
    const size_t n2 = b.nodes.size();
    struct Dp { float cost[7]; uint8_t split[8]; }; // split[i - 1]: 0 = as for i - 1 roots (i > 1) / one root (i == 1); k > 0 = k roots left, rest right; split[7] = the wide node's own split
    std::vector<Dp> dp(greedy ? 0 : n2);
    if (!greedy)
        for (size_t idx = n2; idx-- > 0;) {
            Dp &e = dp[idx];
            const float area = Builder::half_area(b.nodes[idx].lo, b.nodes[idx].hi);
            if (is_leaf2((int32_t)idx)) {
                const int cnt = leaf_count((int32_t)idx);
                float tris = 0.0f;
                for (int k = 0; k < cnt; ++k) tris += Builder::half_area(b.prims[leaf_first((int32_t)idx) + k].lo, b.prims[leaf_first((int32_t)idx) + k].hi) * c_tri;
                for (int i = 1; i <= 7; ++i) {
                    e.cost[i - 1] = i >= cnt ? tris : area * c_node + tris; // fewer roots than triangles: one wide node of triangle slots
                    e.split[i - 1] = 0;
                }
                e.split[7] = 0;
                continue;
            }
            const Dp &l = dp[(size_t)b.nodes[idx].left], &r = dp[(size_t)b.nodes[idx].right];
            // as a wide node: up to eight roots distributed over the two sub-trees
            float wide = 1e30f;
            int wide_k = 1;
            for (int k = 1; k <= 7; ++k) {
                const float c = l.cost[k - 1] + r.cost[8 - k - 1];
                if (c < wide) { wide = c; wide_k = k; }
            }
            e.split[7] = (uint8_t)wide_k;
            e.cost[0] = area * c_node + wide;
            e.split[0] = 0;
            for (int i = 2; i <= 7; ++i) {
                float best = e.cost[i - 2];
                int best_k = 0;
                for (int k = 1; k < i; ++k) {
                    const float c = l.cost[k - 1] + r.cost[i - k - 1];
                    if (c < best) { best = c; best_k = k; }
                }
                e.cost[i - 1] = best;
                e.split[i - 1] = (uint8_t)best_k;
            }
        }
    // the roots of the cheapest forest of at most i roots for the sub-tree of binary node m
    std::function<void(int32_t, int, Kid *, int &)> expand = [&](int32_t m, int i, Kid *kids, int &nk) {
        if (is_leaf2(m)) {
            const int cnt = leaf_count(m);
            if (i >= cnt) for (int k = 0; k < cnt; ++k) kids[nk++] = Kid{-1, leaf_first(m) + k};
            else kids[nk++] = Kid{m, -1};
            return;
        }
        while (i > 1 && dp[(size_t)m].split[i - 1] == 0) --i;
        if (i == 1) { kids[nk++] = Kid{m, -1}; return; }
        const int k = dp[(size_t)m].split[i - 1];
        expand(b.nodes[m].left, k, kids, nk);
        expand(b.nodes[m].right, i - k, kids, nk);
    };
    int max_depth = 0;
    queue.push_back(Item{0, 0});
    for (size_t head = 0; head < queue.size(); ++head) {
        const Item it = queue[head];
        max_depth = std::max(max_depth, (int)it.depth);
        Kid kids[RPTR_BVH_WIDTH];
        int nk = 0;
        if (is_leaf2(it.node2)) { // a leaf of the binary tree: its primitives are the slots
            for (int k = 0; k < leaf_count(it.node2) && nk < RPTR_BVH_WIDTH; ++k) kids[nk++] = Kid{-1, leaf_first(it.node2) + k};
        } else if (!greedy) {
            const int k = dp[(size_t)it.node2].split[7];
            expand(b.nodes[it.node2].left, k, kids, nk);
            expand(b.nodes[it.node2].right, 8 - k, kids, nk);
        } else {
            kids[nk++] = Kid{b.nodes[it.node2].left, -1};
            kids[nk++] = Kid{b.nodes[it.node2].right, -1};
            for (;;) { // open the child with the largest surface area that still fits: an inner node -> its two children, a leaf -> its primitives
                int pick = -1;
                float best = -1.0f;
                for (int k = 0; k < nk; ++k) {
                    if (kids[k].node2 < 0) continue;
                    const int parts = is_leaf2(kids[k].node2) ? leaf_count(kids[k].node2) : 2;
                    if (nk - 1 + parts > RPTR_BVH_WIDTH) continue;
                    const float a = kid_area(kids[k]);
                    if (a > best) { best = a; pick = k; }
                }
                if (pick < 0) break;
                const int32_t open = kids[pick].node2;
                if (is_leaf2(open)) {
                    kids[pick] = Kid{-1, leaf_first(open)};
                    for (int k = 1; k < leaf_count(open); ++k) kids[nk++] = Kid{-1, leaf_first(open) + k};
                } else {
                    kids[pick] = Kid{b.nodes[open].left, -1};
                    kids[nk++] = Kid{b.nodes[open].right, -1};
                }
            }
        }
        float klo[RPTR_BVH_WIDTH][3], khi[RPTR_BVH_WIDTH][3];
        for (int k = 0; k < nk; ++k) {
            const float *lo = kids[k].node2 >= 0 ? b.nodes[kids[k].node2].lo : b.prims[kids[k].prim].lo;
            const float *hi = kids[k].node2 >= 0 ? b.nodes[kids[k].node2].hi : b.prims[kids[k].prim].hi;
            for (int a = 0; a < 3; ++a) { klo[k][a] = lo[a]; khi[k][a] = hi[a]; }
            s.sah_cost += kid_area(kids[k]);
        }
        int slot_of[RPTR_BVH_WIDTH];
        assign_slots(nk, klo, khi, slot_of);
        int kid_in[RPTR_BVH_WIDTH];
        for (int sl = 0; sl < RPTR_BVH_WIDTH; ++sl) kid_in[sl] = -1;
        for (int k = 0; k < nk; ++k) kid_in[slot_of[k]] = k;
        float slo[RPTR_BVH_WIDTH][3], shi[RPTR_BVH_WIDTH][3];
        int kind[RPTR_BVH_WIDTH];
        const int32_t child_base = (int32_t)queue.size(), tri_base = (int32_t)s.leaf_tris.size();
        for (int sl = 0; sl < RPTR_BVH_WIDTH; ++sl) {
            const int k = kid_in[sl];
            kind[sl] = 0;
            for (int a = 0; a < 3; ++a) { slo[sl][a] = 0.0f; shi[sl][a] = 0.0f; }
            if (k < 0) continue;
            for (int a = 0; a < 3; ++a) { slo[sl][a] = klo[k][a]; shi[sl][a] = khi[k][a]; }
            if (kids[k].node2 >= 0) {
                kind[sl] = 1;
                queue.push_back(Item{kids[k].node2, it.depth + 1});
            } else {
                kind[sl] = 2;
                s.leaf_tris.push_back(s.tris[b.prims[kids[k].prim].id]);
            }
        }
        s.nodes.push_back(encode_node(slo, shi, kind, child_base, tri_base));
    }
    return max_depth + 1;
}

} // namespace

void build_bvh(HostScene &s) {
    auto t0 = std::chrono::steady_clock::now();
    s.nodes.clear();
    s.leaf_tris.clear();
    s.sah_cost = 0.0f;
    if (s.tris.empty()) return;
    Builder b;
    b.prims.resize(s.tris.size());
    float extent = 0.0f; // largest |coordinate| of the scene: scale of the absolute part of the box padding
    for (const Tri &t : s.tris)
        extent = fmaxf(extent, fmaxf(fmaxf(fabsf(t.v0x), fabsf(t.v0y)), fabsf(t.v0z)) + fmaxf(fmaxf(fabsf(t.e1x), fabsf(t.e1y)), fabsf(t.e1z)) +
                                   fmaxf(fmaxf(fabsf(t.e2x), fabsf(t.e2y)), fabsf(t.e2z)));
    const float abs_pad = 7.62939453125e-06f * extent; // 2^-17 * extent
    for (size_t i = 0; i < s.tris.size(); ++i) {
        const Tri &t = s.tris[i];
        Prim &p = b.prims[i];
        const float v[3][3] = {{t.v0x, t.v0y, t.v0z}, {t.v0x + t.e1x, t.v0y + t.e1y, t.v0z + t.e1z}, {t.v0x + t.e2x, t.v0y + t.e2y, t.v0z + t.e2z}};
        for (int k = 0; k < 3; ++k) {
            float lo = fminf(v[0][k], fminf(v[1][k], v[2][k])), hi = fmaxf(v[0][k], fmaxf(v[1][k], v[2][k]));
            // conservative padding: box culling may never reject what intersect_tri accepts.  Both tests round at
            // ulp(|origin|) ~ 6e-8 |origin|, so the pad has a part relative to the coordinates (2^-16) and a part
            // relative to the scene extent (2^-17): safe for ray origins within ~16 scene extents.
            float pad = 1.52587890625e-05f * fmaxf(fabsf(lo), fabsf(hi)) + abs_pad + 1e-30f;
            p.lo[k] = lo - pad;
            p.hi[k] = hi + pad;
            p.c[k] = 0.5f * (p.lo[k] + p.hi[k]);
        }
        p.id = (int32_t)i;
    }
    // The traversal stack holds RPTR_STACK_SIZE entries (3 pushes per level): rebuild with an earlier switch to balanced
    // median splits until the wide tree is at most RPTR_MAX_BVH_DEPTH deep (never needed for sane inputs).
    for (int limit = 32;; limit -= 8) {
        std::vector<Prim> keep = b.prims;
        b.sah_depth_limit = limit;
        const auto tb0 = std::chrono::steady_clock::now();
        b.build_all();
        if (std::getenv("RPTR_BUILD_VERBOSE")) fprintf(stderr, "bvh: binary build %.0f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count());
        s.leaf_tris.reserve(s.tris.size());
        const auto tc0 = std::chrono::steady_clock::now();
        const int depth = collapse(b, s);
        if (std::getenv("RPTR_BUILD_VERBOSE")) fprintf(stderr, "bvh: collapse %.0f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
        if (depth <= RPTR_MAX_BVH_DEPTH || limit <= 0) {
            if (depth > RPTR_MAX_BVH_DEPTH) throw std::runtime_error("BVH too deep for the traversal stack");
            break;
        }
        b.prims.swap(keep);
    }
    s.bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace rp
