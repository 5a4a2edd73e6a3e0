// oracle/shading_oracle.h -- TEST INFRASTRUCTURE (CPU oracle). Never included by the product.
//
// Scalar restatement of the reference's shading library for the megakernel path (PT_MEGAKERNEL, LCG RNG),
// written against the RPTR-FP arithmetic contract in fp32.h.  Each function cites the reference lines it follows.
#pragma once
#include "fp32.h"
#include "../include/rptr_types.h"

namespace orc {
using namespace fp;

// ---- RNG: rendering/pointsets/hashing.glsl:11-39, lcg_rng.glsl:15-39 -----------------------------------------
static inline uint32_t murmur_mix(uint32_t hash, uint32_t k) {
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    hash ^= k;
    hash = ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
    return hash;
}
static inline uint32_t murmur_finalize(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
struct Lcg { uint32_t state; };
static inline Lcg lcg_seed(uint32_t index, uint32_t frame, uint32_t linear) {
    uint32_t s = murmur_mix(frame, linear);
    s = murmur_mix(s, index);
    return Lcg{murmur_finalize(s)};
}
static inline uint32_t lcg_next(Lcg &r) {
    r.state = r.state * 1664525u + 1013904223u;
    return r.state;
}
static inline float lcg_randomf(Lcg &r) { return (float)lcg_next(r) * 2.3283064365386963e-10f; /* ldexp(.,-32) */ }

// ---- rendering/util.glsl:70-92 -------------------------------------------------------------------------------
static inline void ortho_basis(V3 &vx, V3 &vy, V3 n) {
    vy = v3(0.0f);
    if (n.x < 0.6f && n.x > -0.6f) vy.x = 1.0f;
    else if (n.y < 0.6f && n.y > -0.6f) vy.y = 1.0f;
    else if (n.z < 0.6f && n.z > -0.6f) vy.z = 1.0f;
    else vy.x = 1.0f;
    vx = normalize(cross(vy, n));
    vy = normalize(cross(n, vx));
}
static inline float pow2(float x) { return x * x; }
static inline float cos_half_angle(float c) { return (1.0f + c) / sqrtf(2.0f + 2.0f * c); }
static inline float mix_fma(float x, float y, float a) { return fmaf(a, y, fmaf(-a, x, x)); }

// ---- material: rendering/rt/material_textures.glsl:95-135 + bsdfs/gltf_bsdf.glsl:14-62 ------------------------
struct GltfMat {
    V3 base_color;
    float metallic, specular, roughness, ior;
    float specular_transmission, transmission_roughness;
    V3 transmission_color;
    uint32_t flags;
};
static inline bool is_textured(float v) { return (f2u(v) & 0x80000000u) != 0; }

// ---- textures: rendering/rt/material_textures.glsl:37-63 -----------------------------------------------------------------
// The texture unit is hardware in the reference (VkSampler of vulkan/render_vulkan.cpp:1655-1671: LINEAR mag / min / mip filters,
// REPEAT addressing, LOD range [0, 16], 12x anisotropy); this is the statement oracle and product share (RPTR-FP): texel centres
// at (i + 0.5) / size, bilinear weights and blends in binary32 as written, UNORM8 -> v / 255, colour channels of an sRGB image
// through the sRGB transfer function (evaluated in double and rounded once) per texel before filtering.  textureGrad follows the
// Vulkan specification's formulas for scale factor, level of detail and anisotropic filtering (the part left to implementations):
// rho_x / rho_y = lengths of the derivatives in base-level texels, eta = min(rho_max / rho_min, 12), N = ceil(eta) taps spread
// along the major derivative at offsets i / (N + 1) - 1 / 2, lambda = log2(rho_max / eta) clamped to the stored levels, linear
// blend of the two nearest levels.  The oracle's textures are RGBA8 with all mip levels back to back (block-compressed input is
// decoded by oracle_scene_create).
struct TextureSet {
    const rptr_texture_desc *tex = nullptr;
    int n = 0;
    struct RGBA { float r, g, b, a; };
    static float decode(int v, bool srgb) {
        if (!srgb) return (float)v / 255.0f;
        double c = (double)v / 255.0;
        return (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
    static int levels_of(const rptr_texture_desc &t) { return t.mip_levels > 0 ? t.mip_levels : 1; }
    struct Level { const uint8_t *px; int w, h; };
    static Level level_of(const rptr_texture_desc &t, int level) {
        Level L{t.texels, t.width, t.height};
        for (int l = 0; l < level; ++l) {
            L.px += (size_t)L.w * L.h * t.channels;
            L.w = L.w > 1 ? L.w / 2 : 1;
            L.h = L.h > 1 ? L.h / 2 : 1;
        }
        return L;
    }
    static RGBA texel_at(const rptr_texture_desc &t, const Level &L, int x, int y) {
        const bool srgb = t.color_space == RPTR_COLOR_SPACE_SRGB;
        const uint8_t *px = L.px + ((size_t)y * L.w + x) * t.channels;
        float ch[4] = {0.0f, 0.0f, 0.0f, 1.0f};
        for (int k = 0; k < t.channels && k < 4; ++k) ch[k] = k == 3 ? (float)px[3] / 255.0f : decode(px[k], srgb);
        return RGBA{ch[0], ch[1], ch[2], ch[3]};
    }
    RGBA texel(uint32_t id) const { return texel_at(tex[id], level_of(tex[id], 0), 0, 0); }
    static int wrap(int i, int n) {
        int r = i % n;
        return r < 0 ? r + n : r;
    }
    static float blend(float a, float b, float t) { return a + (b - a) * t; }
    static RGBA blend(RGBA a, RGBA b, float t) { return RGBA{blend(a.r, b.r, t), blend(a.g, b.g, t), blend(a.b, b.b, t), blend(a.a, b.a, t)}; }
    static RGBA bilinear(const rptr_texture_desc &t, int level, V2 uv) {
        const Level L = level_of(t, level);
        float x = uv.x * (float)L.w - 0.5f, y = uv.y * (float)L.h - 0.5f;
        if (!(fabsf(x) < 1.0e9f) || !(fabsf(y) < 1.0e9f)) { x = 0.0f; y = 0.0f; }
        const float xf = floorf(x), yf = floorf(y);
        const float wx = x - xf, wy = y - yf;
        const int x0 = wrap((int)xf, L.w), x1 = wrap(x0 + 1, L.w), y0 = wrap((int)yf, L.h), y1 = wrap(y0 + 1, L.h);
        const RGBA a = texel_at(t, L, x0, y0), b = texel_at(t, L, x1, y0), c = texel_at(t, L, x0, y1), d = texel_at(t, L, x1, y1);
        return blend(blend(a, b, wx), blend(c, d, wx), wy);
    }
    RGBA sample(uint32_t id, V2 uv) const { return bilinear(tex[id], 0, uv); }               // base level (alpha candidates: zero footprint)
    RGBA sample_lod(uint32_t id, V2 uv, int level) const {                                      // textureLod with a whole level (normal maps)
        const int top = levels_of(tex[id]) - 1;
        return bilinear(tex[id], level < top ? level : top, uv);
    }
    static float log2_positive(float x) { // exponent + Cephes logf kernel, the contract's log2
        uint32_t bits = f2u(x);
        if (bits < 0x00800000u) return -127.0f;
        int e = (int)(bits >> 23) - 127;
        float m = u2f((bits & 0x007fffffu) | 0x3f800000u);
        if (m > 1.41421356237f) { m *= 0.5f; e += 1; }
        const float f = m - 1.0f, z = f * f;
        const float c[9] = {7.0376836292e-2f, -1.1514610310e-1f, 1.1676998740e-1f, -1.2420140846e-1f, 1.4249322787e-1f,
                            -1.6668057665e-1f, 2.0000714765e-1f, -2.4999993993e-1f, 3.3333331174e-1f};
        float p = c[0];
        for (int k = 1; k < 9; ++k) p = fmaf(f, p, c[k]);
        const float y = fmaf(-0.5f, z, f * z * p);
        return fmaf(f + y, 1.44269504088896341f, (float)e);
    }
    RGBA sample_grad(uint32_t id, V2 uv, V2 dx, V2 dy) const { // textureGrad(sampler, uv, dPdx, dPdy)
        const rptr_texture_desc &t = tex[id];
        if (t.width == 1 && t.height == 1) return bilinear(t, 0, uv); // one texel: every tap of every level is that texel (no averaging error)
        const V2 mx{dx.x * (float)t.width, dx.y * (float)t.height}, my{dy.x * (float)t.width, dy.y * (float)t.height};
        const float rho_x = sqrtf(dot(mx, mx)), rho_y = sqrtf(dot(my, my));
        const float rho_max = fmaxf(rho_x, rho_y), rho_min = fminf(rho_x, rho_y);
        if (!(rho_max > 0.0f) || !(rho_max < 1.0e18f) || rho_min != rho_min) return bilinear(t, 0, uv);
        const float eta = rho_min > 0.0f ? fminf(rho_max / rho_min, 12.0f) : 12.0f;
        const int taps = (int)ceilf(eta);
        const int top = levels_of(t) - 1;
        const float lambda = fminf(fmaxf(log2_positive(rho_max / eta), 0.0f), (float)top);
        const float lo = floorf(lambda), w_hi = lambda - lo;
        const int level_lo = (int)lo, level_hi = level_lo + 1 <= top ? level_lo + 1 : top;
        const V2 major = rho_x >= rho_y ? dx : dy;
        RGBA sum_lo{0, 0, 0, 0}, sum_hi{0, 0, 0, 0};
        for (int i = 1; i <= taps; ++i) {
            const float o = (float)i / (float)(taps + 1) - 0.5f;
            const V2 at{uv.x + major.x * o, uv.y + major.y * o};
            const RGBA a = bilinear(t, level_lo, at);
            sum_lo.r += a.r; sum_lo.g += a.g; sum_lo.b += a.b; sum_lo.a += a.a;
            if (w_hi > 0.0f) {
                const RGBA b = bilinear(t, level_hi, at);
                sum_hi.r += b.r; sum_hi.g += b.g; sum_hi.b += b.b; sum_hi.a += b.a;
            }
        }
        const float nt = (float)taps;
        const RGBA avg_lo{sum_lo.r / nt, sum_lo.g / nt, sum_lo.b / nt, sum_lo.a / nt};
        if (!(w_hi > 0.0f)) return avg_lo;
        const RGBA avg_hi{sum_hi.r / nt, sum_hi.g / nt, sum_hi.b / nt, sum_hi.a / nt};
        return blend(avg_lo, avg_hi, w_hi);
    }
    static bool is_handle(float x) { return (f2u(x) & RPTR_TEXTURED_PARAM_MASK) != 0; }
    // textured_color_param(vec4(p.base_color, 1), hit)
    RGBA color_param(const float *rgb, V2 uv, V2 dx = V2{0.0f, 0.0f}, V2 dy = V2{0.0f, 0.0f}) const {
        if (is_handle(rgb[0])) return sample_grad(RPTR_GET_TEXTURE_ID(f2u(rgb[0])), uv, dx, dy);
        return RGBA{rgb[0], rgb[1], rgb[2], 1.0f};
    }
    // textured_scalar_param(x, hit)
    float scalar_param(float x, V2 uv, V2 dx = V2{0.0f, 0.0f}, V2 dy = V2{0.0f, 0.0f}) const {
        if (!is_handle(x)) return x;
        RGBA t = sample_grad(RPTR_GET_TEXTURE_ID(f2u(x)), uv, dx, dy);
        const float ch[4] = {t.r, t.g, t.b, t.a};
        return ch[RPTR_GET_TEXTURE_CHANNEL(f2u(x))];
    }
};
// get_material_alpha (material_textures.glsl:137-145)
static inline float material_alpha(const TextureSet &ts, const rptr_base_material &p, V2 uv) { return ts.color_param(p.base_color, uv).a; }

// unpack_material (material_textures.glsl:95-135, non-unrolled standard-texture semantics of rendering/rt/materials.glsl:42-49);
// returns alpha
static inline float unpack_material(GltfMat &m, V3 &emit, const rptr_base_material &p, bool transmission, const TextureSet &ts, V2 uv,
                                    V2 dx = V2{0.0f, 0.0f}, V2 dy = V2{0.0f, 0.0f}) { // dx, dy: hit.duvdxy[0], hit.duvdxy[1]
    TextureSet::RGBA texel = ts.color_param(p.base_color, uv, dx, dy);
    float alpha = texel.a;
    m.base_color = v3(texel.r, texel.g, texel.b);
    if (alpha > 0.001f) m.base_color = m.base_color / alpha; // PREMULTIPLIED_BASE_COLOR_ALPHA
    m.specular = ts.scalar_param(p.specular, uv, dx, dy);
    m.roughness = ts.scalar_param(p.roughness, uv, dx, dy);
    m.metallic = ts.scalar_param(p.metallic, uv, dx, dy);
    m.ior = ts.scalar_param(p.ior, uv, dx, dy);
    emit = v3(p.base_color[0], p.base_color[1], p.base_color[2]) * p.emission_intensity;
    if (p.emission_intensity != 0.0f) {
        if (TextureSet::is_handle(p.base_color[0])) emit = m.base_color * p.emission_intensity;
        m.base_color = v3(0.0f);
    }
    // load_material (gltf_bsdf.glsl:38-62)
    m.specular_transmission = 0.0f;
    m.transmission_color = v3(0.0f);
    m.transmission_roughness = 0.0f;
    if (transmission) {
        m.specular_transmission = ts.scalar_param(p.specular_transmission, uv, dx, dy);
        if (m.specular_transmission > 0.0f) {
            if (!(m.ior > 1.0f)) {
                alpha *= 1.0f - m.specular_transmission;
                m.specular_transmission = 0.0f;
            } else {
                m.transmission_color = m.base_color;
                m.transmission_roughness = m.roughness;
                m.roughness = sqrtf(ts.scalar_param(p.clearcoat_gloss, uv, dx, dy));
            }
        }
    }
    m.flags = p.flags;
    return alpha;
}

// ---- glTF BSDF: rendering/bsdfs/gltf_bsdf.glsl ---------------------------------------------------------------
static inline float schlick_weight(float c) { // :172-174, pow(x,5) as exact multiplies
    float x = clampf(1.0f - c, 0.0f, 1.0f);
    float x2 = x * x;
    return x2 * x2 * x;
}
static inline float gtr_2(float cos_h, float alpha) { // :193-197
    float a2 = alpha * alpha;
    return INV_PI_F * a2 / pow2(1.0f + (a2 - 1.0f) * cos_h * cos_h);
}
static inline float smith_den1(float ndo, float a2) { // :199-201
    return fabsf(ndo) + sqrtf(a2 + (1.0f - a2) * ndo * ndo);
}
static inline float smith_visibility_ggx(float ndo, float ndi, float alpha) { // :206-211
    float a = alpha * alpha;
    return 1.0f / (smith_den1(ndi, a) * smith_den1(ndo, a));
}
static inline V3 to_pipe_sample(V2 u) { // :215-221
    float s, c;
    sincos_pos(TWO_PI_F * u.x, s, c);
    return v3(c, s, u.y);
}
static inline V3 sample_sphere(V3 up) { // :224-228
    float ct = up.z * 2.0f - 1.0f;
    float st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f));
    return v3(st * up.x, st * up.y, ct);
}
static inline V3 sample_gtr_2_vndf(V3 wo, float ax, float ay, V3 up) { // :233-250
    V3 wi = normalize(v3(ax * wo.x, ay * wo.y, wo.z));
    float z = fmaf(1.0f - up.z, 1.0f + wi.z, -wi.z);
    float st = sqrtf(clampf(1.0f - z * z, 0.0f, 1.0f));
    V3 wm = v3(st * up.x, st * up.y, z) + wi;
    V3 w = v3(wm.x * ax, wm.y * ay, fmaxf(0.0f, wm.z));
    return w / length(w);
}
static inline float gtr_2_vndf_pdf(float ndo, float cos_h, float alpha) { // :253-257
    return gtr_2(cos_h, alpha) * (0.5f / smith_den1(ndo, alpha * alpha));
}
static inline V3 diffuse_basecolor(const GltfMat &m) { return m.base_color * (1.0f - m.metallic); } // :259-261
static inline V3 specular_basecolor(const GltfMat &m, float ior) { // :263-273
    float d = pow2((ior - 1.0f) / (ior + 1.0f));
    return mix(v3(d), m.base_color, m.metallic);
}
static inline float specular_alpha(const GltfMat &m) { return fmaxf(m.roughness * m.roughness, 0.002f); } // :275-277
static inline float transmission_alpha(const GltfMat &m) { // :279-281
    return fmaxf(m.transmission_roughness * m.transmission_roughness, 0.002f);
}
static inline float gltf_schlick_weight(float odh, float ior) { // :284-292
    float f = schlick_weight(odh);
    if (ior < 1.0f) {
        float cc = sqrtf(1.0f - ior * ior);
        f = mix(f, 1.0f, fminf((1.0f - odh) / (1.0f - cc), 1.0f));
    }
    return f;
}
// refract(I,N,eta) as in the GLSL spec
static inline V3 refract(V3 i, V3 n, float eta) {
    float d = dot(n, i);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return v3(0.0f);
    return i * eta - n * (eta * d + sqrtf(k));
}

// gltf_bsdf :294-358.  `tr` = compiled with GLTF_SUPPORT_TRANSMISSION[_ROUGHNESS]
static inline V3 gltf_bsdf(const GltfMat &m, V3 n, V3 wo, V3 wi, bool tr) {
    float idn = dot(n, wi), odn = dot(n, wo);
    float ior = odn < 0.0f ? 1.0f / m.ior : m.ior;
    V3 wh;
    if (idn * odn < 0.0f) {
        if (!tr) return v3(0.0f);
        if (!(m.specular_transmission > 0.0f)) return v3(0.0f);
        if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) wh = wi * (-ior) - wo;
        else wh = reflect(wi, n) + wo;
        if (!(dot(wh, n) > 0.0f)) return v3(0.0f);
    } else
        wh = wi + wo;
    wh = normalize(wh);
    float odh = dot(wo, wh), idh = dot(wi, wh);
    V3 diffuse = diffuse_basecolor(m) * INV_PI_F;
    V3 specular = v3(0.0f);
    if (m.ior > 1.0f) {
        V3 f0 = specular_basecolor(m, m.ior);
        float sa = specular_alpha(m);
        if (tr && idn * odn < 0.0f) sa = transmission_alpha(m);
        float refl = gtr_2(dot(n, wh), sa);
        refl *= smith_visibility_ggx(odn, idn, sa);
        float fw = gltf_schlick_weight(fabsf(odh), ior);
        V3 F = mix(f0, v3(1.0f), fw);
        if (tr && idn * odn < 0.0f) {
            diffuse = v3(0.0f);
            specular = m.transmission_color * (refl * (1.0f - m.metallic) * m.specular_transmission) * (v3(1.0f) - F);
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                float ac = 2.0f * odh / (idh * ior + odh);
                specular = specular * (ac * ac);
            }
        } else {
            if (tr) diffuse = diffuse * (1.0f - m.specular_transmission);
            diffuse = diffuse * (v3(1.0f) - F);
            specular = F * refl;
        }
    }
    return diffuse + specular;
}

struct Components { float w[3]; };
// gltf_component_sampler :368-392
static inline Components component_sampler(const GltfMat &m, float ior, V3 odh, V3 vis, bool tr) {
    Components c;
    float sl = luminance(specular_basecolor(m, m.ior));
    float F0 = mix(sl, 1.0f, gltf_schlick_weight(odh.x, 1.0f));
    float F1 = mix(sl, 1.0f, gltf_schlick_weight(odh.y, 1.0f));
    c.w[0] = (1.0f - F0) * vis.x * (1.0f - m.metallic) * luminance(diffuse_basecolor(m));
    c.w[1] = F1 * vis.y;
    c.w[2] = 0.0f;
    int n = 2;
    if (tr) {
        float F2 = mix(sl, 1.0f, gltf_schlick_weight(odh.z, ior));
        c.w[0] *= (1.0f - m.specular_transmission);
        c.w[2] = (1.0f - F2) * vis.z * (1.0f - m.metallic) * m.specular_transmission;
        n = 3;
    }
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) sum += c.w[i];
    if (sum > 0.0f) {
        for (int i = 0; i < n; ++i) c.w[i] /= sum;
    } else
        c.w[0] = 1.0f;
    return c;
}
// glft_sample_reuse_component :393-409
static inline int sample_reuse_component(const Components &c, float &rnd, float &prob, bool tr) {
    int comp = 0;
    float next_base = 0.0f, base = 0.0f;
    int n = tr ? 3 : 2;
    for (int i = 0; i < n; ++i) {
        float p = c.w[i];
        if (p > 0.0f && rnd >= next_base) {
            comp = i;
            prob = p;
            base = next_base;
        }
        next_base += p;
    }
    rnd = fminf(1.0f, (rnd - base) / prob);
    return comp;
}

// gltf_wpdf :414-494
static inline float gltf_wpdf(const GltfMat &m, V3 n, V3 wo, V3 wi, bool tr) {
    float idn = dot(n, wi), odn = dot(n, wo);
    float ior = odn < 0.0f ? 1.0f / m.ior : m.ior;
    float pdf = INV_PI_F * fabsf(idn);
    if (m.ior > 1.0f) {
        V3 wh;
        if (idn * odn < 0.0f) {
            if (!tr) return 0.0f;
            if (!(m.specular_transmission > 0.0f)) return 0.0f;
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) wh = wi * (-ior) - wo;
            else wh = reflect(wi, n) + wo;
            if (!(dot(wh, n) > 0.0f)) return 0.0f;
        } else
            wh = wi + wo;
        wh = normalize(wh);
        float odh = dot(wo, wh), idh = dot(wi, wh);
        float cth = dot(wh, n);
        V3 vis = v3(1.0f, 0.0f, 0.0f);
        float sa = specular_alpha(m);
        vis.y = 2.0f * fabsf(idn) / smith_den1(idn, sa * sa);
        float ta = sa;
        if (tr) {
            vis.z = vis.y;
            if (m.specular_transmission > 0.0f) {
                ta = transmission_alpha(m);
                vis.z = 2.0f * fabsf(idn) / smith_den1(idn, ta * ta);
            }
        }
        Components c = component_sampler(m, ior, v3(fabsf(odh)), vis, tr);
        if (tr && idn * odn < 0.0f) sa = ta;
        float spec = gtr_2_vndf_pdf(odn, cth, sa);
        if (tr && idn * odn < 0.0f) {
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                float ac = 2.0f * odh / (idh * ior + odh);
                spec *= ac * ac;
            }
            pdf = spec * c.w[2];
        } else {
            pdf *= c.w[0];
            pdf += spec * c.w[1];
        }
    }
    return pdf;
}

// sample_gltf_brdf :496-645.  Returns f*|cos|/pdf; pdf==0 marks failure (mis_wpdf then unspecified -> 0 here).
static inline V3 sample_gltf_brdf(const GltfMat &m, V3 n, V3 wo, V3 &wi, float &pdf, float &mis_wpdf, V2 rng_sample,
                                  V2 fresnel_sample, V3 vx, V3 vy, bool tr) {
    V3 wol = v3(dot(vx, wo), dot(vy, wo), dot(n, wo));
    float odn = wol.z;
    float ior = m.ior;
    mis_wpdf = 0.0f;
    wi = v3(0.0f);
    if (tr) {
        ior = odn < 0.0f ? 1.0f / m.ior : m.ior;
        if (odn < 0.0f) wol.z = -wol.z;
    } else if (odn < 0.0f) {
        pdf = 0.0f;
        return v3(0.0f);
    }
    V3 up = to_pipe_sample(rng_sample);
    V3 wid = normalize(n + sample_sphere(up));
    if (tr && odn < 0.0f) wid = -wid;

    float sa = specular_alpha(m);
    int comp = 0;
    float comp_pdf = 0.0f;
    Components c = {{0.0f, 0.0f, 0.0f}};
    V3 whs = v3(0.0f), wht = v3(0.0f);
    if (m.ior > 1.0f) {
        V3 odh_all = v3(0.0f), vis_all = v3(0.0f);
        odh_all.x = cos_half_angle(dot(wo, wid));
        vis_all.x = 1.0f;
        whs = sample_gtr_2_vndf(wol, sa, sa, up);
        odh_all.y = dot(wol, whs);
        float sidn = reflect(-wol, whs).z;
        vis_all.y = sidn > 0.0f ? 2.0f * sidn / smith_den1(sidn, sa * sa) : 0.0f;
        if (tr) {
            float ta = sa;
            wht = whs;
            odh_all.z = odh_all.y;
            float tidn = sidn;
            if (m.specular_transmission > 0.0f) {
                ta = transmission_alpha(m);
                wht = sample_gtr_2_vndf(wol, ta, ta, up);
                odh_all.z = dot(wol, wht);
                if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) tidn = -refract(-wol, wht, 1.0f / ior).z;
                else tidn = reflect(-wol, wht).z;
                vis_all.z = tidn > 0.0f ? 2.0f * tidn / smith_den1(tidn, ta * ta) : 0.0f;
            }
        }
        c = component_sampler(m, ior, odh_all, vis_all, tr);
        comp = sample_reuse_component(c, fresnel_sample.x, comp_pdf, tr);
    }
    float cth, idh, odh;
    if (comp == 0) {
        wi = wid;
        V3 wh = normalize(wi + wo);
        cth = dot(n, wh);
        idh = odh = dot(wo, wh);
    } else {
        if (tr && comp == 2) {
            sa = transmission_alpha(m);
            whs = wht;
        }
        V3 wh = whs;
        if (tr && odn < 0.0f) wh.z = -wh.z;
        cth = wh.z;
        wh = mat_mul(vx, vy, n, wh);
        idh = odh = dot(wo, wh);
        if (tr && comp != 1) {
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                wi = refract(-wo, wh, 1.0f / ior);
                idh = dot(wi, wh);
            } else
                wi = reflect(reflect(-wo, wh), n);
        } else
            wi = reflect(-wo, wh);
    }
    float idn = dot(n, wi);
    bool bad = tr ? ((idn * odn > 0.0f) != (comp != 2)) : !(idn > 0.0f);
    if (bad) {
        pdf = 0.0f;
        return v3(0.0f);
    }
    pdf = INV_PI_F * fabsf(idn);
    if (m.ior > 1.0f) {
        pdf *= c.w[0];
        float spec = gtr_2_vndf_pdf(odn, cth, sa);
        if (tr && idn * odn < 0.0f) {
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                float ac = 2.0f * odh / (idh * ior + odh);
                spec *= ac * ac;
            }
            pdf = spec * c.w[2];
        } else
            pdf += spec * c.w[1];
    }
    if (!(pdf > 0.0f)) return v3(0.0f);
    V3 f = gltf_bsdf(m, n, wo, wi, tr);
    mis_wpdf = gltf_wpdf(m, n, wo, wi, tr);
    return f * fabsf(idn) / pdf;
}

// ---- triangle lights: rendering/lights/tri.glsl --------------------------------------------------------------
static inline float fast_positive_atan(float y) { // :58-72
    float ay = fabsf(y);
    float rx = (ay > 1.0f) ? (1.0f / ay) : ay;
    float ry = rx * rx;
    float rz = fmaf(ry, 0.02083509974181652f, -0.08513300120830536f);
    rz = fmaf(ry, rz, 0.18014100193977356f);
    rz = fmaf(ry, rz, -0.3302994966506958f);
    ry = fmaf(ry, rz, 0.9998660087585449f);
    rz = fmaf(-2.0f * ry, rx, 0.5f * PI_F);
    rz = (ay > 1.0f) ? rz : 0.0f;
    rx = fmaf(rx, ry, rz);
    return (y < 0.0f) ? (PI_F - rx) : rx;
}
// half_triangle_solid_angle_tan :82-114
static inline float half_tri_solid_angle_tan(V3 v0, V3 v1, V3 v2, V3 &params) {
    float hs = (v0.x > 0.0f) ? -1.0f : 1.0f;
    float hk = 1.0f / (fabsf(v0.x) + 1.0f);
    V2 hyz = V2{v0.y * hk, v0.z * hk};
    float d01 = dot(v0, v1), d02 = dot(v1, v2), d12 = dot(v0, v2);
    float dh0 = fmaf(-hs, v1.x, d01);
    float dh2 = fmaf(-hs, v2.x, d12);
    V2 c0 = V2{fmaf(-dh0, hyz.x, v1.y), fmaf(-dh0, hyz.y, v1.z)};
    V2 c1 = V2{fmaf(-dh2, hyz.x, v2.y), fmaf(-dh2, hyz.y, v2.z)};
    float det = c0.x * c1.y - c1.x * c0.y; // glm determinant(mat2): m00*m11 - m10*m01
    float vol = fabsf(det);
    float d02p12 = d02 + d12;
    float opd01 = 1.0f + d01;
    params = v3(vol, d02p12, opd01);
    return vol / (opd01 + d02p12);
}
static inline float triangle_solid_angle(V3 v0, V3 v1, V3 v2, V3 &params) { // :116-119
    return 2.0f * fast_positive_atan(half_tri_solid_angle_tan(v0, v1, v2, params));
}
// sample_solid_angle_polygon :132-152
static inline V3 sample_solid_angle_polygon(V3 v0, V3 v1, V3 v2, float omega, V3 prm, V2 rnd) {
    float target = omega * rnd.x;
    V3 a0 = v1, a1 = v0, a2 = v2; // vertices[3] = { v1, v0, v2 }
    float s, c;
    sincos_pos(0.5f * target, s, c);
    V3 offset = a0 * (prm.x * c - prm.y * s) + a2 * (prm.z * s);
    float k = 2.0f * (dot(a0, offset) / dot(offset, offset));
    V3 nv2 = V3{fmaf(k, offset.x, -a0.x), fmaf(k, offset.y, -a0.y), fmaf(k, offset.z, -a0.z)};
    float s2 = dot(a1, nv2);
    float sm = mix_fma(1.0f, s2, rnd.y);
    float den = fmaf(-s2, s2, 1.0f);
    float tn = sqrtf(fmaf(-sm, sm, 1.0f) / den);
    tn = (den > 0.0f) ? tn : rnd.y;
    return a1 * fmaf(-tn, s2, sm) + nv2 * tn;
}

// ---- sky: rendering/lights/sky_model_arhosek/sky_model.glsl:40-59 -----------------------------------------------
static inline V3 skymodel_radiance(const rptr_scene_params &sp, V3 sun_dir, V3 view) {
    float ct = clampf(view.y, 0.0f, 1.0f);
    float cg = clampf(dot(view, sun_dir), -1.0f, 1.0f);
    float gamma = acos_f(ct);
    float rayM = cg * cg;
    float zenith = sqrtf(ct);
    float out[3];
    for (int ch = 0; ch < 3; ++ch) {
        const float c0 = sp.sky_configs[0][ch], c1 = sp.sky_configs[1][ch], c2 = sp.sky_configs[2][ch],
                    c3 = sp.sky_configs[3][ch], c4 = sp.sky_configs[4][ch], c5 = sp.sky_configs[5][ch],
                    c6 = sp.sky_configs[6][ch], c7 = sp.sky_configs[7][ch], c8 = sp.sky_configs[8][ch];
        float expM = exp_f(c4 * gamma);
        float b = 1.0f + c8 * c8 - 2.0f * c8 * cg;
        float mieM = (1.0f + cg * cg) / (b * sqrtf(b)); // pow(b, 1.5)
        float lhs = 1.0f + c0 * exp_f(c1 / (ct + 0.01f));
        float rhs = c2 + c3 * expM + c5 * rayM + c6 * mieM + c7 * zenith;
        out[ch] = lhs * rhs * sp.sky_radiances[ch] * 0.01f;
    }
    return v3(out[0], out[1], out[2]);
}

static inline float nee_mis_heuristic(float nf, float pf, float ng, float pg) { // mc/nee_interface.glsl:11-15
    float f = nf * pf, g = ng * pg;
    return f / (f + g);
}
static inline V3 vabs(V3 a) { return v3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline V3 vmax0(V3 a) { return v3(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f), fmaxf(a.z, 0.0f)); }

// compute_sky_illum: vulkan/pt_megakernel.glsl:113-149
static inline V3 compute_sky_illum(const rptr_scene_params &sp, V3 ray_dir, float prev_bsdf_pdf) {
    V3 dir = ray_dir;
    float ocean = 1.0f;
    if (dir.y <= 0.0f) {
        dir.y = -dir.y;
        float b = fmaxf(1.0f - fabsf(dir.y), 0.0f);
        float b2 = b * b;
        ocean = 0.7f * (b2 * b2 * b);
    }
    V3 sun_dir = v3(sp.sun_dir[0], sp.sun_dir[1], sp.sun_dir[2]);
    V3 atm = vmax0(skymodel_radiance(sp, sun_dir, dir)) * ocean;
    V3 sun = v3(0.0f);
    if (dot(dir, sun_dir) >= sp.sun_cos_angle) sun = v3(sp.sun_radiance[0], sp.sun_radiance[1], sp.sun_radiance[2]) * ocean;
    V3 illum = v3(0.0f);
    illum = illum + vabs(atm);
    // eval_direct_sun_light_pdf (mc/nee_interface.glsl:46-48, lights/sun.glsl:17-20)
    float light_pdf = sp.sun_radiance[3] * (1.0f / (TWO_PI_F * (1.0f - sp.sun_cos_angle)));
    float w = nee_mis_heuristic(1.0f, prev_bsdf_pdf, 1.0f, light_pdf);
    illum = illum + vabs(sun) * w;
    return illum;
}

// ---- quantisation: librender/dequantize.glsl ------------------------------------------------------------------
static inline V3 dequantize_position(uint64_t q, const float *scale, const float *offset) { // :8-21
    float ux = (float)(uint32_t)(q & 0x1FFFFFu);
    float uy = (float)(uint32_t)((q >> 21) & 0x1FFFFFu);
    float uz = (float)(uint32_t)((q >> 42) & 0x1FFFFFu);
    return v3(ux * scale[0] + offset[0], uy * scale[1] + offset[1], uz * scale[2] + offset[2]);
}
static inline V3 dequantize_normal(uint32_t w) { // :23-41
    float nx = (float)((int)(w & 0xFFFFu) - 0x8000) / 32767.0f;
    float ny = (float)((int)(w >> 16) - 0x8000) / 32767.0f;
    float nl1 = fabsf(nx) + fabsf(ny);
    if (nl1 >= 1.0f) {
        float tx = (1.0f - fabsf(ny)) * (nx >= 0.0f ? 1.0f : -1.0f);
        float ty = (1.0f - fabsf(nx)) * (ny >= 0.0f ? 1.0f : -1.0f);
        nx = tx;
        ny = ty;
    }
    return normalize(v3(nx, ny, 1.0f - nl1));
}
static inline V2 dequantize_uv(uint32_t w) { // :43-48
    float s = 8.0f / 65535.0f;
    return V2{0.0f + (float)(int)(w & 0xFFFFu) * s, 1.0f + (float)(-(int)(w >> 16)) * s};
}

} // namespace orc
