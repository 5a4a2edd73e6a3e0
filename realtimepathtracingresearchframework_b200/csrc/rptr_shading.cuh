// rptr_shading.cuh -- per-vertex shading of the wavefront path tracer (hit attributes, material unpack, emitter MIS,
// next-event estimation, glTF BSDF sampling, Russian roulette, sky on miss), written against rptr_math.cuh.
// Replaces, as one stage of the wavefront, what the reference runs inside its megakernel thread:
//   vulkan/pt_megakernel.glsl:113-149,578-731 + rendering/mc/{shade_base_material,nee,lights_linear}.glsl +
//   rendering/bsdfs/gltf_bsdf.glsl + rendering/lights/{tri,sun}.glsl + rendering/rt/hit.glsl.
#pragma once
#include "rptr_math.cuh"
#include "rptr_pointsets.cuh"
#include "../../include/rptr_types.h"

namespace rp {

// ---- device-side scene --------------------------------------------------------------------------------------------
struct GeomInst { // the reference's instanced_geometry[] entry (rendering/rt/geometry.h.glsl:72-98), one per (instance, geometry)
    const uint64_t *qverts;
    const uint64_t *qnuv;
    const uint8_t *tri_mat;
    float scale[3], offset[3];
    float w2o[9]; // rows of inverse(mat3(object_to_world))
    int32_t material_id;
    uint32_t flags;
    int32_t instance;
    int32_t has_normals, has_uvs;
    int32_t _pad;
};

#ifdef RPTR_TRI64
struct alignas(64) Tri { // 64 B traversal record (two 256-bit loads), world space
#else
struct Tri { // 48 B traversal record (three 128-bit loads), world space
#endif
    float v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z;
    int32_t id;        // flattened (instance, geometry, primitive) index: the closest-hit tie-break key
    int32_t gi_alpha;  // index into GeomInst[] (low 23 bits) | RPTR_TRI_TEXTURED_ALPHA | 8-bit material alpha << 24; see tri_geom_inst / tri_alpha8
    int32_t prim;
#ifdef RPTR_TRI64
    int32_t pad[4];
#endif
};

// Alpha of the triangle's material as the 8-bit texel it comes from when that is one number per material (constants, 1 x 1
// textures: rptr_host.cpp resolve_materials): 255 = opaque (NOALPHA flags, constant colour or alpha texel 255) -> traversal
// never draws for it.  Bit 23 (RPTR_TRI_TEXTURED_ALPHA) says the alpha comes from a texture larger than 1 x 1 and has to be
// looked up at the candidate's uv (candidate_alpha below).
#define RPTR_TRI_OPAQUE 255
#define RPTR_TRI_TEXTURED_ALPHA 0x00800000
RPTR_HD int32_t tri_geom_inst(const Tri &t) { return t.gi_alpha & 0x007fffff; }
RPTR_HD int32_t tri_alpha8(const Tri &t) { return (int32_t)(((uint32_t)t.gi_alpha) >> 24); }
RPTR_HD int32_t pack_gi_alpha(int32_t geom_inst, int32_t alpha8, bool textured = false) {
    return (int32_t)(((uint32_t)alpha8 << 24) | (textured ? (uint32_t)RPTR_TRI_TEXTURED_ALPHA : 0u) | (uint32_t)geom_inst);
}
RPTR_HD float alpha8_to_float(int32_t a8) { return (float)a8 / 255.0f; } // UNORM8 texel

// A texture on the device: base level only, always four 8-bit channels (missing colour channels 0, missing alpha 255).
struct TexDev {
    const uchar4 *texels;
    int32_t width, height;
    int32_t srgb; // colour channels go through the sRGB transfer function (SceneDev::srgb_lut), alpha never does
    int32_t levels; // mip levels stored back to back behind the base level: level l is max(width >> l, 1) x max(height >> l, 1)
};

struct SceneDev {
    const GeomInst *ginst;
    const rptr_base_material *materials;
    const rptr_tri_light_data *lights;
    const float4 *normal_texels; // per material: rgb = the texel of its 1 x 1 normal map as sampled; w != 0: sample textures[normal_map] at the hit's uv instead
    const TexDev *textures;      // Scene::textures (rptr_scene_desc::textures); parameters of the materials refer to them by handle
    const float *srgb_lut;       // 256 entries: sRGB8 code value -> linear (evaluated in double on the host, rounded once)
};

// ---- the texture unit (RPTR-FP statement; the reference leaves it to the hardware: VkSampler of vulkan/render_vulkan.cpp:1655-1671:
// LINEAR mag / min / mip filters, REPEAT addressing, minLod 0, maxLod 16, 12x anisotropy) ----
// Texel centres at (i + 0.5) / size, UNORM8 -> v / 255, sRGB colour channels decoded per texel BEFORE filtering (as a
// VK_FORMAT_*_SRGB image does), weights and blends in binary32 in the order written.  textureGrad (the megakernel's reads with
// USE_MIPMAPPING, rendering/rt/material_textures.glsl:37-63) follows the formulas the Vulkan specification gives for the scale
// factor, the level of detail and anisotropic filtering -- the part implementations are free to approximate, fixed here:
//   m_x = (du/dx * w, dv/dx * h), m_y likewise; rho_x = |m_x|, rho_y = |m_y|; rho_max, rho_min their max / min
//   eta = min(rho_max / rho_min, 12)  (12 when rho_min = 0);  N = ceil(eta);  lambda = log2(rho_max / eta), clamped to [0, levels - 1]
//   tau(level) = 1 / N * sum_{i = 1..N} bilinear(level, uv + d_major * (i / (N + 1) - 1 / 2)),  d_major = the derivative with the larger rho
//   result = tau(floor(lambda)) * (1 - frac) + tau(floor(lambda) + 1) * frac
// A zero or non-finite footprint (alpha candidates pass mat2(0), vulkan/pt_megakernel.glsl:204) reads the base level with one tap.
// A 1 x 1 image returns its texel whatever the footprint (the host folds such textures into the materials, rptr_host.cpp).
RPTR_HD float4 decode_texel(const SceneDev &sc, const TexDev &t, uchar4 c) {
    if (t.srgb) return f4(sc.srgb_lut[c.x], sc.srgb_lut[c.y], sc.srgb_lut[c.z], (float)c.w / 255.0f);
    return f4((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f, (float)c.w / 255.0f);
}
RPTR_HD int32_t wrap_repeat(int32_t i, int32_t n) {
    i %= n;
    return i < 0 ? i + n : i;
}
RPTR_HD float lerp_tex(float a, float b, float t) { return a + (b - a) * t; }
RPTR_HD float4 lerp_tex4(float4 a, float4 b, float t) { return f4(lerp_tex(a.x, b.x, t), lerp_tex(a.y, b.y, t), lerp_tex(a.z, b.z, t), lerp_tex(a.w, b.w, t)); }
RPTR_HD float4 sample_texture_level(const SceneDev &sc, const TexDev &t, int level, float2 uv) {
    int32_t w = t.width, h = t.height;
    size_t off = 0;
    for (int l = 0; l < level; ++l) {
        off += (size_t)w * (size_t)h;
        if (w > 1) w /= 2;
        if (h > 1) h /= 2;
    }
    const uchar4 *px = t.texels + off;
    float x = uv.x * (float)w - 0.5f, y = uv.y * (float)h - 0.5f;
    if (!(fabsf(x) < 1.0e9f) || !(fabsf(y) < 1.0e9f)) { x = 0.0f; y = 0.0f; } // NaN / out of the integer range: texel (0, 0)
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int32_t i0 = wrap_repeat((int32_t)x0, w), i1 = wrap_repeat(i0 + 1, w);
    const int32_t j0 = wrap_repeat((int32_t)y0, h), j1 = wrap_repeat(j0 + 1, h);
    const float4 c00 = decode_texel(sc, t, px[(size_t)j0 * w + i0]), c10 = decode_texel(sc, t, px[(size_t)j0 * w + i1]);
    const float4 c01 = decode_texel(sc, t, px[(size_t)j1 * w + i0]), c11 = decode_texel(sc, t, px[(size_t)j1 * w + i1]);
    return lerp_tex4(lerp_tex4(c00, c10, fx), lerp_tex4(c01, c11, fx), fy);
}
RPTR_HD float4 sample_texture(const SceneDev &sc, uint32_t id, float2 uv) { return sample_texture_level(sc, sc.textures[id], 0, uv); }
// textureLod with a whole-numbered level (the normal map read of pt_megakernel.glsl:641-647: level = bounce)
RPTR_HD float4 sample_texture_lod(const SceneDev &sc, uint32_t id, float2 uv, int level) {
    const TexDev &t = sc.textures[id];
    return sample_texture_level(sc, t, level < t.levels - 1 ? level : t.levels - 1, uv);
}
// log2 of a positive normal float: exponent + Cephes logf kernel on the mantissa in [sqrt(1/2), sqrt(2)), in +, -, x, fma only
RPTR_HD float log2_pos(float x) {
    const uint32_t u = f2u(x);
    if (u < 0x00800000u) return -127.0f; // zero / subnormal: below every level
    int e = (int)(u >> 23) - 127;
    float m = u2f((u & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356237f) { m *= 0.5f; e += 1; }
    const float f = m - 1.0f, z = f * f;
    float p = fmaf(f, 7.0376836292e-2f, -1.1514610310e-1f);
    p = fmaf(f, p, 1.1676998740e-1f);
    p = fmaf(f, p, -1.2420140846e-1f);
    p = fmaf(f, p, 1.4249322787e-1f);
    p = fmaf(f, p, -1.6668057665e-1f);
    p = fmaf(f, p, 2.0000714765e-1f);
    p = fmaf(f, p, -2.4999993993e-1f);
    p = fmaf(f, p, 3.3333331174e-1f);
    const float y = fmaf(-0.5f, z, f * z * p);
    return fmaf(f + y, 1.44269504088896341f, (float)e);
}
RPTR_HD float4 sample_texture_grad(const SceneDev &sc, uint32_t id, float2 uv, float2 ddx, float2 ddy) {
    const TexDev &t = sc.textures[id];
    if (t.width == 1 && t.height == 1) return sample_texture_level(sc, t, 0, uv);
    const float mxx = ddx.x * (float)t.width, mxy = ddx.y * (float)t.height, myx = ddy.x * (float)t.width, myy = ddy.y * (float)t.height;
    const float rx = sqrtf(fmaf(mxy, mxy, mxx * mxx)), ry = sqrtf(fmaf(myy, myy, myx * myx));
    const float rmax = fmaxf(rx, ry), rmin = fminf(rx, ry);
    if (!(rmax > 0.0f) || !(rmax < 1.0e18f) || !(rmin == rmin)) return sample_texture_level(sc, t, 0, uv);
    const float eta = rmin > 0.0f ? fminf(rmax / rmin, 12.0f) : 12.0f;
    const int n = (int)ceilf(eta);
    float lambda = log2_pos(rmax / eta);
    lambda = fminf(fmaxf(lambda, 0.0f), (float)(t.levels - 1));
    const float l0f = floorf(lambda), frac = lambda - l0f;
    const int l0 = (int)l0f, l1 = l0 + 1 < t.levels ? l0 + 1 : t.levels - 1;
    const float2 major = rx >= ry ? ddx : ddy;
    float4 a0 = f4(0.0f, 0.0f, 0.0f, 0.0f), a1 = f4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int i = 1; i <= n; ++i) {
        const float s_ = (float)i / (float)(n + 1) - 0.5f;
        const float2 p = f2(uv.x + major.x * s_, uv.y + major.y * s_);
        const float4 c0 = sample_texture_level(sc, t, l0, p);
        a0.x += c0.x; a0.y += c0.y; a0.z += c0.z; a0.w += c0.w;
        if (frac > 0.0f) {
            const float4 c1 = sample_texture_level(sc, t, l1, p);
            a1.x += c1.x; a1.y += c1.y; a1.z += c1.z; a1.w += c1.w;
        }
    }
    const float fn = (float)n;
    a0 = f4(a0.x / fn, a0.y / fn, a0.z / fn, a0.w / fn);
    if (!(frac > 0.0f)) return a0;
    a1 = f4(a1.x / fn, a1.y / fn, a1.z / fn, a1.w / fn);
    return lerp_tex4(a0, a1, frac);
}
RPTR_HD bool is_texture_handle(float v) { return (f2u(v) & RPTR_TEXTURED_PARAM_MASK) != 0; }
// textured_scalar_param (rendering/rt/material_textures.glsl:50-63): handles that survived resolve_materials refer to textures larger than 1 x 1
RPTR_HD float textured_scalar(const SceneDev &sc, float v, float2 uv, float2 ddx, float2 ddy) {
    if (!is_texture_handle(v)) return v;
    const float4 t = sample_texture_grad(sc, RPTR_GET_TEXTURE_ID(f2u(v)), uv, ddx, ddy);
    const uint32_t ch = RPTR_GET_TEXTURE_CHANNEL(f2u(v));
    return ch == 0 ? t.x : ch == 1 ? t.y : ch == 2 ? t.z : t.w;
}

struct FrameParams {
    int32_t width, height;
    float cam_pos[3], du[3], dv[3], tl[3];
    uint32_t frame_offset;
    uint32_t first_sample; // frame_id of layer 0 of this batch
    int32_t batch;
    int32_t max_path_depth, rr_path_depth, output_channel, glossy_only_mode, enable_raster_taa;
    int32_t n_lights, n_bins, bin_size;
    int32_t transmission;
    float screen_jitter[2]; // view_params.screen_jitter (raster TAA; zero unless enable_raster_taa)
    float pixel_radius;     // render_params.pixel_radius: scale of the texture footprint of a pixel (pt_megakernel.glsl:347-348)
    int32_t image_textures; // the scene has textures larger than 1 x 1: footprints are tracked and looked up (a feature-complete shade
                            // variant also runs scenes without any)
    float vp[16];           // view_params.VP, column-major (render_vulkan.cpp:2926-2930)
    float vp_reference[16]; // view_params.VP_reference: the VP of the previous begin_frame (:1986-1998, :2911)
    int32_t rng_variant;  // RenderBackendOptions::rng_variant (librender/render_params.glsl.h:34-37)
    PointsetTables pts;   // tables of the Sobol / blue-noise samplers (null unless rng_variant needs them)
    rptr_scene_params sp; // sun_radiance[3] already carries the light-count rule (vulkan/render_sky.cpp:67-70)
};

// ---- which pixel a path slot belongs to --------------------------------------------------------------------------------------
// slot = layer * local_pixels + lp.  Frames: lp enumerates the rows this GPU owns (interleaved bands of `rows` rows: band b
// belongs to rank b % world), row-major.  Ray queries (query_wgs_x > 0; render_ray_queries, vulkan/render_vulkan.cpp:1867-1876):
// lp is the query index = gl_GlobalInvocationIndex of a 2-D dispatch of 32 x 16 workgroups, query_wgs_x of them per row
// (record_frame, :3050-3056), and its "pixel" -- what seeds the samplers -- is the swizzled gl_GlobalInvocationID.xy of
// vulkan/setup_pixel_assignment.glsl:17-22, linearised with the width of the real frame (rendering/pointsets/lcg_rng.glsl:36-39).
struct TileMap {
    int32_t width, height;
    int32_t rank, world, rows;
    int32_t local_rows;
    int32_t local_pixels;
    int32_t query_wgs_x; // 0: frame; > 0: ray queries
};
RPTR_HD int32_t local_row_to_global(const TileMap &t, int32_t lr) {
    int32_t band = lr / t.rows;
    return (band * t.world + t.rank) * t.rows + lr % t.rows;
}
RPTR_HD void query_pixel(uint32_t q, uint32_t wgs_x, uint32_t &px, uint32_t &py) {
    const uint32_t wg = q >> 9, l = q & 511u; // WORKGROUP_SIZE 32 x 16 (vulkan/gpu_params.glsl, CMakeLists.txt:53-55)
    const uint32_t gx = (wg % wgs_x) * 32u + (l & 31u), gy = (wg / wgs_x) * 16u + (l >> 5);
    px = (gx & ~0x18u) + ((gy & 0x3u) << 3);
    py = (gy & ~0x3u) + ((gx & 0x18u) >> 3);
}
RPTR_HD void tile_pixel(const TileMap &t, uint32_t lp, uint32_t &px, uint32_t &py) {
    if (t.query_wgs_x > 0) {
        query_pixel(lp, (uint32_t)t.query_wgs_x, px, py);
        return;
    }
    px = lp % (uint32_t)t.width;
    py = (uint32_t)local_row_to_global(t, (int32_t)(lp / (uint32_t)t.width));
}
RPTR_HD uint32_t tile_pixel_linear(const TileMap &t, uint32_t lp) {
    uint32_t px, py;
    tile_pixel(t, lp, px, py);
    return px + py * (uint32_t)t.width;
}

// ---- RNG: rendering/pointsets/hashing.glsl:11-39, lcg_rng.glsl:15-39 ------------------------------------------------
RPTR_HD uint32_t murmur_mix(uint32_t hash, uint32_t k) {
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    hash ^= k;
    hash = ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
    return hash;
}
RPTR_HD uint32_t murmur_finalize(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
RPTR_HD uint32_t lcg_seed(uint32_t index, uint32_t frame, uint32_t linear) {
    uint32_t s = murmur_mix(frame, linear);
    s = murmur_mix(s, index);
    return murmur_finalize(s);
}
RPTR_HD float lcg_randomf(uint32_t &state) {
    state = state * 1664525u + 1013904223u;
    return (float)state * 2.3283064365386963e-10f; // ldexp(float(state), -32)
}

// ---- rendering/util.glsl:70-92 ---------------------------------------------------------------------------------------
RPTR_HD void ortho_basis(float3 &vx, float3 &vy, float3 n) {
    vy = f3(0.0f);
    if (n.x < 0.6f && n.x > -0.6f) vy.x = 1.0f;
    else if (n.y < 0.6f && n.y > -0.6f) vy.y = 1.0f;
    else if (n.z < 0.6f && n.z > -0.6f) vy.z = 1.0f;
    else vy.x = 1.0f;
    vx = normalize(cross(vy, n));
    vy = normalize(cross(n, vx));
}
RPTR_HD float cos_half_angle(float c) { return (1.0f + c) / sqrtf(2.0f + 2.0f * c); }
RPTR_HD float mix_fma(float x, float y, float a) { return fmaf(a, y, fmaf(-a, x, x)); }

// ---- material ----------------------------------------------------------------------------------------------------------
struct GltfMat {
    float3 base_color;
    float metallic, specular, roughness, ior;
    float specular_transmission, transmission_roughness;
    float3 transmission_color;
    uint32_t flags;
};

// unpack_material + load_material (rendering/rt/material_textures.glsl:95-135, gltf_bsdf.glsl:38-62; non-unrolled standard-texture
// semantics of rendering/rt/materials.glsl:42-49).  Parameters that refer to 1 x 1 textures were folded into constants on the
// host; TEX = false compiles the lookups of larger textures out (scenes without any).
template <bool TEX>
RPTR_HD float unpack_material(GltfMat &m, float3 &emit, const rptr_base_material &p, bool transmission, const SceneDev &sc, float2 uv,
                             float2 ddx = f2(0.0f, 0.0f), float2 ddy = f2(0.0f, 0.0f)) { // ddx / ddy: hit.duvdxy[0], [1] (textureGrad)
    float alpha = 1.0f;
    m.base_color = f3(p.base_color[0], p.base_color[1], p.base_color[2]);
    if (TEX && is_texture_handle(p.base_color[0])) { // textured_color_param(vec4(base_color, 1), hit)
        const float4 t = sample_texture_grad(sc, RPTR_GET_TEXTURE_ID(f2u(p.base_color[0])), uv, ddx, ddy);
        m.base_color = f3(t.x, t.y, t.z);
        alpha = t.w;
    }
    if (alpha > 0.001f) m.base_color = m.base_color / alpha; // PREMULTIPLIED_BASE_COLOR_ALPHA
    m.specular = TEX ? textured_scalar(sc, p.specular, uv, ddx, ddy) : p.specular;
    m.roughness = TEX ? textured_scalar(sc, p.roughness, uv, ddx, ddy) : p.roughness;
    m.metallic = TEX ? textured_scalar(sc, p.metallic, uv, ddx, ddy) : p.metallic;
    m.ior = TEX ? textured_scalar(sc, p.ior, uv, ddx, ddy) : p.ior;
    emit = f3(p.base_color[0], p.base_color[1], p.base_color[2]) * p.emission_intensity;
    if (p.emission_intensity != 0.0f) m.base_color = f3(0.0f); // (emitters with a textured colour are refused by set_scene)
    m.specular_transmission = 0.0f;
    m.transmission_color = f3(0.0f);
    m.transmission_roughness = 0.0f;
    if (transmission) {
        m.specular_transmission = TEX ? textured_scalar(sc, p.specular_transmission, uv, ddx, ddy) : p.specular_transmission;
        if (m.specular_transmission > 0.0f) {
            if (!(m.ior > 1.0f)) {
                alpha *= 1.0f - m.specular_transmission;
                m.specular_transmission = 0.0f;
            } else {
                m.transmission_color = m.base_color;
                m.transmission_roughness = m.roughness;
                m.roughness = sqrtf(TEX ? textured_scalar(sc, p.clearcoat_gloss, uv, ddx, ddy) : p.clearcoat_gloss);
            }
        }
    }
    m.flags = p.flags;
    return alpha;
}
// the constants-only form (no texture larger than 1 x 1 in reach)
RPTR_HD float unpack_material(GltfMat &m, float3 &emit, const rptr_base_material &p, bool transmission) {
    return unpack_material<false>(m, emit, p, transmission, SceneDev{}, f2(0.0f, 0.0f));
}

// ---- glTF BSDF (rendering/bsdfs/gltf_bsdf.glsl:172-645) -----------------------------------------------------------------
RPTR_HD float schlick_weight(float c) {
    float x = clampf(1.0f - c, 0.0f, 1.0f);
    float x2 = x * x;
    return x2 * x2 * x;
}
RPTR_HD float gtr_2(float cos_h, float alpha) {
    float a2 = alpha * alpha;
    return RPTR_INV_PI * a2 / pow2(1.0f + (a2 - 1.0f) * cos_h * cos_h);
}
RPTR_HD float smith_den1(float ndo, float a2) { return fabsf(ndo) + sqrtf(a2 + (1.0f - a2) * ndo * ndo); }
RPTR_HD float smith_visibility_ggx(float ndo, float ndi, float alpha) {
    float a = alpha * alpha;
    return 1.0f / (smith_den1(ndi, a) * smith_den1(ndo, a));
}
RPTR_HD float3 to_pipe_sample(float2 u) {
    float s, c;
    sincos_pos(RPTR_TWO_PI * u.x, s, c);
    return f3(c, s, u.y);
}
RPTR_HD float3 sample_sphere(float3 up) {
    float ct = up.z * 2.0f - 1.0f;
    float st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f));
    return f3(st * up.x, st * up.y, ct);
}
RPTR_HD float3 sample_gtr_2_vndf(float3 wo, float ax, float ay, float3 up) {
    float3 wi = normalize(f3(ax * wo.x, ay * wo.y, wo.z));
    float z = fmaf(1.0f - up.z, 1.0f + wi.z, -wi.z);
    float st = sqrtf(clampf(1.0f - z * z, 0.0f, 1.0f));
    float3 wm = f3(st * up.x, st * up.y, z) + wi;
    float3 w = f3(wm.x * ax, wm.y * ay, fmaxf(0.0f, wm.z));
    return w / length(w);
}
RPTR_HD float gtr_2_vndf_pdf(float ndo, float cos_h, float alpha) {
    return gtr_2(cos_h, alpha) * (0.5f / smith_den1(ndo, alpha * alpha));
}
RPTR_HD float3 diffuse_basecolor(const GltfMat &m) { return m.base_color * (1.0f - m.metallic); }
RPTR_HD float3 specular_basecolor(const GltfMat &m, float ior) {
    float d = pow2((ior - 1.0f) / (ior + 1.0f));
    return mix3(f3(d), m.base_color, m.metallic);
}
RPTR_HD float specular_alpha(const GltfMat &m) { return fmaxf(m.roughness * m.roughness, 0.002f); }
RPTR_HD float transmission_alpha(const GltfMat &m) { return fmaxf(m.transmission_roughness * m.transmission_roughness, 0.002f); }
RPTR_HD float gltf_schlick_weight(float odh, float ior) {
    float f = schlick_weight(odh);
    if (ior < 1.0f) {
        float cc = sqrtf(1.0f - ior * ior);
        f = mixf(f, 1.0f, fminf((1.0f - odh) / (1.0f - cc), 1.0f));
    }
    return f;
}

RPTR_HD float3 gltf_bsdf(const GltfMat &m, float3 n, float3 wo, float3 wi, bool tr) {
    float idn = dot(n, wi), odn = dot(n, wo);
    float ior = odn < 0.0f ? 1.0f / m.ior : m.ior;
    float3 wh;
    if (idn * odn < 0.0f) {
        if (!tr) return f3(0.0f);
        if (!(m.specular_transmission > 0.0f)) return f3(0.0f);
        if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) wh = wi * (-ior) - wo;
        else wh = reflect3(wi, n) + wo;
        if (!(dot(wh, n) > 0.0f)) return f3(0.0f);
    } else
        wh = wi + wo;
    wh = normalize(wh);
    float odh = dot(wo, wh), idh = dot(wi, wh);
    float3 diffuse = diffuse_basecolor(m) * RPTR_INV_PI;
    float3 specular = f3(0.0f);
    if (m.ior > 1.0f) {
        float3 f0 = specular_basecolor(m, m.ior);
        float sa = specular_alpha(m);
        if (tr && idn * odn < 0.0f) sa = transmission_alpha(m);
        float refl = gtr_2(dot(n, wh), sa);
        refl *= smith_visibility_ggx(odn, idn, sa);
        float fw = gltf_schlick_weight(fabsf(odh), ior);
        float3 F = mix3(f0, f3(1.0f), fw);
        if (tr && idn * odn < 0.0f) {
            diffuse = f3(0.0f);
            specular = m.transmission_color * (refl * (1.0f - m.metallic) * m.specular_transmission) * (f3(1.0f) - F);
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                float ac = 2.0f * odh / (idh * ior + odh);
                specular = specular * (ac * ac);
            }
        } else {
            if (tr) diffuse = diffuse * (1.0f - m.specular_transmission);
            diffuse = diffuse * (f3(1.0f) - F);
            specular = F * refl;
        }
    }
    return diffuse + specular;
}

struct Components { float w0, w1, w2; };
RPTR_HD Components component_sampler(const GltfMat &m, float ior, float3 odh, float3 vis, bool tr) {
    Components c;
    float sl = luminance(specular_basecolor(m, m.ior));
    float F0 = mixf(sl, 1.0f, gltf_schlick_weight(odh.x, 1.0f));
    float F1 = mixf(sl, 1.0f, gltf_schlick_weight(odh.y, 1.0f));
    c.w0 = (1.0f - F0) * vis.x * (1.0f - m.metallic) * luminance(diffuse_basecolor(m));
    c.w1 = F1 * vis.y;
    c.w2 = 0.0f;
    float sum = 0.0f;
    if (tr) {
        float F2 = mixf(sl, 1.0f, gltf_schlick_weight(odh.z, ior));
        c.w0 *= (1.0f - m.specular_transmission);
        c.w2 = (1.0f - F2) * vis.z * (1.0f - m.metallic) * m.specular_transmission;
        sum += c.w0; sum += c.w1; sum += c.w2;
    } else {
        sum += c.w0; sum += c.w1;
    }
    if (sum > 0.0f) {
        c.w0 /= sum;
        c.w1 /= sum;
        if (tr) c.w2 /= sum;
    } else
        c.w0 = 1.0f;
    return c;
}
RPTR_HD int sample_reuse_component(const Components &c, float &rnd, float &prob, bool tr) {
    int comp = 0;
    float next_base = 0.0f, base = 0.0f;
    float w[3] = {c.w0, c.w1, c.w2};
    int n = tr ? 3 : 2;
    for (int i = 0; i < n; ++i) {
        float p = w[i];
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

RPTR_HD float gltf_wpdf(const GltfMat &m, float3 n, float3 wo, float3 wi, bool tr) {
    float idn = dot(n, wi), odn = dot(n, wo);
    float ior = odn < 0.0f ? 1.0f / m.ior : m.ior;
    float pdf = RPTR_INV_PI * fabsf(idn);
    if (m.ior > 1.0f) {
        float3 wh;
        if (idn * odn < 0.0f) {
            if (!tr) return 0.0f;
            if (!(m.specular_transmission > 0.0f)) return 0.0f;
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) wh = wi * (-ior) - wo;
            else wh = reflect3(wi, n) + wo;
            if (!(dot(wh, n) > 0.0f)) return 0.0f;
        } else
            wh = wi + wo;
        wh = normalize(wh);
        float odh = dot(wo, wh), idh = dot(wi, wh);
        float cth = dot(wh, n);
        float3 vis = f3(1.0f, 0.0f, 0.0f);
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
        Components c = component_sampler(m, ior, f3(fabsf(odh)), vis, tr);
        if (tr && idn * odn < 0.0f) sa = ta;
        float spec = gtr_2_vndf_pdf(odn, cth, sa);
        if (tr && idn * odn < 0.0f) {
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                float ac = 2.0f * odh / (idh * ior + odh);
                spec *= ac * ac;
            }
            pdf = spec * c.w2;
        } else {
            pdf *= c.w0;
            pdf += spec * c.w1;
        }
    }
    return pdf;
}

// returns f*|cos|/pdf; pdf == 0 marks failure (mis_wpdf = 0 then)
RPTR_HD float3 sample_gltf_brdf(const GltfMat &m, float3 n, float3 wo, float3 &wi, float &pdf, float &mis_wpdf, float2 rng_sample,
                               float2 fresnel_sample, float3 vx, float3 vy, bool tr) {
    float3 wol = f3(dot(vx, wo), dot(vy, wo), dot(n, wo));
    float odn = wol.z;
    float ior = m.ior;
    mis_wpdf = 0.0f;
    wi = f3(0.0f);
    if (tr) {
        ior = odn < 0.0f ? 1.0f / m.ior : m.ior;
        if (odn < 0.0f) wol.z = -wol.z;
    } else if (odn < 0.0f) {
        pdf = 0.0f;
        return f3(0.0f);
    }
    float3 up = to_pipe_sample(rng_sample);
    float3 wid = normalize(n + sample_sphere(up));
    if (tr && odn < 0.0f) wid = -wid;

    float sa = specular_alpha(m);
    int comp = 0;
    float comp_pdf = 0.0f;
    Components c;
    c.w0 = c.w1 = c.w2 = 0.0f;
    float3 whs = f3(0.0f), wht = f3(0.0f);
    if (m.ior > 1.0f) {
        float3 odh_all = f3(0.0f), vis_all = f3(0.0f);
        odh_all.x = cos_half_angle(dot(wo, wid));
        vis_all.x = 1.0f;
        whs = sample_gtr_2_vndf(wol, sa, sa, up);
        odh_all.y = dot(wol, whs);
        float sidn = reflect3(-wol, whs).z;
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
                if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) tidn = -refract3(-wol, wht, 1.0f / ior).z;
                else tidn = reflect3(-wol, wht).z;
                vis_all.z = tidn > 0.0f ? 2.0f * tidn / smith_den1(tidn, ta * ta) : 0.0f;
            }
        }
        c = component_sampler(m, ior, odh_all, vis_all, tr);
        comp = sample_reuse_component(c, fresnel_sample.x, comp_pdf, tr);
    }
    float cth, idh, odh;
    if (comp == 0) {
        wi = wid;
        float3 wh = normalize(wi + wo);
        cth = dot(n, wh);
        idh = odh = dot(wo, wh);
    } else {
        if (tr && comp == 2) {
            sa = transmission_alpha(m);
            whs = wht;
        }
        float3 wh = whs;
        if (tr && odn < 0.0f) wh.z = -wh.z;
        cth = wh.z;
        wh = mat_mul(vx, vy, n, wh);
        idh = odh = dot(wo, wh);
        if (tr && comp != 1) {
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                wi = refract3(-wo, wh, 1.0f / ior);
                idh = dot(wi, wh);
            } else
                wi = reflect3(reflect3(-wo, wh), n);
        } else
            wi = reflect3(-wo, wh);
    }
    float idn = dot(n, wi);
    bool bad = tr ? ((idn * odn > 0.0f) != (comp != 2)) : !(idn > 0.0f);
    if (bad) {
        pdf = 0.0f;
        return f3(0.0f);
    }
    pdf = RPTR_INV_PI * fabsf(idn);
    if (m.ior > 1.0f) {
        pdf *= c.w0;
        float spec = gtr_2_vndf_pdf(odn, cth, sa);
        if (tr && idn * odn < 0.0f) {
            if (m.flags & RPTR_BASE_MATERIAL_ONESIDED) {
                float ac = 2.0f * odh / (idh * ior + odh);
                spec *= ac * ac;
            }
            pdf = spec * c.w2;
        } else
            pdf += spec * c.w1;
    }
    if (!(pdf > 0.0f)) return f3(0.0f);
    float3 f = gltf_bsdf(m, n, wo, wi, tr);
    mis_wpdf = gltf_wpdf(m, n, wo, wi, tr);
    return f * fabsf(idn) / pdf;
}

// ---- triangle lights (rendering/lights/tri.glsl:58-152) ------------------------------------------------------------------
RPTR_HD float fast_positive_atan(float y) {
    float ay = fabsf(y);
    float rx = (ay > 1.0f) ? (1.0f / ay) : ay;
    float ry = rx * rx;
    float rz = fmaf(ry, 0.02083509974181652f, -0.08513300120830536f);
    rz = fmaf(ry, rz, 0.18014100193977356f);
    rz = fmaf(ry, rz, -0.3302994966506958f);
    ry = fmaf(ry, rz, 0.9998660087585449f);
    rz = fmaf(-2.0f * ry, rx, 0.5f * RPTR_PI);
    rz = (ay > 1.0f) ? rz : 0.0f;
    rx = fmaf(rx, ry, rz);
    return (y < 0.0f) ? (RPTR_PI - rx) : rx;
}
RPTR_HD float half_tri_solid_angle_tan(float3 v0, float3 v1, float3 v2, float3 &params) {
    float hs = (v0.x > 0.0f) ? -1.0f : 1.0f;
    float hk = 1.0f / (fabsf(v0.x) + 1.0f);
    float hy = v0.y * hk, hz = v0.z * hk;
    float d01 = dot(v0, v1), d02 = dot(v1, v2), d12 = dot(v0, v2);
    float dh0 = fmaf(-hs, v1.x, d01);
    float dh2 = fmaf(-hs, v2.x, d12);
    float c0x = fmaf(-dh0, hy, v1.y), c0y = fmaf(-dh0, hz, v1.z);
    float c1x = fmaf(-dh2, hy, v2.y), c1y = fmaf(-dh2, hz, v2.z);
    float det = c0x * c1y - c1x * c0y;
    float vol = fabsf(det);
    float d02p12 = d02 + d12;
    float opd01 = 1.0f + d01;
    params = f3(vol, d02p12, opd01);
    return vol / (opd01 + d02p12);
}
RPTR_HD float triangle_solid_angle(float3 v0, float3 v1, float3 v2, float3 &params) {
    return 2.0f * fast_positive_atan(half_tri_solid_angle_tan(v0, v1, v2, params));
}
RPTR_HD float3 sample_solid_angle_polygon(float3 v0, float3 v1, float3 v2, float omega, float3 prm, float2 rnd) {
    float target = omega * rnd.x;
    float3 a0 = v1, a1 = v0, a2 = v2;
    float s, c;
    sincos_pos(0.5f * target, s, c);
    float3 offset = a0 * (prm.x * c - prm.y * s) + a2 * (prm.z * s);
    float k = 2.0f * (dot(a0, offset) / dot(offset, offset));
    float3 nv2 = f3(fmaf(k, offset.x, -a0.x), fmaf(k, offset.y, -a0.y), fmaf(k, offset.z, -a0.z));
    float s2 = dot(a1, nv2);
    float sm = mix_fma(1.0f, s2, rnd.y);
    float den = fmaf(-s2, s2, 1.0f);
    float tn = sqrtf(fmaf(-sm, sm, 1.0f) / den);
    tn = (den > 0.0f) ? tn : rnd.y;
    return a1 * fmaf(-tn, s2, sm) + nv2 * tn;
}

RPTR_HD float3 ld3(const float *p) { return f3(p[0], p[1], p[2]); }

// sample_tri_lights (rendering/mc/lights_linear.glsl:19-127), BINNED_LIGHTS_BIN_MAX_SIZE = 16
RPTR_HD float3 sample_tri_lights(const FrameParams &fp, const rptr_tri_light_data *lights, float3 hit_p, float3 hit_n, float2 dir_sample,
                                float2 sel, float3 &light_dir, float &light_dist, float &pdf, float &mis_wpdf) {
    int num_lights = fp.n_lights, num_bins = fp.n_bins, bin_size = fp.bin_size;
    sel.x *= (float)num_bins;
    int bin_id = (int)(uint32_t)sel.x;
    bin_id = bin_id < num_bins - 1 ? bin_id : num_bins - 1;
    float sel_p = 1.0f / (float)num_bins;
    float contribs[RPTR_BINNED_LIGHTS_BIN_MAX_SIZE];
    float total = 0.0f;
    const float MIN_IRRADIANCE = 6.2e-4f * 0.001f;
    int bin_end = bin_size * (bin_id + 1);
    bin_end = bin_end < num_lights ? bin_end : num_lights;
#pragma unroll 1
    for (int i = 0; i < RPTR_BINNED_LIGHTS_BIN_MAX_SIZE; ++i) {
        int light_id = bin_size * bin_id + i;
        if (!(light_id < bin_end)) break;
        const rptr_tri_light_data &L = lights[light_id];
        float3 a = ld3(L.v0) - hit_p, b = ld3(L.v1) - hit_p, c = ld3(L.v2) - hit_p;
        bool front = dot(cross(a, b), c) < 0.0f;
        float contrib = luminance(ld3(L.radiance));
        if ((dot(a, hit_n) > 0.0f || dot(b, hit_n) > 0.0f || dot(c, hit_n) > 0.0f) && front) {
            float3 prm;
            contrib *= triangle_solid_angle(normalize(a), normalize(b), normalize(c), prm);
        } else
            contrib = 0.0f;
        contrib += MIN_IRRADIANCE;
        contribs[i] = contrib;
        total += contrib;
    }
    float p = 0.0f, t = 0.0f;
    int light_id = 0;
#pragma unroll 1
    for (int i = 0; i < RPTR_BINNED_LIGHTS_BIN_MAX_SIZE; ++i) {
        light_id = bin_size * bin_id + i;
        if (!(light_id < bin_end)) break;
        p = contribs[i] / total;
        t += p;
        if (sel.y < t) break;
    }
    light_id = light_id < num_lights - 1 ? light_id : num_lights - 1;
    sel_p *= p;
    const rptr_tri_light_data &L = lights[light_id];
    float3 v0 = ld3(L.v0), v1 = ld3(L.v1), v2 = ld3(L.v2);
    float3 d0 = normalize(v0 - hit_p), d1 = normalize(v1 - hit_p), d2 = normalize(v2 - hit_p);
    float3 prm;
    float omega = triangle_solid_angle(d0, d1, d2, prm);
    light_dir = sample_solid_angle_polygon(d0, d1, d2, omega, prm, dir_sample);
    pdf = 1.0f / omega;
    float3 e_n = cross(v1 - v0, v2 - v0);
    light_dist = dot(v0 - hit_p, e_n) / dot(light_dir, e_n);
    mis_wpdf = 2.0f * light_dist * light_dist / fabsf(dot(light_dir, e_n));
    pdf *= sel_p;
    mis_wpdf /= (float)num_bins;
    return ld3(L.radiance) / pdf;
}

// ---- sky (rendering/lights/sky_model_arhosek/sky_model.glsl:40-59, vulkan/pt_megakernel.glsl:113-149) -------------------
RPTR_HD float3 skymodel_radiance(const rptr_scene_params &sp, float3 sun_dir, float3 view) {
    float ct = clampf(view.y, 0.0f, 1.0f);
    float cg = clampf(dot(view, sun_dir), -1.0f, 1.0f);
    float gamma = acos_f(ct);
    float rayM = cg * cg;
    float zenith = sqrtf(ct);
    float out[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float c0 = sp.sky_configs[0][ch], c1 = sp.sky_configs[1][ch], c2 = sp.sky_configs[2][ch], c3 = sp.sky_configs[3][ch],
                    c4 = sp.sky_configs[4][ch], c5 = sp.sky_configs[5][ch], c6 = sp.sky_configs[6][ch], c7 = sp.sky_configs[7][ch],
                    c8 = sp.sky_configs[8][ch];
        float expM = exp_f(c4 * gamma);
        float b = 1.0f + c8 * c8 - 2.0f * c8 * cg;
        float mieM = (1.0f + cg * cg) / (b * sqrtf(b));
        float lhs = 1.0f + c0 * exp_f(c1 / (ct + 0.01f));
        float rhs = c2 + c3 * expM + c5 * rayM + c6 * mieM + c7 * zenith;
        out[ch] = lhs * rhs * sp.sky_radiances[ch] * 0.01f;
    }
    return f3(out[0], out[1], out[2]);
}
RPTR_HD float nee_mis_heuristic(float nf, float pf, float ng, float pg) {
    float f = nf * pf, g = ng * pg;
    return f / (f + g);
}
RPTR_HD float3 compute_sky_illum(const rptr_scene_params &sp, float3 ray_dir, float prev_bsdf_pdf) {
    float3 dir = ray_dir;
    float ocean = 1.0f;
    if (dir.y <= 0.0f) {
        dir.y = -dir.y;
        float b = fmaxf(1.0f - fabsf(dir.y), 0.0f);
        float b2 = b * b;
        ocean = 0.7f * (b2 * b2 * b);
    }
    float3 sun_dir = ld3(sp.sun_dir);
    float3 atm = max0(skymodel_radiance(sp, sun_dir, dir)) * ocean;
    float3 sun = f3(0.0f);
    if (dot(dir, sun_dir) >= sp.sun_cos_angle) sun = ld3(sp.sun_radiance) * ocean;
    float3 illum = f3(0.0f);
    illum = illum + abs3(atm);
    float light_pdf = sp.sun_radiance[3] * (1.0f / (RPTR_TWO_PI * (1.0f - sp.sun_cos_angle)));
    float w = nee_mis_heuristic(1.0f, prev_bsdf_pdf, 1.0f, light_pdf);
    illum = illum + abs3(sun) * w;
    return illum;
}

// ---- quantisation (librender/dequantize.glsl:8-48) -----------------------------------------------------------------------
RPTR_HD float3 dequantize_position(uint64_t q, const float *scale, const float *offset) {
    float ux = (float)(uint32_t)(q & 0x1FFFFFu);
    float uy = (float)(uint32_t)((q >> 21) & 0x1FFFFFu);
    float uz = (float)(uint32_t)((q >> 42) & 0x1FFFFFu);
    return f3(ux * scale[0] + offset[0], uy * scale[1] + offset[1], uz * scale[2] + offset[2]);
}
RPTR_HD float3 dequantize_normal(uint32_t w) {
    float nx = (float)((int)(w & 0xFFFFu) - 0x8000) / 32767.0f;
    float ny = (float)((int)(w >> 16) - 0x8000) / 32767.0f;
    float nl1 = fabsf(nx) + fabsf(ny);
    if (nl1 >= 1.0f) {
        float tx = (1.0f - fabsf(ny)) * (nx >= 0.0f ? 1.0f : -1.0f);
        float ty = (1.0f - fabsf(nx)) * (ny >= 0.0f ? 1.0f : -1.0f);
        nx = tx;
        ny = ty;
    }
    return normalize(f3(nx, ny, 1.0f - nl1));
}
RPTR_HD float2 dequantize_uv(uint32_t w) {
    float s = 8.0f / 65535.0f;
    return f2(0.0f + (float)(int)(w & 0xFFFFu) * s, 1.0f + (float)(-(int)(w >> 16)) * s);
}

// ---- hit attributes (rendering/rt/hit.glsl:49-128,162-203) -----------------------------------------------------------------
RPTR_HD int calc_hit_material_id(const GeomInst &g, uint32_t prim);
// the uv part of calc_hit_attributes (hit.glsl:75-84), also evaluated for alpha candidates during traversal (generate_candidate_hit)
RPTR_HD float2 hit_uv(const GeomInst &g, uint32_t prim, float ax, float ay) {
    if (!g.has_uvs) return f2(0.0f, 0.0f);
    const uint64_t *qn = g.qnuv + 3 * (size_t)prim;
    const float2 uva = dequantize_uv((uint32_t)(qn[0] >> 32)), uvb = dequantize_uv((uint32_t)(qn[1] >> 32)), uvc = dequantize_uv((uint32_t)(qn[2] >> 32));
    const float bx = 1.0f - ax - ay;
    return f2(fmaf(uvc.x, ay, fmaf(uvb.x, ax, uva.x * bx)), fmaf(uvc.y, ay, fmaf(uvb.y, ax, uva.y * bx)));
}
// get_material_alpha (material_textures.glsl:137-145) of a traversal candidate: the per-material texel packed into the triangle
// record, or -- RPTR_TRI_TEXTURED_ALPHA -- the alpha channel of the material's base-colour texture at the candidate's uv
RPTR_HD float candidate_alpha(const SceneDev &sc, int32_t gi_alpha, int32_t prim, float u, float v) {
    if (!(gi_alpha & RPTR_TRI_TEXTURED_ALPHA)) return alpha8_to_float((int32_t)(((uint32_t)gi_alpha) >> 24));
    const GeomInst &g = sc.ginst[gi_alpha & 0x007fffff];
    const rptr_base_material &m = sc.materials[calc_hit_material_id(g, (uint32_t)prim)];
    return sample_texture(sc, RPTR_GET_TEXTURE_ID(f2u(m.base_color[0])), hit_uv(g, (uint32_t)prim, u, v)).w;
}
struct RTHit {
    float3 normal; float dist; float3 geo_normal; int material_id; float3 tangent; float bitangent_l; float2 uv;
};
RPTR_HD int calc_hit_material_id(const GeomInst &g, uint32_t prim) {
    if (g.material_id < 0) return (int)g.tri_mat[prim] - g.material_id - 1;
    return g.material_id;
}
RPTR_HD RTHit calc_hit_attributes(const GeomInst &g, float ray_t, uint32_t prim, float ax, float ay) {
    RTHit h;
    h.dist = ray_t;
    const uint64_t *qv = g.qverts + 3 * (size_t)prim;
    float3 p0 = dequantize_position(qv[0], g.scale, g.offset);
    float3 p1 = dequantize_position(qv[1], g.scale, g.offset);
    float3 p2 = dequantize_position(qv[2], g.scale, g.offset);
    float3 gn = cross(p1 - p0, p2 - p0);
    float3 bary = f3(1.0f - ax - ay, ax, ay);
    float3 n = gn;
    uint64_t qa = 0, qb = 0, qc = 0;
    if (g.has_normals || g.has_uvs) {
        const uint64_t *qn = g.qnuv + 3 * (size_t)prim;
        qa = qn[0]; qb = qn[1]; qc = qn[2];
    }
    if (g.has_normals) {
        n = mat_mul(dequantize_normal((uint32_t)qa), dequantize_normal((uint32_t)qb), dequantize_normal((uint32_t)qc), bary);
        if (dot(n, gn) < 0.0f) gn = -gn;
    }
    h.geo_normal = gn * 0.5f;
    h.normal = n;
    float2 uva = f2(0, 0), uvb = f2(0, 0), uvc = f2(0, 0);
    h.uv = f2(0.0f, 0.0f);
    if (g.has_uvs) {
        uva = dequantize_uv((uint32_t)(qa >> 32));
        uvb = dequantize_uv((uint32_t)(qb >> 32));
        uvc = dequantize_uv((uint32_t)(qc >> 32));
        h.uv = f2(fmaf(uvc.x, bary.z, fmaf(uvb.x, bary.y, uva.x * bary.x)), fmaf(uvc.y, bary.z, fmaf(uvb.y, bary.y, uva.y * bary.x)));
    }
    h.material_id = calc_hit_material_id(g, prim);
    float3 r0 = f3(g.w2o[0], g.w2o[1], g.w2o[2]), r1 = f3(g.w2o[3], g.w2o[4], g.w2o[5]), r2 = f3(g.w2o[6], g.w2o[7], g.w2o[8]);
    h.geo_normal = mat_mul(r0, r1, r2, h.geo_normal);
    h.normal = normalize(mat_mul(r0, r1, r2, h.normal));
    bool requires_tangent = true;
    if (g.has_uvs) {
        float det = length(gn);
        float3 frame_n = gn / (det * det);
        float3 dp2perp = cross(p2 - p0, frame_n);
        float3 dp1perp = cross(frame_n, p1 - p0);
        float2 duv1 = f2(uvb.x - uva.x, uvb.y - uva.y), duv2 = f2(uvc.x - uva.x, uvc.y - uva.y);
        float3 T = dp2perp * duv1.x + dp1perp * duv2.x;
        float3 B = dp2perp * duv1.y + dp1perp * duv2.y;
        T = mat_mul(r0, r1, r2, T);
        B = mat_mul(r0, r1, r2, B);
        float Tlen = length(T);
        if (Tlen > 0.0f && Tlen <= 3.402823466e+38f) { // > 0, not inf, not nan
            h.tangent = T;
            h.bitangent_l = dot(normalize(cross(h.geo_normal, T)), B);
            requires_tangent = false;
        }
    }
    if (requires_tangent) {
        h.tangent = normalize(mat_mul(r0, r1, r2, cross(p2 - p0, gn)));
        h.bitangent_l = 1.0f;
    }
    return h;
}

RPTR_HD float geometry_scale_to_tmin(float3 orig, float scale) { return (length(orig) + scale) * 0.000005f; }

// ---- one path vertex ---------------------------------------------------------------------------------------------------------
// ---- ray footprints for texture level of detail (USE_MIPMAPPING: rendering/rt/footprint.glsl, pt_megakernel.glsl:336-350, 583-605, 698-702) ----
// A GLSL mat2 F, stored as the reference holds it (all four entries): mCR = F[C][R] = column C, row R.  Matrix products follow the
// GLSL definition (A * B)[c][r] = sum_k A[k][r] * B[c][k], each sum as the RPTR-FP dot product (fma chain, last term innermost).
struct Footprint { float m00, m01, m10, m11; };
RPTR_HD float dot2f(float ax, float ay, float bx, float by) { return fmaf(ay, by, ax * bx); }
// dpdxy_to_footprint (footprint.glsl:10-15)
RPTR_HD Footprint dpdxy_to_footprint(float3 ray_dir, float3 dpdx, float3 dpdy) {
    float3 t, b;
    ortho_basis(t, b, ray_dir);
    // M = transpose(mat2x3(t, b)) * mat2x3(dpdx, dpdy): M[0] = (t . dpdx, b . dpdx), M[1] = (t . dpdy, b . dpdy)
    const float m00 = dot(t, dpdx), m01 = dot(b, dpdx), m10 = dot(t, dpdy), m11 = dot(b, dpdy);
    Footprint f; // F = M * transpose(M): F[c][r] = M[0][r] * M[0][c] + M[1][r] * M[1][c]
    f.m00 = dot2f(m00, m10, m00, m10);
    f.m01 = dot2f(m01, m11, m00, m10);
    f.m10 = dot2f(m00, m10, m01, m11);
    f.m11 = dot2f(m01, m11, m01, m11);
    return f;
}
// transform_footprint(dst_ray_dir, T, src_ray_dir, F) with T given by its columns (footprint.glsl:28-35)
RPTR_HD Footprint transform_footprint(float3 dst_ray_dir, float3 tc0, float3 tc1, float3 tc2, float3 src_ray_dir, Footprint F) {
    float3 t, b;
    ortho_basis(t, b, src_ray_dir);
    const float3 u0 = mat_mul(tc0, tc1, tc2, t), u1 = mat_mul(tc0, tc1, tc2, b); // T2 = T * mat2x3(t, b)
    ortho_basis(t, b, dst_ray_dir);
    const float a00 = dot(t, u0), a01 = dot(b, u0), a10 = dot(t, u1), a11 = dot(b, u1); // T3 = transpose(mat2x3(t, b)) * T2
    // G = T3 * F: G[c][r] = T3[0][r] * F[c][0] + T3[1][r] * F[c][1]
    const float g00 = dot2f(a00, a10, F.m00, F.m01), g01 = dot2f(a01, a11, F.m00, F.m01);
    const float g10 = dot2f(a00, a10, F.m10, F.m11), g11 = dot2f(a01, a11, F.m10, F.m11);
    Footprint h; // H = G * transpose(T3): H[c][r] = G[0][r] * T3[0][c] + G[1][r] * T3[1][c]
    h.m00 = dot2f(g00, g10, a00, a10);
    h.m01 = dot2f(g01, g11, a00, a10);
    h.m10 = dot2f(g00, g10, a01, a11);
    h.m11 = dot2f(g01, g11, a01, a11);
    return h;
}
// reflect_footprint (footprint.glsl:38-42): R = mat3(1) - 2 * outerProduct(n, n), n = normalize(dst - src)
RPTR_HD Footprint reflect_footprint(float3 dst_ray_dir, float3 src_ray_dir, Footprint F) {
    const float3 n = normalize(dst_ray_dir - src_ray_dir);
    const float3 c0 = f3(1.0f - 2.0f * (n.x * n.x), 0.0f - 2.0f * (n.y * n.x), 0.0f - 2.0f * (n.z * n.x));
    const float3 c1 = f3(0.0f - 2.0f * (n.x * n.y), 1.0f - 2.0f * (n.y * n.y), 0.0f - 2.0f * (n.z * n.y));
    const float3 c2 = f3(0.0f - 2.0f * (n.x * n.z), 0.0f - 2.0f * (n.y * n.z), 1.0f - 2.0f * (n.z * n.z));
    return transform_footprint(dst_ray_dir, c0, c1, c2, src_ray_dir, F);
}
// footprint_to_dpdxy (footprint.glsl:44-61): the principal axes of F as world-space differentials
RPTR_HD void footprint_to_dpdxy(float3 &dpdx, float3 &dpdy, float3 ray_dir, Footprint F) {
    const float B = F.m00 + F.m11;
    const float C = F.m00 * F.m11 - F.m01 * F.m10;
    const float D = sqrtf(B * B * 0.25f - C);
    const float ev0 = 0.5f * B - D, ev1 = 0.5f * B + D;
    float x0x = 1.0f, x0y = 0.0f, x1x = 0.0f, x1y = 1.0f; // X = mat2(1)
    if (fabsf(F.m01) > 3.0e-39f) {
        x0x = F.m10; x0y = ev0 - F.m00;
        x1x = ev1 - F.m11; x1y = F.m01;
    }
    float3 t, b;
    ortho_basis(t, b, ray_dir);
    const float i0 = 1.0f / sqrtf(dot2f(x0x, x0y, x0x, x0y)), i1 = 1.0f / sqrtf(dot2f(x1x, x1y, x1x, x1y));
    const float n0x = x0x * i0, n0y = x0y * i0, n1x = x1x * i1, n1y = x1y * i1;
    const float s0 = sqrtf(ev0), s1 = sqrtf(ev1);
    dpdx = f3(fmaf(b.x, n0y, t.x * n0x), fmaf(b.y, n0y, t.y * n0x), fmaf(b.z, n0y, t.z * n0x)) * s0;
    dpdy = f3(fmaf(b.x, n1y, t.x * n1x), fmaf(b.y, n1y, t.y * n1x), fmaf(b.z, n1y, t.z * n1x)) * s1;
}

struct PathState {
    float3 o, d;
    float tmin, tmax;
    float3 thr;
    float prev_pdf;
    float3 illum;
    float total_t;
    uint32_t rng;     // RANDOM_STATE word that advances with every draw (LCG state; Sobol scramble LCG; BN sampleID)
    int bounce;
    uint32_t rng_b;   // second word, constant along the path (Sobol index / BN pixelID; unused by the LCG)
    int32_t rng_dim;  // RANDOM_SET_DIM / RANDOM_SHIFT_DIM cursor
    Footprint foot;   // texture_footprint (pt_megakernel.glsl:338-350, 700): read by the texture lookups of textured scenes only
};
struct ShadowRay {
    float3 o, d;
    float tmin, tmax;
    float3 contrib; // throughput * L * w * |cos| * f, added to illum iff unoccluded
};
enum ShadeResult { SHADE_TERMINATE = 0, SHADE_CONTINUE = 1 };
// First-vertex attributes the megakernel stores into its fp16 AOV images (vulkan/accumulate.glsl:89-103;
// pt_megakernel.glsl:482-486, 670-672; shade_base_material.glsl:28-31), as floats before the fp16 store
struct AovSample {
    float3 albedo;  // path_throughput * base_color (emitters: 0)
    float roughness; // ior != 1 ? roughness : 1
    float3 normal;  // interaction.n
    float depth;    // |p - cam_pos|
    float motion[2]; // screen-space motion of the vertex against the reference view (accumulate.glsl:77-87)
    float jitter[2]; // view_params.screen_jitter
};
// mat4 * vec4(p, 1) in RPTR-FP order: columns accumulated left to right
RPTR_HD void project_point(const float *m, float3 p, float &x, float &y, float &w) {
    x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * 1.0f;
    y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * 1.0f;
    w = ((m[3] * p.x + m[7] * p.y) + m[11] * p.z) + m[15] * 1.0f;
}
// store_motion_jitter_aovs (vulkan/accumulate.glsl:77-87) with motion_vector = 0 (static geometry)
RPTR_HD void aov_motion_jitter(const FrameParams &fp, float3 position, AovSample &a) {
    const float3 moved = position + f3(0.0f);
    float rx, ry, rw, cx, cy, cw;
    project_point(fp.vp_reference, moved, rx, ry, rw);
    project_point(fp.vp, position, cx, cy, cw);
    const float rd = rw < 0.0f ? 0.0f : rw, cd = cw < 0.0f ? 0.0f : cw; // max(w, 0)
    a.motion[0] = rx / rd - cx / cd;
    a.motion[1] = ry / rd - cy / cd;
    a.jitter[0] = fp.screen_jitter[0];
    a.jitter[1] = fp.screen_jitter[1];
}
RPTR_HD AovSample aov_of_miss(const FrameParams &fp) { // store_geometry_aovs(0, vec3(2e32), 0) + store_material_aovs(0, 1, 1)
    AovSample a;
    a.albedo = f3(0.0f); a.roughness = 1.0f; a.normal = f3(0.0f);
    a.depth = length(f3(2.e32f) - ld3(fp.cam_pos));
    aov_motion_jitter(fp, f3(2.e32f), a);
    return a;
}

// FEAT: code paths compiled in (the reference compiles its shader variants from #defines the same way, e.g.
// GLTF_SUPPORT_TRANSMISSION, RBO_rng_variant); a kernel built without a feature must only be launched when the frame
// does not use it.
#define RPTR_FEAT_TRANSMISSION 1 // fp.transmission may be set
#define RPTR_FEAT_TRI_LIGHTS 2   // the scene has binned triangle lights (p_sun < 1)
#define RPTR_FEAT_AOV 4          // fp.output_channel may be non-zero
#define RPTR_FEAT_QMC 8          // fp.rng_variant may select a Sobol / blue-noise sampler
#define RPTR_FEAT_NORMAL_MAPS 16 // some material has a normal map
#define RPTR_FEAT_TEXTURES 32    // some material parameter refers to a texture larger than 1 x 1
#define RPTR_FEAT_ALL 63
// RANDOM_FLOAT1(rng, d) of the selected pointset (rendering/pointsets/selected_rng.glsl, rendering/defaults.glsl:23-28)
template <int FEAT>
RPTR_HD float path_rand(const FrameParams &fp, PathState &ps, int d) {
    if (!(FEAT & RPTR_FEAT_QMC) || fp.rng_variant == 0) return lcg_randomf(ps.rng);
    Sampler sm;
    sm.a = ps.rng; sm.b = ps.rng_b; sm.dim = ps.rng_dim;
    const float r = sampler_next(fp.rng_variant, fp.pts, sm, d);
    ps.rng = sm.a;
    return r;
}

// alpha_rng of a path when the pointset is not the LCG (pt_megakernel.glsl:354-358): the LCG the UNIFORM pointset would start from
RPTR_HD uint32_t alpha_lcg_seed(const FrameParams &fp, int px, int py, uint32_t sample_index) {
    return sampler_init(0, fp.pts, sample_index, fp.first_sample, fp.frame_offset, (uint32_t)px, (uint32_t)py, (uint32_t)fp.width).a;
}

// texture footprint of a pixel at the start of a path (pt_megakernel.glsl:341-351); ray queries run the same block on their own direction
RPTR_HD void init_footprint(const FrameParams &fp, PathState &ps) {
    const float3 dpdx = (ld3(fp.du) / (float)fp.width) * fp.pixel_radius, dpdy = (ld3(fp.dv) / (float)fp.height) * fp.pixel_radius;
    ps.foot = dpdxy_to_footprint(ps.d, dpdx, dpdy);
}
// primary ray + path state (vulkan/pt_megakernel.glsl:310-365)
RPTR_HD void generate_primary(const FrameParams &fp, int px, int py, uint32_t sample_index, PathState &ps) {
    const Sampler sm = sampler_init(fp.rng_variant, fp.pts, sample_index, fp.first_sample, fp.frame_offset, (uint32_t)px, (uint32_t)py,
                                    (uint32_t)fp.width);
    ps.rng = sm.a; ps.rng_b = sm.b; ps.rng_dim = 0;
    float ptx = (float)px + 0.5f, pty = (float)py + 0.5f;
    if (fp.enable_raster_taa == 0) {
        float ux = path_rand<RPTR_FEAT_ALL>(fp, ps, RPTR_DIM_PIXEL_X);
        float uy = path_rand<RPTR_FEAT_ALL>(fp, ps, RPTR_DIM_PIXEL_X + 1);
        ptx += ux - 0.5f;
        pty += uy - 0.5f;
    }
    ptx /= (float)fp.width;
    pty /= (float)fp.height;
    if (fp.enable_raster_taa != 0) { // pt_megakernel.glsl:319-320
        ptx += 0.5f * fp.screen_jitter[0];
        pty += 0.5f * fp.screen_jitter[1];
    }
    ps.o = ld3(fp.cam_pos);
    ps.d = normalize(ld3(fp.du) * ptx + ld3(fp.dv) * pty + ld3(fp.tl));
    ps.tmin = 0.0f;
    ps.tmax = 2.e32f;
    ps.thr = f3(1.0f);
    ps.prev_pdf = 2.e16f;
    ps.illum = f3(0.0f);
    ps.total_t = 0.0f;
    ps.bounce = 0;
    init_footprint(fp, ps);
}

// Shades the vertex found by the closest-hit stage (tri < 0: miss).  On SHADE_CONTINUE ps holds the next ray.
// sh.tmax < 0 means "no shadow ray"; sh.tmax == 0 means "visible without tracing" (the contribution is then
// already added to ps.illum).  Restates pt_megakernel.glsl:480-731 + shade_base_material.glsl:14-96 + nee.glsl:32-90.
// A path that leaves the scene: sky + sun disc, weighted against the sun's NEE pdf (pt_megakernel.glsl:113-149, 480-489).
// The wavefront defers this to the resolve kernel (the miss is always the last event of a path, so the sum is the same).
RPTR_HD float3 shade_miss(const rptr_scene_params &sp, float3 illum, float3 thr, float3 dir, float prev_pdf) {
    return illum + thr * compute_sky_illum(sp, dir, prev_pdf);
}

template <int FEAT = RPTR_FEAT_ALL>
RPTR_HD ShadeResult shade_hit(const FrameParams &fp, const SceneDev &sc, PathState &ps, float hit_t, float hit_u, float hit_v, const Tri *tri,
                             ShadowRay &sh, AovSample *aov = nullptr) {
    sh.tmax = -1.0f;
    const rptr_scene_params &sp = fp.sp;
    const bool tr = (FEAT & RPTR_FEAT_TRANSMISSION) && fp.transmission != 0;
    const int output_channel = (FEAT & RPTR_FEAT_AOV) ? fp.output_channel : 0;
    ps.rng_dim = RPTR_DIM_CAMERA_END + ps.bounce * (RPTR_DIM_VERTEX_END + RPTR_DIM_LIGHT_END); // RANDOM_SET_DIM, pt_megakernel.glsl:423
    const GeomInst &g = sc.ginst[tri_geom_inst(*tri)];
    RTHit h = calc_hit_attributes(g, hit_t, (uint32_t)tri->prim, hit_u, hit_v);
    float approx_sa = length(h.geo_normal);
    h.geo_normal = h.geo_normal / approx_sa;
    approx_sa *= fabsf(dot(h.geo_normal, ps.d)) / (h.dist * h.dist);
    ps.total_t += h.dist;
    float geometry_scale = ps.total_t;
    float3 w_o = -ps.d;
    float3 ip = ps.o + ps.d * h.dist;
    float3 ign = h.geo_normal, in_ = h.normal;
    const rptr_base_material &mp = sc.materials[h.material_id];
    if (dot(w_o, ign) < 0.0f) {
        if (mp.flags & RPTR_BASE_MATERIAL_VOLUME) {
            ip = ps.o;
            h.dist = 0.0f;
        } else if (!(mp.flags & RPTR_BASE_MATERIAL_ONESIDED)) {
            in_ = -in_;
            ign = -ign;
        }
    }
    if ((FEAT & RPTR_FEAT_NORMAL_MAPS) && mp.normal_map != -1) { // pt_megakernel.glsl:634-654 (1x1-texel mode: one texel per material)
        float3 t_y = normalize(cross(h.normal, h.tangent));
        float3 t_x = cross(t_y, h.normal);
        t_x = t_x * length(h.tangent);
        t_y = t_y * h.bitangent_l;
        float4 tx = sc.normal_texels[h.material_id];
        if ((FEAT & RPTR_FEAT_TEXTURES) && tx.w != 0.0f) tx = sample_texture_lod(sc, (uint32_t)mp.normal_map, h.uv, ps.bounce); // textureLod(.., hit.uv, float(bounce)), :641-647
        float3 map_nrm = f3(2.0f * tx.x - 1.0f, 2.0f * tx.y - 1.0f, 1.0f * tx.z - 0.0f);
        map_nrm.z = sqrtf(fmaxf(1.0f - map_nrm.x * map_nrm.x - map_nrm.y * map_nrm.y, 0.0f));
        in_ = normalize(mat_mul(t_x, t_y, in_ * sp.normal_z_scale, map_nrm));
    }
    {
        float nw = dot(w_o, in_), gnw = dot(w_o, ign);
        if (nw * gnw <= 0.0f) {
            float blend = gnw / (gnw - nw);
            in_ = normalize(mix3(ign, in_, blend - 0.0001f));
        }
    }
    float3 v_y = normalize(cross(in_, h.tangent));
    float3 v_x = cross(v_y, in_);

    GltfMat mat;
    float3 emit;
    float2 duvdx = f2(0.0f, 0.0f), duvdy = f2(0.0f, 0.0f);
    if ((FEAT & RPTR_FEAT_TEXTURES) && fp.image_textures) { // hit.duvdxy, pt_megakernel.glsl:583-605 (total_t already includes this segment)
        float3 dpdx, dpdy;
        footprint_to_dpdxy(dpdx, dpdy, ps.d, ps.foot);
        const float3 dir_tangent_un = ps.d - h.geo_normal * dot(ps.d, h.geo_normal);
        const float cos_theta2 = fmaxf(1.0f - dot(dir_tangent_un, dir_tangent_un), 0.0f);
        const float3 dir_tangent_elong = dir_tangent_un / (sqrtf(cos_theta2) + cos_theta2);
        const float3 dpdx_ = dpdx + dir_tangent_elong * dot(dpdx, dir_tangent_un);
        const float3 dpdy_ = dpdy + dir_tangent_elong * dot(dpdy, dir_tangent_un);
        const float3 bitangent = cross(h.geo_normal, normalize(h.tangent)) * h.bitangent_l;
        duvdx = f2(dot(h.tangent, dpdx_) * ps.total_t, dot(bitangent, dpdx_) * ps.total_t);
        duvdy = f2(dot(h.tangent, dpdy_) * ps.total_t, dot(bitangent, dpdy_) * ps.total_t);
    }
    unpack_material<(FEAT & RPTR_FEAT_TEXTURES) != 0>(mat, emit, mp, tr, sc, h.uv, duvdx, duvdy);
    if (aov && ps.bounce == 0) {
        aov->normal = in_;
        aov->depth = length(ip - ld3(fp.cam_pos));
        aov->albedo = ps.thr * mat.base_color;
        aov->roughness = mat.ior != 1.0f ? mat.roughness : 1.0f;
        aov_motion_jitter(fp, ip, *aov);
    }
    const float p_sun = sp.sun_radiance[3];
    if (output_channel == 0 && !is_zero(emit)) {
        float light_pdf = (1.0f - p_sun) * (1.0f / ((float)fp.n_bins * approx_sa));
        float w = nee_mis_heuristic(1.0f, ps.prev_pdf, 1.0f, light_pdf);
        ps.illum = ps.illum + ps.thr * w * emit;
    }
    if (output_channel != 0) {
        float reliability = u2f((uint32_t)(127 - 2 * ps.bounce) << 23); // pow(0.25, bounce)
        if (output_channel == 1) ps.illum = ps.illum + ps.thr * mat.base_color * reliability;
        else if (output_channel == 2) ps.illum = ps.illum + in_ * reliability;
        else if (output_channel == 3) ps.illum = ps.illum + ip * reliability;
    }
    if (ps.bounce + 1 >= fp.max_path_depth) return SHADE_TERMINATE;
    if (output_channel == 0) {
        float2 dir_sample, sel_sample;
        dir_sample.x = path_rand<FEAT>(fp, ps, RPTR_DIM_POSITION_X);
        dir_sample.y = path_rand<FEAT>(fp, ps, RPTR_DIM_POSITION_X + 1);
        sel_sample.x = path_rand<FEAT>(fp, ps, RPTR_DIM_LIGHT_SEL_1);
        sel_sample.y = path_rand<FEAT>(fp, ps, RPTR_DIM_LIGHT_SEL_1 + 1);
        float3 li = f3(0.0f), light_dir = f3(0.0f);
        float light_dist = 2.e16f, light_pdf = 0.0f, mis_pdf = 0.0f;
        if (!(FEAT & RPTR_FEAT_TRI_LIGHTS) || sel_sample.x <= p_sun) {
            sel_sample.x /= p_sun;
            float sn, cs;
            sincos_pos(RPTR_TWO_PI * dir_sample.x, sn, cs);
            float cosT = mixf(1.0f, sp.sun_cos_angle, dir_sample.y);
            float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
            float3 sun_dir = ld3(sp.sun_dir);
            float3 fx, fy;
            ortho_basis(fx, fy, sun_dir);
            light_dir = mat_mul(fx, fy, sun_dir, f3(sinT * cs, sinT * sn, cosT));
            float pdf = 1.0f / (RPTR_TWO_PI * (1.0f - sp.sun_cos_angle));
            li = li + (f3(1.0f) / pdf) * (ld3(sp.sun_radiance) / p_sun);
            light_pdf = pdf * p_sun;
            mis_pdf = light_pdf;
        } else {
            sel_sample.x = (sel_sample.x - p_sun) / (1.0f - p_sun);
            float tri_mis = 0.0f;
            li = li + sample_tri_lights(fp, sc.lights, ip, in_, dir_sample, sel_sample, light_dir, light_dist, light_pdf, tri_mis) / (1.0f - p_sun);
            light_pdf *= 1.0f - p_sun;
            mis_pdf = tri_mis * (1.0f - p_sun);
        }
        if (light_pdf > 0.0f && dot(light_dir, ign) * dot(light_dir, in_) > 0.0f) {
            float bsdf_pdf = gltf_wpdf(mat, in_, w_o, light_dir, tr);
            if (bsdf_pdf >= 0.0f) {
                float3 bsdf = gltf_bsdf(mat, in_, w_o, light_dir, tr);
                float w = nee_mis_heuristic(1.0f, mis_pdf, 1.0f, bsdf_pdf);
                float3 contrib = ps.thr * (li * (bsdf * (w * fabsf(dot(light_dir, in_)))));
                // raytrace_test_visibility (pt_megakernel.glsl:216-272)
                float eps = geometry_scale_to_tmin(ip, geometry_scale);
                if (light_dist - 2.0f * eps > 0.0f) {
                    sh.o = ip;
                    sh.d = light_dir;
                    sh.tmin = eps;
                    sh.tmax = light_dist - eps;
                    sh.contrib = contrib;
                } else {
                    ps.illum = ps.illum + contrib;
                    sh.tmax = 0.0f;
                }
            }
        }
    }
    ps.rng_dim += RPTR_DIM_LIGHT_END; // RANDOM_SHIFT_DIM, shade_base_material.glsl:66
    if (fp.glossy_only_mode != 0 && !(mat.roughness < RPTR_GLOSSY_MODE_ROUGHNESS_THRESHOLD && mat.ior != 1.0f)) return SHADE_TERMINATE;
    float2 lobe, dirs;
    lobe.x = path_rand<FEAT>(fp, ps, RPTR_DIM_LOBE);
    lobe.y = path_rand<FEAT>(fp, ps, RPTR_DIM_LOBE + 1);
    dirs.x = path_rand<FEAT>(fp, ps, RPTR_DIM_DIRECTION_X);
    dirs.y = path_rand<FEAT>(fp, ps, RPTR_DIM_DIRECTION_X + 1);
    float3 w_i;
    float sampling_pdf = 0.0f, mis_wpdf = 0.0f;
    float3 bsdf = sample_gltf_brdf(mat, in_, w_o, w_i, sampling_pdf, mis_wpdf, dirs, lobe, v_x, v_y, tr);
    ps.rng_dim += RPTR_DIM_VERTEX_END; // shade_base_material.glsl:82
    ++ps.bounce;
    if (mis_wpdf == 0.0f || is_zero(bsdf) || !(dot(w_i, in_) * dot(w_i, ign) > 0.0f)) return SHADE_TERMINATE;
    ps.thr = ps.thr * bsdf;
    ps.prev_pdf = mis_wpdf;
    if ((FEAT & RPTR_FEAT_TEXTURES) && fp.image_textures && dot(w_i, in_) * dot(w_o, in_) > -0.999f) ps.foot = reflect_footprint(w_i, ps.d, ps.foot); // :698-702
    ps.d = w_i;
    ps.o = ip;
    ps.tmin = geometry_scale_to_tmin(ps.o, ps.total_t);
    ps.tmax = 1e20f;
    if (ps.bounce >= fp.rr_path_depth) {
        float prefix = fmaxf(ps.thr.x, fmaxf(ps.thr.y, ps.thr.z));
        float rr_prob = prefix;
        float rr_sample = path_rand<FEAT>(fp, ps, RPTR_DIM_RR);
        if (ps.bounce > 6) rr_prob = fminf(0.95f, rr_prob);
        else rr_prob = fminf(1.0f, rr_prob);
        if (rr_sample < rr_prob) ps.thr = ps.thr / rr_prob;
        else return SHADE_TERMINATE;
    }
    return SHADE_CONTINUE;
}

// One path vertex (hit or miss) -- the megakernel-order statement used by the host-side simulation and the unit tests.
RPTR_HD ShadeResult shade_vertex(const FrameParams &fp, const SceneDev &sc, PathState &ps, float hit_t, float hit_u, float hit_v,
                                const Tri *tri, ShadowRay &sh, AovSample *aov = nullptr) {
    sh.tmax = -1.0f;
    if (!tri) {
        if (aov && ps.bounce == 0) *aov = aov_of_miss(fp);
        ps.illum = shade_miss(fp.sp, ps.illum, ps.thr, ps.d, ps.prev_pdf);
        return SHADE_TERMINATE;
    }
    return shade_hit(fp, sc, ps, hit_t, hit_u, hit_v, tri, sh, aov);
}

} // namespace rp
