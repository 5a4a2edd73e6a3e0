// tests/hostsim/hostsim.cpp -- TEST-ONLY harness (never part of librptr_cuda.so).
// Compiles the product's __host__ __device__ shading / traversal code (csrc/rptr_shading.cuh, rptr_bvh.cuh) and its
// host scene ingestion (csrc/rptr_host.cpp) with g++ and runs them one sample at a time on the CPU, so that
// tests/test_hostsim_parity.py can check -- without a GPU -- that the code the kernels execute agrees bit for bit with
// the independent oracle.  The wavefront kernels themselves (queues, atomics, resolve) are covered by the -m gpu tests.
#include "../../realtimepathtracingresearchframework_b200/csrc/rptr_host.hpp"
#include <cmath>
#include <cstring>

using namespace rp;

extern "C" {

struct hostsim_args { // same layout as oracle_render_args
    int32_t width, height;
    rptr_camera_params camera;
    rptr_render_params params;
    rptr_light_sampling_config lighting;
    rptr_scene_params scene_params;
    uint32_t frame_offset, first_sample;
    int32_t n_samples;
    int32_t x0, y0, x1, y1;
    int32_t transmission, n_threads;
    int32_t rng_variant, batch_spp;
    const uint32_t *pointset_tables[4];
    float vp_reference[16];
};

struct hostsim_scene {
    HostScene hs;
    std::vector<GeomInst> gi;
    std::vector<TexDev> tex;
    SceneDev dev() const { // the device-side scene tables, here with host pointers
        return SceneDev{gi.data(), hs.materials.data(), hs.lights.data(), reinterpret_cast<const float4 *>(hs.normal_texels.data()), tex.data(), hs.srgb_lut};
    }
};

hostsim_scene *hostsim_scene_create(const rptr_scene_desc *d, const rptr_light_sampling_config *ls) {
    hostsim_scene *s = new hostsim_scene();
    try {
        build_host_scene(*d, *ls, s->hs);
    } catch (const std::exception &e) {
        fprintf(stderr, "hostsim: %s\n", e.what());
        delete s;
        return nullptr;
    }
    for (const HostGeomInst &h : s->hs.ginst) {
        GeomInst g = h.g;
        g.qverts = s->hs.qverts[h.geometry].data();
        g.qnuv = s->hs.qnuv[h.geometry].empty() ? nullptr : s->hs.qnuv[h.geometry].data();
        g.tri_mat = s->hs.tri_mat[h.pmesh].empty() ? nullptr : s->hs.tri_mat[h.pmesh].data() + h.prim_offset;
        s->gi.push_back(g);
    }
    for (const HostTexture &t : s->hs.textures)
        s->tex.push_back(TexDev{t.rgba.empty() ? nullptr : reinterpret_cast<const uchar4 *>(t.rgba.data()), t.width, t.height, t.srgb, t.levels});
    return s;
}
void hostsim_scene_destroy(hostsim_scene *s) { delete s; }
int32_t hostsim_num_lights(const hostsim_scene *s) { return (int32_t)s->hs.lights.size(); }
void hostsim_get_lights(const hostsim_scene *s, rptr_tri_light_data *out) { memcpy(out, s->hs.lights.data(), s->hs.lights.size() * sizeof(*out)); }
int32_t hostsim_num_nodes(const hostsim_scene *s) { return (int32_t)s->hs.nodes.size(); }
// FNV-1a over the 4-wide nodes and the leaf-ordered triangle records: the BVH must not depend on the builder's thread count
uint64_t hostsim_bvh_hash(const hostsim_scene *s) {
    uint64_t h = 1469598103934665603ull;
    auto eat = [&](const void *p, size_t n) {
        const unsigned char *b = (const unsigned char *)p;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    eat(s->hs.nodes.data(), s->hs.nodes.size() * sizeof(BvhNode));
    eat(s->hs.leaf_tris.data(), s->hs.leaf_tris.size() * sizeof(Tri));
    return h;
}

static FrameParams make_frame(const hostsim_scene *s, const hostsim_args *a) {
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.width = a->width; fp.height = a->height;
    memcpy(fp.cam_pos, a->camera.pos, 12);
    view_params(a->camera, a->width, a->height, fp.du, fp.dv, fp.tl);
    fp.frame_offset = a->frame_offset;
    fp.first_sample = a->first_sample;
    fp.batch = 1;
    fp.max_path_depth = a->params.max_path_depth;
    fp.rr_path_depth = a->params.rr_path_depth;
    fp.output_channel = a->params.output_channel;
    fp.glossy_only_mode = a->params.glossy_only_mode;
    fp.enable_raster_taa = a->params.enable_raster_taa;
    fp.pixel_radius = a->params.pixel_radius;
    fp.image_textures = s->hs.any_textured ? 1 : 0;
    if (fp.enable_raster_taa > 0) screen_jitter(a->frame_offset, a->first_sample, a->width, a->height, fp.screen_jitter);
    fp.n_lights = (int)s->hs.lights.size();
    fp.bin_size = a->lighting.bin_size;
    fp.n_bins = fp.bin_size > 0 ? (fp.n_lights + fp.bin_size - 1) / fp.bin_size : 0;
    fp.transmission = a->transmission;
    fp.rng_variant = a->rng_variant;
    fp.pts.sobol_matrix = a->pointset_tables[0];
    fp.pts.sobol_tile_invert = a->pointset_tables[1];
    fp.pts.bn_sobol = a->pointset_tables[2];
    fp.pts.bn_scrambling = a->pointset_tables[3];
    view_projection(a->camera, a->width, a->height, fp.vp);
    memcpy(fp.vp_reference, a->vp_reference, sizeof(fp.vp_reference));
    fp.sp = a->scene_params;
    if (fp.n_lights > 0) fp.sp.sun_radiance[3] *= 0.5f;
    else fp.sp.sun_radiance[3] = 1.0f;
    return fp;
}

// un-averaged sample layer `sample_index` for the region; rgba is W*H*4; aov (optional) = W*H*8 floats per pixel:
// albedo.rgb, roughness, normal.xyz, depth of the first path vertex (the values behind the fp16 AOV images)
static int render_sample(const hostsim_scene *s, const hostsim_args *a, uint32_t sample_index, float *rgba, float *aov_out, int aov_stride = 8) {
    FrameParams fp = make_frame(s, a);
    const SceneDev sc = s->dev();
    BvhDev bvh{s->hs.nodes.data(), s->hs.leaf_tris.data(), (int32_t)s->hs.nodes.size(), (int32_t)s->hs.leaf_tris.size()};
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = a->y0; y < a->y1; ++y)
        for (int x = a->x0; x < a->x1; ++x) {
            PathState ps;
            generate_primary(fp, x, y, sample_index, ps);
            uint32_t alpha_lcg = alpha_lcg_seed(fp, x, y, sample_index);
            TraceCounters cnt{0, 0};
            AovSample as;
            memset(&as, 0, sizeof(as));
            for (;;) {
                HitRec h;
                // stochastic alpha draws come from the path's LCG (the LCG pointset; the QMC pointsets keep a separate one)
                bool found = closest_hit_filtered(bvh, sc, ps.o, ps.d, ps.tmin, ps.tmax, fp.rng_variant == 0 ? ps.rng : alpha_lcg, h, cnt);
                ShadowRay sh;
                ShadeResult r = shade_vertex(fp, sc, ps, h.t, h.u, h.v, found ? &bvh.tris[h.tri] : nullptr, sh, aov_out ? &as : nullptr);
                if (sh.tmax > 0.0f) {
                    HitRec o;
                    const AlphaFilter af{sc, fp.first_sample, fp.frame_offset, (uint32_t)x + (uint32_t)y * (uint32_t)fp.width};
                    if (!trace_ray<true>(bvh, sh.o, sh.d, sh.tmin, sh.tmax, o, cnt, sh.tmin, 0x7fffffff, &af)) ps.illum = ps.illum + sh.contrib;
                }
                if (r == SHADE_TERMINATE) break;
            }
            float *px = rgba + 4 * ((size_t)y * a->width + x);
            px[0] = ps.illum.x; px[1] = ps.illum.y; px[2] = ps.illum.z; px[3] = ps.bounce == 0 ? 0.0f : 1.0f;
            if (aov_out) {
                const float m[12] = {as.albedo.x, as.albedo.y, as.albedo.z, as.roughness, as.normal.x, as.normal.y, as.normal.z, as.depth,
                                     as.motion[0], as.motion[1], as.jitter[0], as.jitter[1]};
                memcpy(aov_out + (size_t)aov_stride * ((size_t)y * a->width + x), m, sizeof(float) * aov_stride);
            }
        }
    return 0;
}
int hostsim_render_sample(const hostsim_scene *s, const hostsim_args *a, uint32_t sample_index, float *rgba) {
    return render_sample(s, a, sample_index, rgba, nullptr);
}
int hostsim_render_sample_aov(const hostsim_scene *s, const hostsim_args *a, uint32_t sample_index, float *rgba, float *aov) {
    return render_sample(s, a, sample_index, rgba, aov);
}
// aov = 12 floats per pixel: the 8 above + motion.xy, screen_jitter.xy (the motion / jitter image)
int hostsim_render_sample_aov3(const hostsim_scene *s, const hostsim_args *a, uint32_t sample_index, float *rgba, float *aov) {
    return render_sample(s, a, sample_index, rgba, aov, 12);
}
void hostsim_view_params(const rptr_camera_params *cam, int32_t w, int32_t h, float *out) { view_params(*cam, w, h, out, out + 3, out + 6); }
void hostsim_view_projection(const rptr_camera_params *cam, int32_t w, int32_t h, float *out) { view_projection(*cam, w, h, out); }

// sampler calls replayed through the product's rptr_pointsets.cuh (same protocol as oracle_pointset_replay)
int hostsim_pointset_replay(int variant, const uint32_t *const *tables, uint32_t sample_index, uint32_t frame_id, uint32_t frame_offset, uint32_t px,
                            uint32_t py, uint32_t w, const int32_t *ops, const int32_t *args, int n_ops, float *out, uint32_t *state_out) {
    PointsetTables t{tables[0], tables[1], tables[2], tables[3]};
    Sampler s = sampler_init(variant, t, sample_index, frame_id, frame_offset, px, py, w);
    if (state_out) {
        state_out[0] = s.b;
        state_out[1] = s.a;
    }
    int n = 0;
    for (int i = 0; i < n_ops; ++i) {
        if (ops[i] == 0) out[n++] = sampler_next(variant, t, s, args[i]);
        else if (ops[i] == 1) s.dim = args[i];
        else s.dim += args[i];
    }
    return n;
}
// the product's view of material `mid`: resolve_materials (host, set_scene) followed by the device-side unpack_material;
// same output layout as ref_unpack_material ([15] = alpha as unpack returns it for the resolved constants, [16] = alpha texel / 255)
void hostsim_unpack_material(const hostsim_scene *s, int32_t mid, int32_t transmission, float *out) {
    GltfMat m;
    float3 e;
    memset(out, 0, 17 * sizeof(float));
    out[15] = unpack_material(m, e, s->hs.materials[mid], transmission != 0);
    out[16] = alpha8_to_float(s->hs.material_alpha8[mid]);
    out[0] = m.base_color.x; out[1] = m.base_color.y; out[2] = m.base_color.z;
    out[3] = m.metallic; out[4] = m.specular; out[5] = m.roughness; out[6] = m.ior;
    if (transmission) {
        out[7] = m.specular_transmission; out[8] = m.transmission_roughness;
        out[9] = m.transmission_color.x; out[10] = m.transmission_color.y; out[11] = m.transmission_color.z;
    }
    out[12] = e.x; out[13] = e.y; out[14] = e.z;
}
void hostsim_halton_23(int32_t k, float *out) { halton_23(k, out); }
// the product's texture unit (csrc/rptr_shading.cuh sample_texture) on texture `id` of the scene: out = rgba
void hostsim_sample_texture(const hostsim_scene *s, uint32_t id, float u, float v, float *out) {
    const float4 c = sample_texture(s->dev(), id, f2(u, v));
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
// textureGrad / textureLod of the product's texture unit on texture `id` of the scene
void hostsim_sample_texture_grad(const hostsim_scene *s, uint32_t id, float u, float v, const float *ddx, const float *ddy, float *out) {
    const float4 c = sample_texture_grad(s->dev(), id, f2(u, v), f2(ddx[0], ddx[1]), f2(ddy[0], ddy[1]));
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
void hostsim_sample_texture_lod(const hostsim_scene *s, uint32_t id, float u, float v, int32_t level, float *out) {
    const float4 c = sample_texture_lod(s->dev(), id, f2(u, v), level);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
float hostsim_log2(float x) { return log2_pos(x); }
// texture ingestion of the product (csrc/rptr_host.cpp decode_texture): all levels as RGBA8; returns the number of bytes (out may be NULL)
int64_t hostsim_decode_texture(const rptr_texture_desc *td, uint8_t *out, int64_t capacity) {
    HostTexture ht;
    try { decode_texture(*td, ht); } catch (const std::exception &) { return -1; }
    if (out && (int64_t)ht.rgba.size() <= capacity) memcpy(out, ht.rgba.data(), ht.rgba.size());
    return (int64_t)ht.rgba.size();
}
// rendering/rt/footprint.glsl as the product states it: op 0 dpdxy_to_footprint(in: dir3, dpdx3, dpdy3) -> 4; op 1 reflect_footprint(in: dst3, src3,
// F4) -> 4; op 2 footprint_to_dpdxy(in: dir3, F4) -> dpdx3, dpdy3.  F = m00, m01, m10, m11 (GLSL F[c][r])
void hostsim_footprint_op(int32_t op, const float *in, float *out) {
    if (op == 0) {
        const Footprint f = dpdxy_to_footprint(f3(in[0], in[1], in[2]), f3(in[3], in[4], in[5]), f3(in[6], in[7], in[8]));
        out[0] = f.m00; out[1] = f.m01; out[2] = f.m10; out[3] = f.m11;
    } else if (op == 1) {
        const Footprint f = reflect_footprint(f3(in[0], in[1], in[2]), f3(in[3], in[4], in[5]), Footprint{in[6], in[7], in[8], in[9]});
        out[0] = f.m00; out[1] = f.m01; out[2] = f.m10; out[3] = f.m11;
    } else {
        float3 dx, dy;
        footprint_to_dpdxy(dx, dy, f3(in[0], in[1], in[2]), Footprint{in[3], in[4], in[5], in[6]});
        out[0] = dx.x; out[1] = dx.y; out[2] = dx.z; out[3] = dy.x; out[4] = dy.y; out[5] = dy.z;
    }
}
// the product's stochastic alpha test (csrc/rptr_bvh.cuh): closest-hit form (draws from the given LCG) and shadow-ray form
// (own LCG per candidate); 1 = rejected / 1 = the candidate blocks the ray
int32_t hostsim_alpha_rejects(float alpha, uint32_t *state) { return alpha_rejects(alpha, *state) ? 1 : 0; }
int32_t hostsim_shadow_candidate_passes(int32_t alpha8, int32_t prim, int32_t instance, uint32_t frame_id, uint32_t frame_offset, uint32_t pixel_linear) {
    GeomInst g;
    memset(&g, 0, sizeof(g));
    g.instance = instance;
    SceneDev sc{};
    sc.ginst = &g;
    const AlphaFilter af{sc, frame_id, frame_offset, pixel_linear};
    return shadow_candidate_passes(af, pack_gi_alpha(0, alpha8), prim, 0.0f, 0.0f) ? 1 : 0;
}
// generate_primary for one pixel sample: out = origin(3), dir(3), bits(first sampler word afterwards), tmin, tmax
void hostsim_primary_ray(const hostsim_scene *s, const hostsim_args *a, int32_t px, int32_t py, uint32_t sample_index, float *out) {
    FrameParams fp = make_frame(s, a);
    PathState ps;
    generate_primary(fp, px, py, sample_index, ps);
    out[0] = ps.o.x; out[1] = ps.o.y; out[2] = ps.o.z; out[3] = ps.d.x; out[4] = ps.d.y; out[5] = ps.d.z;
    memcpy(out + 6, &ps.rng, 4);
    out[7] = ps.tmin; out[8] = ps.tmax;
}
// the sky / sun-disc term of a missed path (k_resolve's shade_miss) with illum = 0, throughput = 1
void hostsim_shade_miss(const rptr_scene_params *sp, const float *dir, float prev_pdf, float *out) {
    const float3 r = shade_miss(*sp, f3(0.0f), f3(1.0f), f3(dir[0], dir[1], dir[2]), prev_pdf);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void hostsim_screen_jitter(uint32_t frame_offset, uint32_t frame_id, int32_t w, int32_t h, float *out) { screen_jitter(frame_offset, frame_id, w, h, out); }
uint32_t hostsim_morton_sample_id(uint32_t sample_id, uint32_t px, uint32_t py, uint32_t tw, uint32_t th, int hash_tile, int hash_sample) {
    return morton_sample_id(sample_id, px, py, tw, th, hash_tile != 0, hash_sample != 0);
}

} // extern "C"

// render_ray_queries through the product's shared code: one sample layer (sample index `layer`, accumulation_frame_offset = 0)
// of every query; the query's "pixel" comes from the product's TileMap in query mode (tile_pixel).  out = 4 floats per query.
extern "C" int hostsim_ray_query_layer(const hostsim_scene *s, const hostsim_args *a, const rptr_render_ray_query *queries, int32_t n, uint32_t layer, float *out) {
    FrameParams fp = make_frame(s, a);
    const SceneDev sc = s->dev();
    BvhDev bvh{s->hs.nodes.data(), s->hs.leaf_tris.data(), (int32_t)s->hs.nodes.size(), (int32_t)s->hs.leaf_tris.size()};
    TileMap tm{};
    tm.width = a->width; tm.height = a->height; tm.world = 1; tm.rows = 1; tm.local_pixels = n;
    tm.query_wgs_x = ((int)ceilf(sqrtf((float)n)) + 31) / 32;
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t q = 0; q < n; ++q) {
        uint32_t px, py;
        tile_pixel(tm, (uint32_t)q, px, py);
        PathState ps;
        generate_primary(fp, (int)px, (int)py, layer, ps);
        ps.o = f3(queries[q].origin[0], queries[q].origin[1], queries[q].origin[2]);
        ps.d = f3(queries[q].dir[0], queries[q].dir[1], queries[q].dir[2]);
        ps.tmax = queries[q].t_max;
        init_footprint(fp, ps); // the footprint block runs on the ray actually traced (pt_megakernel.glsl:326-351)
        uint32_t alpha_lcg = alpha_lcg_seed(fp, (int)px, (int)py, layer);
        TraceCounters cnt{0, 0};
        for (;;) {
            HitRec h;
            bool found = closest_hit_filtered(bvh, sc, ps.o, ps.d, ps.tmin, ps.tmax, fp.rng_variant == 0 ? ps.rng : alpha_lcg, h, cnt);
            ShadowRay sh;
            ShadeResult r = shade_vertex(fp, sc, ps, h.t, h.u, h.v, found ? &bvh.tris[h.tri] : nullptr, sh, nullptr);
            if (sh.tmax > 0.0f) {
                HitRec o;
                const AlphaFilter af{sc, fp.first_sample, fp.frame_offset, tile_pixel_linear(tm, (uint32_t)q)};
                if (!trace_ray<true>(bvh, sh.o, sh.d, sh.tmin, sh.tmax, o, cnt, sh.tmin, 0x7fffffff, &af)) ps.illum = ps.illum + sh.contrib;
            }
            if (r == SHADE_TERMINATE) break;
        }
        out[4 * q + 0] = ps.illum.x; out[4 * q + 1] = ps.illum.y; out[4 * q + 2] = ps.illum.z; out[4 * q + 3] = ps.bounce == 0 ? 0.0f : 1.0f;
    }
    return 0;
}

// traversal statistics of the product's tree over a set of closest-hit queries: out = nodes visited, triangles tested (totals)
extern "C" void hostsim_trace_stats(const hostsim_scene *s, const rptr_render_ray_query *q, int32_t n, uint64_t *out) {
    BvhDev bvh{s->hs.nodes.data(), s->hs.leaf_tris.data(), (int32_t)s->hs.nodes.size(), (int32_t)s->hs.leaf_tris.size()};
    uint64_t nodes = 0, tris = 0;
#pragma omp parallel for reduction(+ : nodes, tris)
    for (int32_t i = 0; i < n; ++i) {
        HitRec h;
        TraceCounters cnt{0, 0};
        float3 o = f3(q[i].origin[0], q[i].origin[1], q[i].origin[2]), d = f3(q[i].dir[0], q[i].dir[1], q[i].dir[2]);
        trace_ray<false>(bvh, o, d, RPTR_RAY_EPSILON * length(o), q[i].t_max, h, cnt);
        nodes += cnt.nodes; tris += cnt.tris;
    }
    out[0] = nodes; out[1] = tris;
}

// the product's query -> sampler pixel map (TileMap in query mode, csrc/rptr_shading.cuh); wgs_x as rptr_cuda_render_ray_queries computes it
extern "C" void hostsim_query_pixel(uint32_t q, int32_t n, uint32_t *out) {
    TileMap tm{};
    tm.width = 1; tm.world = 1; tm.rows = 1; tm.local_pixels = n;
    tm.query_wgs_x = ((int)std::ceil(std::sqrt((float)n)) + 31) / 32;
    tile_pixel(tm, q, out[0], out[1]);
}

// RQ_CLOSEST through the product's traversal code (tmin = q.mode_or_data reinterpreted as float when any != 0)
extern "C" int hostsim_trace(const hostsim_scene *s, const rptr_render_ray_query *q, int32_t n, float *results, float *hit_t, int32_t any) {
    BvhDev bvh{s->hs.nodes.data(), s->hs.leaf_tris.data(), (int32_t)s->hs.nodes.size(), (int32_t)s->hs.leaf_tris.size()};
    for (int32_t i = 0; i < n; ++i) {
        HitRec h;
        TraceCounters cnt{0, 0};
        if (q[i].mode_or_data < 0) continue;
        float3 o = f3(q[i].origin[0], q[i].origin[1], q[i].origin[2]), d = f3(q[i].dir[0], q[i].dir[1], q[i].dir[2]);
        const float tmin = RPTR_RAY_EPSILON * length(o);
        bool ok = any ? trace_ray<true>(bvh, o, d, tmin, q[i].t_max, h, cnt) : trace_ray<false>(bvh, o, d, tmin, q[i].t_max, h, cnt);
        int32_t gi = -1, prim = -1;
        if (ok) { gi = tri_geom_inst(bvh.tris[h.tri]); prim = bvh.tris[h.tri].prim; }
        results[4 * i + 0] = ok ? h.u : -1.0f;
        results[4 * i + 1] = ok ? h.v : -1.0f;
        memcpy(&results[4 * i + 2], &gi, 4);
        memcpy(&results[4 * i + 3], &prim, 4);
        if (hit_t) hit_t[i] = ok ? h.t : -1.0f;
    }
    return 0;
}

// Structure and conservativeness of the wide tree (rptr_bvh.cuh): every node and every leaf-order triangle is referenced exactly
// once, no slot is both an inner child and a triangle, and the vertices of every triangle lie strictly inside the decoded box of its own slot and of every
// ancestor slot on the way down from the root (what culling relies on).  out = {nodes, triangles, violations, deepest level}.
#include <functional>
extern "C" void hostsim_validate_bvh(const hostsim_scene *s, int64_t *out) {
    const std::vector<BvhNode> &nodes = s->hs.nodes;
    const std::vector<Tri> &tris = s->hs.leaf_tris;
    std::vector<int> seen_node(nodes.size(), 0), seen_tri(tris.size(), 0);
    int64_t bad = 0, depth = 0;
    struct Bx { float lo[3], hi[3]; };
    std::vector<Bx> path;
    std::function<void(int32_t)> walk = [&](int32_t ni) {
        depth = std::max<int64_t>(depth, (int64_t)path.size() + 1);
        const BvhNode &nd = nodes[ni];
        const uint32_t im = node_imask(nd), lm = node_lmask(nd);
        if (im & lm) bad++;
        for (int k = 0; k < RPTR_BVH_WIDTH; ++k) {
            if (!(((im | lm) >> k) & 1u)) continue;
            Bx bx;
            decode_child(nd, k, bx.lo, bx.hi);
            path.push_back(bx);
            if ((lm >> k) & 1u) {
                const int32_t ti = nd.tri_base + popc8(lm & ((1u << k) - 1u));
                if (ti < 0 || ti >= (int32_t)tris.size()) bad++;
                else {
                    seen_tri[ti]++;
                    const Tri &t = tris[ti];
                    const float v[3][3] = {{t.v0x, t.v0y, t.v0z}, {t.v0x + t.e1x, t.v0y + t.e1y, t.v0z + t.e1z}, {t.v0x + t.e2x, t.v0y + t.e2y, t.v0z + t.e2z}};
                    for (const Bx &anc : path)
                        for (int c = 0; c < 3; ++c)
                            for (int a = 0; a < 3; ++a)
                                if (!(anc.lo[a] < v[c][a] && v[c][a] < anc.hi[a])) bad++;
                }
            } else {
                const int32_t ci = nd.child_base + popc8(im & ((1u << k) - 1u));
                if (ci <= ni || ci >= (int32_t)nodes.size()) bad++;
                else { seen_node[ci]++; walk(ci); }
            }
            path.pop_back();
        }
    };
    if (!nodes.empty()) { seen_node[0] = 1; walk(0); }
    for (int c : seen_node) bad += c != 1;
    for (int c : seen_tri) bad += c != 1;
    out[0] = (int64_t)nodes.size(); out[1] = (int64_t)tris.size(); out[2] = bad; out[3] = depth;
}

// ---- the temporal passes (csrc/rptr_post.cuh), image by image -----------------------------------------------------------------
#include "../../realtimepathtracingresearchframework_b200/csrc/rptr_post.cuh"
extern "C" {
// process_samples.comp:106-113 over a whole frame; accum_cur = this frame's samples, stored / shown = accumulator and display colour
void hostsim_reproject(int32_t w, int32_t h, const float *accum_cur, const float *history, const uint16_t *nd_history, const uint16_t *nd,
                       const uint16_t *mj, float min_sample_weight, int32_t batch, float *stored, float *shown) {
    const ResolveImages im{w, h, reinterpret_cast<const float4 *>(history), reinterpret_cast<const ushort4 *>(nd_history),
                           reinterpret_cast<const ushort4 *>(nd), reinterpret_cast<const ushort4 *>(mj)};
    for (int32_t y = 0; y < h; ++y)
        for (int32_t x = 0; x < w; ++x) {
            const size_t i = (size_t)y * w + x;
            float4 st;
            const float4 sh = reproject_and_accumulate(im, reinterpret_cast<const float4 *>(accum_cur)[i], x, y, min_sample_weight, batch, &st);
            reinterpret_cast<float4 *>(stored)[i] = st;
            reinterpret_cast<float4 *>(shown)[i] = sh;
        }
}
void hostsim_taa(int32_t w, int32_t h, int32_t upscale, int32_t rw, int32_t rh, const uint8_t *current, const uint8_t *history, const uint16_t *mj,
                 uint8_t *out) {
    const TaaImages im{w, h, upscale, rw, rh, reinterpret_cast<const uchar4 *>(current), reinterpret_cast<const uchar4 *>(history),
                       reinterpret_cast<const ushort4 *>(mj)};
    for (int32_t y = 0; y < h; ++y)
        for (int32_t x = 0; x < w; ++x) reinterpret_cast<uchar4 *>(out)[(size_t)y * w + x] = process_taa_pixel(im, x, y);
}
}
