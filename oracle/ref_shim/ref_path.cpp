// oracle/ref_shim/ref_path.cpp -- TEST INFRASTRUCTURE.
// ONE PATH SAMPLE of the megakernel (vulkan/pt_megakernel.glsl: main_spp) composed from the reference-executed pieces of the
// other files of this directory, in the order the megakernel runs them:
//   ray-generation head (gen/raygen.inc)                                   -> ref_camera_ray
//   per bounce:  closest hit                                                -> CALLBACK (the reference leaves it to the driver)
//                miss: compute_sky_illum (gen/compute_sky_illum.inc)        -> ref_compute_sky_illum
//                calc_hit_vertices / calc_hit_attributes (rt/hit.glsl)      -> ref_hit_attributes
//                total_t, geometry_scale (gen/total_t.inc, cut out of the megakernel for this file)
//                bounce prologue (gen/prologue.inc)                         -> ref_bounce_prologue
//                shade_base_material (mc/shade_base_material.glsl + nee.glsl + the glTF BSDF)  -> ref_shade_base_material, whose
//                    raytrace_test_visibility calls back into this file: range from the reference's own function
//                    (gen/test_visibility.inc -> ref_test_visibility), occlusion from the CALLBACK
//                next ray (gen/next_ray.inc, cut out of the megakernel for this file)
//                Russian roulette (gen/rr.inc)                              -> ref_russian_roulette
//   result vec4(illum, bounce == 0 ? 0 : 1) (gen/path_result.inc)
// and folded into the accumulator by the resolve's running mean (gen/running_mean.inc -> ref_running_mean).
// What is written here by hand: the for / break skeleton of the bounce loop (the reference builds it from a recursive
// #include and EXPLICIT_MASK macros) and the marshalling between the pieces.  Opaque scenes only (no alpha candidates).
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "../../include/rptr_types.h"

extern "C" {
void ref_camera_ray(const float *cam, uint32_t width, uint32_t height, uint32_t px, uint32_t py, uint32_t sample_index, uint32_t frame_offset,
                    int32_t enable_raster_taa, const float *screen_jitter, float *out);
void ref_compute_sky_illum(const rptr_scene_params *sp, const float *ray_origin, const float *ray_dir, float prev_bsdf_pdf, float *out);
void ref_hit_attributes(const uint64_t *qverts3, const uint64_t *qnuv3, const float *scale, const float *offset, int has_normals, int has_uvs,
                        const float *w2o, int material_id, const uint32_t *id_4pack, uint32_t prim, float t, float u, float v, float *out);
void ref_bounce_prologue(const float *in, uint32_t material_flags, int32_t normal_map, float normal_z_scale, float *out);
void ref_shade_base_material(const rptr_base_material *p, int bounce, int output_channel, float prev_bounce_pdf, const float *illum,
                             const float *throughput, float approx_sa, const float *wo, const float *ia, uint32_t rng_state, int max_path_depth,
                             int glossy_only_mode, const float *sun_dir, float sun_cos_angle, const float *sun_radiance,
                             const rptr_tri_light_data *lights, int n_lights, int bin_size, float *out);
void ref_shade_set_visibility(int (*cb)(void *, const float *, const float *, float), void *user);
int32_t ref_russian_roulette(int32_t bounce, int32_t rr_path_depth, float *throughput, float rr_sample);
float ref_geometry_scale_to_tmin(const float *orig, float geometry_scale);
void ref_running_mean(const float *x, float *history, uint32_t sample_base_index, uint32_t sample_batch_size);
void ref_test_visibility(const float *from, const float *dir, float dist, float geom_scale, uint32_t frame_id, uint32_t frame_offset, uint32_t px,
                         uint32_t py, uint32_t width, uint32_t height, const float *cands, int32_t n, int32_t opaque_hit, float *out);
float ref_lcg_randomf(uint32_t *state);
}

// what the closest-hit callback returns about the triangle it found: the inputs of calc_hit_vertices / calc_hit_attributes
struct ref_path_hit {
    float t, u, v;
    const uint64_t *qverts3; // the triangle's three quantised vertices
    const uint64_t *qnuv3;   // its three normal / uv words (or null)
    float scale[3], offset[3];
    int32_t has_normals, has_uvs;
    float w2o[9];            // rows of inverse(mat3(object_to_world))
    int32_t material_id;     // RenderMeshParams::material_id (negative: per-triangle ids)
    const uint32_t *id_4pack; // per-triangle material ids of the geometry, four per word (or null)
    uint32_t prim;
};
typedef int (*ref_closest_fn)(void *user, const float *o, const float *d, float tmin, float tmax, ref_path_hit *hit);
typedef int (*ref_occluded_fn)(void *user, const float *o, const float *d, float tmin, float tmax);

struct ref_path_args {
    float cam[12]; // cam_pos, cam_du, cam_dv, cam_dir_top_left
    uint32_t width, height, frame_offset, frame_id;
    int32_t max_path_depth, rr_path_depth, glossy_only_mode, output_channel;
    rptr_scene_params sp; // sun_radiance[3] = p_sun as the shader sees it
    const rptr_base_material *materials;
    const rptr_tri_light_data *lights;
    int32_t n_lights, bin_size;
    ref_closest_fn closest;
    ref_occluded_fn occluded;
    void *user;
};

namespace refpath {
using namespace glm;
typedef unsigned int uint;
#define SHADING_RESULT_TERMINATE -1 // rendering/mc/shading_interface.glsl:7-10
#define SHADING_RESULT_NULL 0
#define SHADING_RESULT_BOUNCE 1
struct ShadingStateStandIn { int bounce; };

struct VisCtx { const ref_path_args *a; float geometry_scale; uint32_t px, py; };
static VisCtx g_vis;
// raytrace_test_visibility as the megakernel defines it (:216-272): the reference's own statements decide whether a ray is cast
// and over which range (first call, scripted "nothing hit"), the occlusion answer comes from the callback, the verdict from a
// second run of the same statements with that answer
static int visibility(void *, const float *from, const float *dir, float dist) {
    float out[5 + 64];
    ref_test_visibility(from, dir, dist, g_vis.geometry_scale, g_vis.a->frame_id, g_vis.a->frame_offset, g_vis.px, g_vis.py, g_vis.a->width,
                        g_vis.a->height, nullptr, 0, 0, out);
    if (out[1] == 0.0f) return out[0] != 0.0f; // no ray query started: the skip rule decided
    const int occ = g_vis.a->occluded(g_vis.a->user, from, dir, out[2], out[3]);
    ref_test_visibility(from, dir, dist, g_vis.geometry_scale, g_vis.a->frame_id, g_vis.a->frame_offset, g_vis.px, g_vis.py, g_vis.a->width,
                        g_vis.a->height, nullptr, 0, occ, out);
    return out[0] != 0.0f;
}
inline float geometry_scale_to_tmin(vec3 o, float s) { const float f[3] = {o.x, o.y, o.z}; return ref_geometry_scale_to_tmin(f, s); }

static vec4 path_sample(const ref_path_args &a, uint32_t px, uint32_t py, uint32_t sample_index) {
    float cr[9];
    const float no_jitter[2] = {0.0f, 0.0f};
    ref_camera_ray(a.cam, a.width, a.height, px, py, sample_index, a.frame_offset, 0, no_jitter, cr);
    vec3 ray_origin(cr[0], cr[1], cr[2]), ray_dir(cr[3], cr[4], cr[5]);
    uint32_t rng_state;
    std::memcpy(&rng_state, &cr[6], 4);
    float t_min = cr[7], t_max = cr[8];
    float total_t = 0.0f, geometry_scale = 0.0f;
    vec3 illum(0.0f), path_throughput(1.0f);
    ShadingStateStandIn shading_state{0};
    float prev_bounce_pdf = 2.e16f; // init_shading_sample_state (rendering/mc/shading_interface.glsl:19-22)
    g_vis.a = &a; g_vis.px = px; g_vis.py = py;
    ref_shade_set_visibility(visibility, nullptr);
    for (int unrollBounceIdx = 0; unrollBounceIdx < a.max_path_depth; ++unrollBounceIdx) {
        ref_path_hit ph;
        std::memset(&ph, 0, sizeof(ph));
        const float o3[3] = {ray_origin.x, ray_origin.y, ray_origin.z}, d3[3] = {ray_dir.x, ray_dir.y, ray_dir.z};
        const bool was_miss = !a.closest(a.user, o3, d3, t_min, t_max, &ph);
        if (was_miss) {
            float sky[3];
            ref_compute_sky_illum(&a.sp, o3, d3, prev_bounce_pdf, sky);
            illum += path_throughput * vec3(sky[0], sky[1], sky[2]);
            break;
        }
        float ha[14];
        ref_hit_attributes(ph.qverts3, ph.qnuv3, ph.scale, ph.offset, ph.has_normals, ph.has_uvs, ph.w2o, ph.material_id, ph.id_4pack, ph.prim, ph.t,
                           ph.u, ph.v, ha);
        struct { float dist; } hit{ha[3]};
        {
#include "gen/total_t.inc"
        }
        g_vis.geometry_scale = geometry_scale;
        const int material_id = (int)ha[7];
        const rptr_base_material &mp = a.materials[material_id];
        float pin[20], pout[17];
        pin[0] = ha[0]; pin[1] = ha[1]; pin[2] = ha[2]; pin[3] = ha[3]; pin[4] = ha[4]; pin[5] = ha[5]; pin[6] = ha[6];
        pin[7] = ha[8]; pin[8] = ha[9]; pin[9] = ha[10]; pin[10] = ha[11];
        pin[11] = o3[0]; pin[12] = o3[1]; pin[13] = o3[2]; pin[14] = d3[0]; pin[15] = d3[1]; pin[16] = d3[2];
        pin[17] = pin[18] = pin[19] = 0.0f;
        ref_bounce_prologue(pin, mp.flags, -1, a.sp.normal_z_scale, pout);
        const float wo[3] = {-d3[0], -d3[1], -d3[2]};
        const float il[3] = {illum.x, illum.y, illum.z}, thr[3] = {path_throughput.x, path_throughput.y, path_throughput.z};
        float so[19];
        ref_shade_base_material(&mp, shading_state.bounce, a.output_channel, prev_bounce_pdf, il, thr, pout[0], wo, pout + 1, rng_state, a.max_path_depth,
                                a.glossy_only_mode, a.sp.sun_dir, a.sp.sun_cos_angle, a.sp.sun_radiance, a.lights, a.n_lights, a.bin_size, so);
        const int shading_result = (int)so[0];
        shading_state.bounce = (int)so[1];
        prev_bounce_pdf = so[2];
        illum = vec3(so[3], so[4], so[5]);
        path_throughput = vec3(so[6], so[7], so[8]);
        const vec3 w_i(so[9], so[10], so[11]);
        std::memcpy(&rng_state, &so[13], 4);
        struct { vec3 p; } interaction{vec3(pout[1], pout[2], pout[3])};
        if (shading_result == SHADING_RESULT_TERMINATE) break;
        {
#include "gen/next_ray.inc"
        }
        if (shading_state.bounce >= a.rr_path_depth) { // the condition is part of gen/rr.inc too; the draw must only happen when it holds
            const float rr_sample = ref_lcg_randomf(&rng_state);
            float t3[3] = {path_throughput.x, path_throughput.y, path_throughput.z};
            const bool alive = ref_russian_roulette(shading_state.bounce, a.rr_path_depth, t3, rr_sample) != 0;
            path_throughput = vec3(t3[0], t3[1], t3[2]);
            if (!alive) break;
        }
    }
    ref_shade_set_visibility(nullptr, nullptr);
#include "gen/path_result.inc"
}
} // namespace refpath

extern "C" {
// running mean over n_samples frames of batch_spp = 1 (sample indices first_sample ...) for the pixels of the region, into
// rgba (W * H * 4 floats, row-major); first_sample > 0 continues an accumulation
void ref_path_render(const ref_path_args *a, int32_t x0, int32_t y0, int32_t x1, int32_t y1, uint32_t first_sample, int32_t n_samples, float *rgba) {
    for (int32_t y = y0; y < y1; ++y)
        for (int32_t x = x0; x < x1; ++x) {
            float *px = rgba + 4 * ((size_t)y * a->width + x);
            for (int32_t k = 0; k < n_samples; ++k) {
                const uint32_t frame_id = first_sample + (uint32_t)k;
                const glm::vec4 c = refpath::path_sample(*a, (uint32_t)x, (uint32_t)y, frame_id);
                const float xs[4] = {c.x, c.y, c.z, c.w};
                if (frame_id > 0) ref_running_mean(xs, px, frame_id, 1u); // process_samples.comp:116-127
                else std::memcpy(px, xs, 16);
            }
        }
}
} // extern "C"
