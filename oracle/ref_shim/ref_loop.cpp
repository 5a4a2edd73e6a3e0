// oracle/ref_shim/ref_loop.cpp -- TEST INFRASTRUCTURE.
// Blocks of the megakernel (vulkan/pt_megakernel.glsl, inside main_spp) executed as C++: the ray-generation head, the bounce
// prologue (approximate solid angle, shading point, face-forwarding, normal map, "fix incident direction", tangent frame)
// and the Russian-roulette step; plus geometry_scale_to_tmin (vulkan/geometry.glsl) and the running-mean statement of the
// resolve pass (vulkan/process_samples.comp).  Most are not functions in the reference, so oracle/Makefile cuts the line
// ranges out of the files where they lie (by their first / last statements) into the git-ignored build directory
// oracle/_ref/gen/ for the duration of the compile; this file supplies the local variables and uniforms the blocks read.
// Compiled without USE_MIPMAPPING: the ray-differential footprint between the two halves of the prologue only feeds texture
// LODs (irrelevant for 1 x 1 textures) and total_t, which the caller accumulates.
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "../../include/rptr_types.h"

namespace refloop {
using namespace glm;
typedef unsigned int uint;
#define UNROLL_STANDARD_TEXTURES // vulkan/gpu_params.glsl: only names the slot the stand-in sampler below ignores
#include "rendering/language.hpp"
#include "rendering/defaults.glsl"
#include "rendering/pointsets/lcg_rng.glsl"
#include "rendering/pathspace.h"
#include "rendering/util.glsl"
#include "rendering/bsdfs/base_material.h.glsl"
#include "rendering/bsdfs/hit_point.glsl"
#define DEFAULT_GEOMETRY_BUFFER_TYPES
#define QUANTIZED_POSITIONS
#define QUANTIZED_NORMALS_AND_UVS
#define NEED_MESH_ID_FOR_VISUALIZATION 0
#include "rendering/rt/hit.glsl"

struct { float normal_z_scale; } scene_params;
struct { int rr_path_depth; int enable_raster_taa; } render_params;
// the GLSL swizzles the ray-generation head spells out (.xy / .xyz) as members of stand-in types
struct SwzU3 { uvec2 xy; };
struct SwzU2 { uvec2 xy; operator vec2() const { return vec2(xy); } }; // uvec2 -> vec2 is implicit in GLSL
struct SwzF3 { vec3 xyz; };
struct { SwzU2 frame_dims; vec2 screen_jitter; SwzF3 cam_pos, cam_du, cam_dv, cam_dir_top_left; } view_params;
static SwzU3 gl_GlobalInvocationID;
#define SAMPLE_PIXEL_FILTER(urand) (urand - vec2(0.5f)) // vulkan/gpu_params.glsl:42 (no PIXEL_FILTER_TENT_WINDOW in the build)
#define RAY_EPSILON 0.000005f                           // vulkan/gpu_params.glsl:27-29
#include "gen/geometry_scale.inc"
struct { int bounce; } shading_state;
static BaseMaterial material_params[1];
struct Texel { vec3 rgb; };
static Texel g_normal_texel;
struct SamplerStandIn {};
inline SamplerStandIn get_standard_texture_sampler(int, int, int) { return SamplerStandIn(); }
inline Texel textureLod(SamplerStandIn, vec2, float) { return g_normal_texel; }
static float g_rr_sample;
#define nonuniformEXT(x) (x)
#define EXPLICIT_MASK_BEGIN
#define EXPLICIT_MASK_END
#define RANDOM_FLOAT1(rng, dim) g_rr_sample

static void prologue(RTHit hit, vec3 ray_origin, vec3 ray_dir, float *out) {
    int instanceIdx = 0, primitiveIdx = 0;
    vec3 motion_vector = vec3(0.0f);
#include "gen/prologue.inc"
    out[0] = approx_tri_solid_angle;
    out[1] = interaction.p.x; out[2] = interaction.p.y; out[3] = interaction.p.z;
    out[4] = interaction.gn.x; out[5] = interaction.gn.y; out[6] = interaction.gn.z;
    out[7] = interaction.n.x; out[8] = interaction.n.y; out[9] = interaction.n.z;
    out[10] = interaction.v_x.x; out[11] = interaction.v_x.y; out[12] = interaction.v_x.z;
    out[13] = interaction.v_y.x; out[14] = interaction.v_y.y; out[15] = interaction.v_y.z;
    out[16] = hit.dist;
#undef w_o
}

static void camera_head(uint sample_index, uint rnd_offset, float *out) {
#include "gen/raygen.inc"
    out[0] = ray_origin.x; out[1] = ray_origin.y; out[2] = ray_origin.z;
    out[3] = ray_dir.x; out[4] = ray_dir.y; out[5] = ray_dir.z;
    std::memcpy(out + 6, &rng.state, 4);
    out[7] = t_min; out[8] = t_max;
}

// the resolve's running mean (process_samples.comp:121-127) around a history texel
struct HistoryStandIn { vec4 v; };
inline vec4 textureLod(HistoryStandIn h, vec2, float) { return h.v; }
static vec4 running_mean(vec4 accum_color, vec4 history, uint sample_base_index, uint sample_batch_size) {
    HistoryStandIn history_buffer{history};
    ivec2 fb_pixel(0, 0);
    ivec2 fb_dims(1, 1);
#include "gen/running_mean.inc"
    return accum_color;
}

static bool russian_roulette(vec3 &path_throughput) {
    do {
#include "gen/rr.inc"
        return true;
    } while (false);
    return false;
}
} // namespace refloop

extern "C" {

// in: [0..2] hit.normal, [3] hit.dist, [4..6] hit.geo_normal (area-scaled), [7..9] hit.tangent, [10] hit.bitangent_l,
//     [11..13] ray_origin, [14..16] ray_dir, [17..19] normal-map texel as sampled
// out: [0] approx_tri_solid_angle, [1..3] p, [4..6] gn, [7..9] n, [10..12] v_x, [13..15] v_y, [16] hit.dist afterwards
void ref_bounce_prologue(const float *in, uint32_t material_flags, int32_t normal_map, float normal_z_scale, float *out) {
    using namespace refloop;
    RTHit hit;
    std::memset(&hit, 0, sizeof(hit));
    hit.normal = glm::vec3(in[0], in[1], in[2]);
    hit.dist = in[3];
    hit.geo_normal = glm::vec3(in[4], in[5], in[6]);
    hit.tangent = glm::vec3(in[7], in[8], in[9]);
    hit.bitangent_l = in[10];
    hit.material_id = 0;
    std::memset(&material_params[0], 0, sizeof(material_params[0]));
    material_params[0].flags = material_flags;
    material_params[0].normal_map = normal_map;
    scene_params.normal_z_scale = normal_z_scale;
    shading_state.bounce = 0;
    g_normal_texel.rgb = glm::vec3(in[17], in[18], in[19]);
    prologue(hit, glm::vec3(in[11], in[12], in[13]), glm::vec3(in[14], in[15], in[16]), out);
}

// returns 1 when the path survives; throughput is updated in place
int32_t ref_russian_roulette(int32_t bounce, int32_t rr_path_depth, float *throughput, float rr_sample) {
    using namespace refloop;
    shading_state.bounce = bounce;
    render_params.rr_path_depth = rr_path_depth;
    g_rr_sample = rr_sample;
    glm::vec3 t(throughput[0], throughput[1], throughput[2]);
    const bool alive = russian_roulette(t);
    throughput[0] = t.x; throughput[1] = t.y; throughput[2] = t.z;
    return alive ? 1 : 0;
}

// head of main_spp (vulkan/pt_megakernel.glsl:311-325).  cam = cam_pos(3), cam_du(3), cam_dv(3), cam_dir_top_left(3);
// out = origin(3), dir(3), bits(LCG state afterwards), t_min, t_max
void ref_camera_ray(const float *cam, uint32_t width, uint32_t height, uint32_t px, uint32_t py, uint32_t sample_index, uint32_t frame_offset,
                    int32_t enable_raster_taa, const float *screen_jitter, float *out) {
    using namespace refloop;
    view_params.frame_dims.xy = glm::uvec2(width, height);
    view_params.screen_jitter = glm::vec2(screen_jitter[0], screen_jitter[1]);
    view_params.cam_pos.xyz = glm::vec3(cam[0], cam[1], cam[2]);
    view_params.cam_du.xyz = glm::vec3(cam[3], cam[4], cam[5]);
    view_params.cam_dv.xyz = glm::vec3(cam[6], cam[7], cam[8]);
    view_params.cam_dir_top_left.xyz = glm::vec3(cam[9], cam[10], cam[11]);
    render_params.enable_raster_taa = enable_raster_taa;
    gl_GlobalInvocationID.xy = glm::uvec2(px, py);
    camera_head(sample_index, frame_offset, out);
}

float ref_geometry_scale_to_tmin(const float *orig, float geometry_scale) {
    return refloop::geometry_scale_to_tmin(glm::vec3(orig[0], orig[1], orig[2]), geometry_scale);
}

// history += (x - history) / float(sample_base_index + sample_batch_size), in place on history[4]
void ref_running_mean(const float *x, float *history, uint32_t sample_base_index, uint32_t sample_batch_size) {
    glm::vec4 r = refloop::running_mean(glm::vec4(x[0], x[1], x[2], x[3]), glm::vec4(history[0], history[1], history[2], history[3]), sample_base_index,
                                        sample_batch_size);
    history[0] = r.x; history[1] = r.y; history[2] = r.z; history[3] = r.w;
}

} // extern "C"
