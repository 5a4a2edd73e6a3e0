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
struct { SwzU2 frame_dims; vec2 screen_jitter; SwzF3 cam_pos, cam_du, cam_dv, cam_dir_top_left; uint frame_id, frame_offset; } view_params;
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

// ---- raytrace_test_visibility (vulkan/pt_megakernel.glsl:216-272) over a scripted ray query ------------------------------
// The GL_EXT_ray_query calls are answered from a list of non-opaque candidates (+ "an opaque triangle was hit", which the
// traversal commits by itself); generate_candidate_hit() is replaced by a recorder that notes the alpha LCG it was handed
// and answers from the script.  What runs from the reference: the epsilon / range / skip rule, the per-candidate seed,
// the loop's termination and the final verdict.
struct Candidate { float t; int prim, inst, geom, accept; };
static const Candidate *g_cands = nullptr;
static int g_ncands = 0, g_opaque_hit = 0;
struct rayQueryEXT { int cursor; bool terminated; };
static struct { vec3 from, dir; float tmin, tmax; int inits; } g_rq;
static uint g_seeds[64];
static int g_nseeds;
static int scene;
static float geometry_scale;
enum { gl_RayFlagsTerminateOnFirstHitEXT = 4, gl_RayFlagsSkipClosestHitShaderEXT = 8, gl_RayQueryCandidateIntersectionTriangleEXT = 0 };
inline void rayQueryInitializeEXT(rayQueryEXT &q, int, uint, uint, vec3 o, float tmin, vec3 d, float tmax) {
    g_rq.from = o; g_rq.dir = d; g_rq.tmin = tmin; g_rq.tmax = tmax; ++g_rq.inits;
    q.cursor = -1; q.terminated = false;
}
inline bool rayQueryProceedEXT(rayQueryEXT &q) { return !q.terminated && ++q.cursor < g_ncands; }
inline void rayQueryTerminateEXT(rayQueryEXT &q) { q.terminated = true; }
// committed == false: type of the current candidate (always a triangle); committed == true: 1 when an opaque triangle was hit
inline uint rayQueryGetIntersectionTypeEXT(rayQueryEXT &, bool committed) { return committed ? (g_opaque_hit ? 1u : 0u) : 0u; }
inline float rayQueryGetIntersectionTEXT(rayQueryEXT &q, bool) { return g_cands[q.cursor].t; }
inline vec2 rayQueryGetIntersectionBarycentricsEXT(rayQueryEXT &, bool) { return vec2(0.25f); }
inline int rayQueryGetIntersectionInstanceIdEXT(rayQueryEXT &q, bool) { return g_cands[q.cursor].inst; }
inline int rayQueryGetIntersectionInstanceCustomIndexEXT(rayQueryEXT &q, bool) { return g_cands[q.cursor].geom; }
inline int rayQueryGetIntersectionGeometryIndexEXT(rayQueryEXT &, bool) { return 0; }
inline int rayQueryGetIntersectionPrimitiveIndexEXT(rayQueryEXT &q, bool) { return g_cands[q.cursor].prim; }
inline vec3 rayQueryGetIntersectionObjectRayOriginEXT(rayQueryEXT &, bool) { return g_rq.from; }
inline vec3 rayQueryGetIntersectionObjectRayDirectionEXT(rayQueryEXT &, bool) { return g_rq.dir; }
static int g_candidate_cursor;
inline bool generate_candidate_hit(float, vec2, int, int, int primitiveIdx, vec3, vec3, RTHit &, LCGRand &alpha_rng, bool) {
    if (g_nseeds < 64) g_seeds[g_nseeds++] = alpha_rng.state;
    for (int i = 0; i < g_ncands; ++i)
        if (g_cands[i].prim == primitiveIdx) return !g_cands[i].accept; // true = candidate rejected (transparent)
    return false;
}
#define IMPLICIT_INSTANCE_PARAMS // vulkan/gpu_params.glsl:16
#include "gen/test_visibility.inc"

// the alpha test at the end of generate_candidate_hit (:202-210)
static float g_alpha;
inline float get_material_alpha(int, const BaseMaterial &, const HitPoint &) { return g_alpha; }
#define MATERIAL_PARAMS BaseMaterial
static bool alpha_tail(RTHit hit, LCGRand &alpha_rng) {
    vec3 local_ray_orig(0.0f), local_ray_dir(0.0f, 0.0f, 1.0f);
    float dist = 1.0f;
#include "gen/alpha_tail.inc"
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

// raytrace_test_visibility(from, dir, dist) with geometry_scale and the frame counters / pixel of the invocation, over the scripted
// candidates cands[n] = (t, prim, inst, accept) and opaque_hit.  out: [0] visible, [1] number of ray queries started,
// [2] tmin, [3] tmax, [4] number of candidates judged, [5..] bits of the alpha LCG state handed to each
void ref_test_visibility(const float *from, const float *dir, float dist, float geom_scale, uint32_t frame_id, uint32_t frame_offset, uint32_t px,
                         uint32_t py, uint32_t width, uint32_t height, const float *cands, int32_t n, int32_t opaque_hit, float *out) {
    using namespace refloop;
    static Candidate list[64];
    for (int i = 0; i < n && i < 64; ++i) list[i] = Candidate{cands[4 * i], (int)cands[4 * i + 1], (int)cands[4 * i + 2], 0, (int)cands[4 * i + 3]};
    g_cands = list; g_ncands = n < 64 ? n : 64; g_opaque_hit = opaque_hit;
    g_rq.inits = 0; g_rq.tmin = 0.0f; g_rq.tmax = 0.0f; g_nseeds = 0;
    geometry_scale = geom_scale;
    view_params.frame_id = frame_id; view_params.frame_offset = frame_offset;
    view_params.frame_dims.xy = glm::uvec2(width, height);
    gl_GlobalInvocationID.xy = glm::uvec2(px, py);
    const bool visible = raytrace_test_visibility(glm::vec3(from[0], from[1], from[2]), glm::vec3(dir[0], dir[1], dir[2]), dist);
    out[0] = visible ? 1.0f : 0.0f; out[1] = (float)g_rq.inits; out[2] = g_rq.tmin; out[3] = g_rq.tmax; out[4] = (float)g_nseeds;
    std::memcpy(out + 5, g_seeds, sizeof(uint) * g_nseeds);
}

// 1 = the candidate is rejected (the ray passes through); *lcg_state advances when a draw was needed
int32_t ref_alpha_filter(float alpha, uint32_t material_flags, uint32_t *lcg_state) {
    using namespace refloop;
    RTHit hit;
    std::memset(&hit, 0, sizeof(hit));
    std::memset(&material_params[0], 0, sizeof(material_params[0]));
    material_params[0].flags = material_flags;
    g_alpha = alpha;
    LCGRand rng;
    rng.state = *lcg_state;
    const bool rejected = alpha_tail(hit, rng);
    *lcg_state = rng.state;
    return rejected ? 1 : 0;
}

} // extern "C"

// ---- camera basis of RenderVulkan::update_view_parameters (vulkan/render_vulkan.cpp:2887-2895), host C++ of the reference ----
// (the statements from `glm::vec2 img_plane_size;` to `dir_top_left`, cut out by the Makefile; glm calls go to the shim)
namespace refcam {
struct RenderTargetStandIn { glm::uvec2 d; glm::uvec2 dims() const { return d; } };
static void camera_basis(const glm::vec3 &dir, const glm::vec3 &up, const float fovy, uint32_t w, uint32_t h, float *out) {
    RenderTargetStandIn rt{glm::uvec2(w, h)};
    RenderTargetStandIn *render_targets[1] = {&rt};
#include "gen/camera_basis.inc"
    out[0] = dir_du.x; out[1] = dir_du.y; out[2] = dir_du.z;
    out[3] = dir_dv.x; out[4] = dir_dv.y; out[5] = dir_dv.z;
    out[6] = dir_top_left.x; out[7] = dir_top_left.y; out[8] = dir_top_left.z;
}
} // namespace refcam

extern "C" {
// out = du(3), dv(3), top_left(3)
void ref_view_params(const rptr_camera_params *cam, uint32_t w, uint32_t h, float *out) {
    refcam::camera_basis(glm::vec3(cam->dir[0], cam->dir[1], cam->dir[2]), glm::vec3(cam->up[0], cam->up[1], cam->up[2]), cam->fovy, w, h, out);
}
} // extern "C"
