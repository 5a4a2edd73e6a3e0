// oracle/ref_shim/ref_loop.cpp -- TEST INFRASTRUCTURE.
// Two blocks of the megakernel's bounce loop (vulkan/pt_megakernel.glsl, inside main_spp) executed as C++: the bounce
// prologue (approximate solid angle, shading point, face-forwarding, normal map, "fix incident direction", tangent frame)
// and the Russian-roulette step.  They are not functions in the reference, so oracle/Makefile cuts the two line ranges out
// of the file where it lies (by their first / last statements) into the git-ignored build directory oracle/_ref/gen/ for
// the duration of the compile; this file supplies the local variables and uniforms the blocks read.
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
#include "rendering/util.glsl"
#include "rendering/bsdfs/base_material.h.glsl"
#include "rendering/bsdfs/hit_point.glsl"
#define DEFAULT_GEOMETRY_BUFFER_TYPES
#define QUANTIZED_POSITIONS
#define QUANTIZED_NORMALS_AND_UVS
#define NEED_MESH_ID_FOR_VISUALIZATION 0
#include "rendering/rt/hit.glsl"

struct { float normal_z_scale; } scene_params;
struct { int rr_path_depth; } render_params;
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

} // extern "C"
