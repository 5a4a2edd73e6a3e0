// oracle/ref_shim/ref_pointsets.cpp -- TEST INFRASTRUCTURE.
// extern "C" wrappers around the reference's OWN low-discrepancy samplers, #included from where they lie under REF:
// rendering/pointsets/sobol.glsl (plain and with Z_ORDER_SHUFFLING), sample_order.glsl, bn_rng.glsl and their tables
// (sobol_tables.h, bn_tables.h), exactly the way rendering/tests/compile.cpp:9-37 and vulkan/pointsets/render_{sobol,bn}.cpp
// use them.  Nothing here restates reference arithmetic; the glue replays a list of (op, arg) sampler calls.
#include <glm/glm.hpp>
#include <cstdint>
#include <cstring>

#include "rendering/pointsets/sobol_tables.h"
#include "rendering/pointsets/bn_tables.h"
#include "librender/halton.h"

#define MAKE_RANDOM_TABLE(TYPE, NAME) static TYPE NAME;

namespace refps {
using namespace glm;
typedef unsigned int uint;
#include "rendering/language.hpp"
#include "rendering/util.glsl"
#include "rendering/pointsets/lcg_rng.glsl"
#include "rendering/pointsets/sobol_data.h"
namespace plain {
#include "rendering/pointsets/sobol.glsl"
}
#undef SOBOL_RNG_GLSL
#undef RANDOM_STATE
#undef RANDOM_FLOAT1
#undef RANDOM_SHIFT_DIM
#undef RANDOM_SET_DIM
#undef GET_RNG
#undef PACK_RNG
#undef UNPACK_RNG
#undef COMPRESSED_RANDOM_STATE
#define Z_ORDER_SHUFFLING
namespace zorder {
#include "rendering/pointsets/sobol.glsl"
}
#undef Z_ORDER_SHUFFLING
#undef RANDOM_STATE
#undef RANDOM_FLOAT1
#undef RANDOM_SHIFT_DIM
#undef RANDOM_SET_DIM
#undef GET_RNG
#undef PACK_RNG
#undef UNPACK_RNG
#undef COMPRESSED_RANDOM_STATE
#define out /* GLSL qualifier on two pack helpers that are not called here */
namespace bn {
#include "rendering/pointsets/bn_rng.glsl"
}
#undef out
} // namespace refps

static bool g_loaded = false;
static void load_tables() { // vulkan/pointsets/render_sobol.cpp:77-104, render_bn.cpp:77-110
    if (g_loaded) return;
    using namespace refps;
    static_assert(sizeof(plain::sobol_table.matrix) == sizeof(SobolMatrix), "");
    static_assert(sizeof(plain::sobol_table.tile_invert_1_0) == sizeof(SobolInversion_1_0), "");
    std::memcpy(&plain::sobol_table.matrix, SobolMatrix, sizeof(SobolMatrix));
    std::memcpy(&plain::sobol_table.tile_invert_1_0, SobolInversion_1_0, sizeof(SobolInversion_1_0));
    std::memcpy(&zorder::sobol_table, &plain::sobol_table, sizeof(plain::sobol_table));
    std::memcpy(&bn::bn_pointset_table.sobol_spp_d, sobol_256spp_256d, sizeof(sobol_256spp_256d));
    std::memcpy(&bn::bn_pointset_table.tile_scrambling_yx_d_1spp, scramblingTile_yx_d_1spp, sizeof(scramblingTile_yx_d_1spp));
    g_loaded = true;
}

extern "C" {

// which: 0 = Sobol matrices [1024*32], 1 = Sobol tile inversion [256*256], 2 = BN sobol_256spp_256d [256*256],
// 3 = BN scramblingTile_yx_d_1spp [128*128*8].  Returns the element count; copies when out != NULL.
int ref_pointset_table(int which, uint32_t *out_) {
    load_tables();
    const void *src = nullptr;
    int n = 0;
    switch (which) {
    case 0: src = SobolMatrix; n = sizeof(SobolMatrix) / 4; break;
    case 1: src = SobolInversion_1_0; n = sizeof(SobolInversion_1_0) / 4; break;
    case 2: src = sobol_256spp_256d; n = sizeof(sobol_256spp_256d) / 4; break;
    case 3: src = scramblingTile_yx_d_1spp; n = sizeof(scramblingTile_yx_d_1spp) / 4; break;
    default: return 0;
    }
    if (out_) std::memcpy(out_, src, (size_t)n * 4);
    return n;
}

// Replays sampler calls on RANDOM_STATE rng = GET_RNG(sample_index, frame_offset, uvec4(px, py, w, h)) of the given
// variant (librender/render_params.glsl.h:34-37: 1 = BN, 2 = SOBOL, 3 = Z_SBL).  ops[i] = 0: RANDOM_FLOAT1(rng, args[i])
// -> *out++; 1: RANDOM_SET_DIM(rng, args[i]); 2: RANDOM_SHIFT_DIM(rng, args[i]).  BN takes (frame_id, frame_offset) as
// its GET_RNG does (bn_rng.glsl:112).  state_out (optional): index / pixelID, sampleID after construction.
int ref_pointset_replay(int variant, uint32_t sample_index, uint32_t frame_id, uint32_t frame_offset, uint32_t px, uint32_t py, uint32_t w,
                        uint32_t h, const int32_t *ops, const int32_t *args, int n_ops, float *out_, uint32_t *state_out) {
    load_tables();
    using namespace refps;
    const uvec4 pd(px, py, w, h);
    int n_out = 0;
    if (variant == 2 || variant == 3) {
#define REPLAY(NS)                                                                                     \
    {                                                                                                  \
        NS::SobolRand rng = NS::get_sobol_rng(sample_index, frame_offset, pd);                         \
        if (state_out) { state_out[0] = rng.index; state_out[1] = rng.scramble.state; }                \
        for (int i = 0; i < n_ops; ++i) {                                                              \
            if (ops[i] == 0) out_[n_out++] = NS::sobol_randomf(rng, (uint32_t)args[i]);                \
            else if (ops[i] == 1) NS::sobol_set_dim(rng, (uint32_t)args[i]);                           \
            else NS::sobol_shift_dim(rng, (uint32_t)args[i]);                                          \
        }                                                                                              \
    }
        if (variant == 2) REPLAY(plain) else REPLAY(zorder)
#undef REPLAY
    } else if (variant == 1) {
        bn::BNDState rng = bn::get_bnd_rng(frame_id, frame_offset, pd);
        if (state_out) { state_out[0] = rng.pixelID; state_out[1] = rng.sampleID; }
        for (int i = 0; i < n_ops; ++i) {
            if (ops[i] == 0) out_[n_out++] = bn::sample_bnd(rng.pixelID, rng.sampleID, (uint32_t)(rng.dimension + args[i]));
            else if (ops[i] == 1) rng.dimension = args[i];
            else rng.dimension += args[i];
        }
    } else
        return -1;
    return n_out;
}

// librender/halton.h:12-81: the table update_view_parameters reads for the raster-TAA screen jitter (render_vulkan.cpp:2917-2926)
int ref_halton_23(float *out_) {
    const int n = (int)halton_23_size;
    if (out_) std::memcpy(out_, halton_23, sizeof(float) * 2 * n);
    return n;
}

// the raster-TAA screen jitter statements of update_view_parameters (vulkan/render_vulkan.cpp:2922-2924), cut out of the
// reference's host code by the Makefile; RASTER_TAA_NUM_SAMPLES = 16 is the build's definition (CMakeLists.txt:30)
void ref_screen_jitter(uint32_t frame_offset, uint32_t frame_id, uint32_t w, uint32_t h, float *out_) {
    constexpr size_t num_sample_offsets = 16;
    struct { glm::uvec2 frame_dims; glm::vec2 screen_jitter; } viewParams;
    viewParams.frame_dims = glm::uvec2(w, h);
#include "gen/screen_jitter.inc"
    out_[0] = viewParams.screen_jitter.x; out_[1] = viewParams.screen_jitter.y;
}

uint32_t ref_morton_sample_id(uint32_t sample_id, uint32_t px, uint32_t py, uint32_t tw, uint32_t th, int hash_tile_id, int hash_sample_id) {
    return refps::zorder::morton_sample_id(sample_id, glm::uvec2(px, py), glm::uvec2(tw, th), hash_tile_id != 0, hash_sample_id != 0);
}

} // extern "C"
