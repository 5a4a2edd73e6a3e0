// rptr_pointsets.cuh -- the reference's low-discrepancy samplers (RenderBackendOptions::rng_variant):
//   RNG_VARIANT_BN     rendering/pointsets/bn_rng.glsl:31-118  (blue-noise dithered Sobol, 1-spp optimised tables)
//   RNG_VARIANT_SOBOL  rendering/pointsets/sobol.glsl:81-213   (Sobol + a fresh LCG digit scramble per draw)
//   RNG_VARIANT_Z_SBL  the same with Z_ORDER_SHUFFLING (sample_order.glsl:21-73: Morton-ordered, hashed 256x256 tiles)
// and the dimension bookkeeping of rendering/pathspace.h:9-38.  All integer work is exact; the only float step is
// uint -> float (round to nearest) times a power of two.  Tables are the reference's own data
// (sobol_tables.h / bn_tables.h), handed over through rptr_cuda_set_pointset_table().
#pragma once
#include "rptr_math.cuh"

namespace rp {

#define RPTR_SOBOL_DIMS 1024      // sobol_data.h:7-11
#define RPTR_SOBOL_MATRIX_SIZE 32
#define RPTR_SOBOL_TILE 256
#define RPTR_SOBOL_TILE_BITS 8
#define RPTR_BN_SAMPLES 256       // bn_data.h:7-10
#define RPTR_BN_DIMS 256
#define RPTR_BN_SCRAMBLING_DIMS 8
#define RPTR_BN_TILE 128

// rendering/pathspace.h: dimensions of one path vertex (the megakernel does not define USE_SIMPLIFIED_CAMERA)
#define RPTR_DIM_PIXEL_X 0
#define RPTR_DIM_CAMERA_END 6
#define RPTR_DIM_DIRECTION_X 0
#define RPTR_DIM_LOBE 2
#define RPTR_DIM_VERTEX_END 4
#define RPTR_DIM_RR (-1) // DIM_FREE_PATH - DIM_VERTEX_END
#define RPTR_DIM_LIGHT_SEL_1 0
#define RPTR_DIM_POSITION_X 2
#define RPTR_DIM_LIGHT_END 4

struct PointsetTables {
    const uint32_t *sobol_matrix;      // [1024 * 32]   SobolMatrix
    const uint32_t *sobol_tile_invert; // [256 * 256]   SobolInversion_1_0
    const uint32_t *bn_sobol;          // [256 * 256]   sobol_256spp_256d
    const uint32_t *bn_scrambling;     // [128*128*8]   scramblingTile_yx_d_1spp
};

RPTR_HD uint32_t ps_murmur_mix(uint32_t hash, uint32_t k) { // rendering/pointsets/hashing.glsl:11-26
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    hash ^= k;
    return ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
}
RPTR_HD uint32_t ps_murmur_finalize(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu;
    h ^= h >> 13; h *= 0xc2b2ae35u;
    return h ^ (h >> 16);
}
RPTR_HD uint32_t ps_lcg_next(uint32_t &s) { return s = s * 1664525u + 1013904223u; }
RPTR_HD float ps_unorm32(uint32_t v) { return (float)v * 2.3283064365386963e-10f; } // ldexp(float(v), -32)

RPTR_HD int ps_msb(uint32_t v) { // findMSB
    int r = -1;
    while (v) { v >>= 1; ++r; }
    return r;
}
RPTR_HD uint32_t ps_pow2_ceil(uint32_t v) { // sample_order.glsl:22-24: smallest power of two >= v
    uint32_t p = 1u << ps_msb(v);
    return p != v ? p << 1 : p;
}
RPTR_HD uint32_t ps_spread16(uint32_t x) { // util.glsl:156-163: bit i of the low half-word moves to bit 2i
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    return (x | (x << 1)) & 0x55555555u;
}

// sample_order.glsl:21-73.  Consecutive sample ids along a Z curve inside tiles of tile_w x tile_h pixels; every bit
// pair of the curve is permuted (and optionally transposed) by a hash of the more significant part of the index.
RPTR_HD uint32_t morton_sample_id(uint32_t sample_id, uint32_t px, uint32_t py, uint32_t tile_w, uint32_t tile_h, bool hash_tile_id,
                                 bool hash_sample_id) {
    const uint32_t pw = ps_pow2_ceil(tile_w), ph = ps_pow2_ceil(tile_h);
    const uint32_t tile_count = pw * ph;
    const uint32_t sx = ps_spread16(px), sy = ps_spread16(py);
    const uint32_t square = (pw - 1u) & (ph - 1u);                     // bits present in both dimensions
    const uint32_t interleaved = (square + 1u) * (square + 1u) - 1u;   // ... and their interleaved range
    uint32_t id = ((sy << 1) + sx) & interleaved;
    id |= ((px | py) & ~square) * (square + 1u);                       // bits of the longer dimension go on top
    if (!hash_tile_id) id &= tile_count - 1u;
    uint32_t out = id;
    uint32_t transpose = sx ^ sy;
    transpose |= transpose << 1;
    const uint32_t seed = hash_sample_id ? ps_murmur_mix(0u, sample_id) : 0u;
    for (int e = 2 * ps_msb(square + 1u); e > 0;) {
        const uint32_t h = ps_murmur_finalize(ps_murmur_mix(seed, id >> e));
        e -= 2;
        out ^= ((h & 3u) << e) & interleaved;
        const uint32_t pair = (h & 4u) ? (3u << e) : 0u;
        if (pair == (interleaved & pair)) out ^= transpose & pair;
    }
    if (hash_tile_id) out &= tile_count - 1u;
    return sample_id * tile_count + out;
}

// XOR of the columns of generator matrix `dim` selected by the bits of index (sobol.glsl:81-91, 116-127)
RPTR_HD uint32_t sobol_xor(const uint32_t *matrix, uint32_t dim, uint32_t index, uint32_t acc) {
    const uint32_t *col = matrix + dim * RPTR_SOBOL_MATRIX_SIZE;
    for (; index != 0u; index >>= 1, ++col)
        if (index & 1u) acc ^= *col;
    return acc;
}
RPTR_HD float sobol_point(const PointsetTables &t, uint32_t index, uint32_t dimension, uint32_t scramble, bool z_order) {
    dimension &= (uint32_t)(RPTR_SOBOL_DIMS - 1);
    uint32_t r = sobol_xor(t.sobol_matrix, dimension, index, scramble);
    if (z_order && dimension < 2u) r ^= r << RPTR_SOBOL_TILE_BITS; // sobol.glsl:93-108
    return ps_unorm32(r);
}
// sobol.glsl:113-133: the sample of the (0,1) projection that falls into the same pixel of the 256x256 tile
RPTR_HD uint32_t sobol_shift_invert(const PointsetTables &t, uint32_t index, uint32_t index_shift) {
    index += index_shift;
    const uint32_t r0 = sobol_xor(t.sobol_matrix, 0u, index, 0u) >> (32 - RPTR_SOBOL_TILE_BITS);
    const uint32_t r1 = sobol_xor(t.sobol_matrix, 1u, index, 0u) >> (32 - RPTR_SOBOL_TILE_BITS);
    return index_shift + t.sobol_tile_invert[r1 * RPTR_SOBOL_TILE + r0];
}

// bn_rng.glsl:31-82 with BN_OPTIMIZED_DIMENSION_REPEAT and BN_OPTIMIZED_SPP == 1
RPTR_HD float sample_bnd(const PointsetTables &t, uint32_t pixel_id, uint32_t sample_id, uint32_t d) {
    const uint32_t T = RPTR_BN_TILE, S = RPTR_BN_SCRAMBLING_DIMS;
    const uint32_t x_doffset = d / S;
    pixel_id = ((pixel_id + x_doffset) & (T - 1u)) + (pixel_id & ~(T - 1u));
    d = (d & (S - 1u)) + x_doffset / T * S;
    d &= (uint32_t)(RPTR_BN_DIMS - 1);
    if (sample_id & 1u) pixel_id ^= T - 1u;
    if (sample_id & 2u) pixel_id ^= (T - 1u) * T;
    const uint32_t xs = sample_id * 73u, ys = sample_id * 97u;
    pixel_id = ((pixel_id + xs) & (T - 1u)) + (pixel_id & ~(T - 1u));
    pixel_id = ((pixel_id + ys * T) & (T * (T - 1u))) + (pixel_id & ~(T * (T - 1u)));
    sample_id = 0u; // sampleID & (BN_OPTIMIZED_SPP - 1)
    const uint32_t rank = pixel_id * S + (d & (S - 1u));
    uint32_t v = t.bn_sobol[d + sample_id * RPTR_BN_DIMS];
    v ^= t.bn_scrambling[rank];
    return (0.5f + (float)v) / 256.0f;
}

// RANDOM_STATE of the selected variant, packed into the two words the wavefront keeps per path:
//   UNIFORM: a = LCG state                         SOBOL / Z_SBL: a = scramble LCG state, b = Sobol index
//   BN:      a = sampleID, b = pixelID             dim: RANDOM_SET_DIM / RANDOM_SHIFT_DIM cursor (not stored: 6 + 8 * bounce)
struct Sampler {
    uint32_t a, b;
    int32_t dim;
};

// GET_RNG(sample_index, frame_offset, uvec4(pixel, frame_dims)) (pt_megakernel.glsl:314; bn_rng.glsl:112 takes frame_id instead)
RPTR_HD Sampler sampler_init(int variant, const PointsetTables &t, uint32_t sample_index, uint32_t frame_id, uint32_t frame_offset, uint32_t px,
                            uint32_t py, uint32_t width) {
    Sampler s;
    s.dim = 0;
    if (variant == 1) { // BN
        s.b = (px & (RPTR_BN_TILE - 1u)) + (py & (RPTR_BN_TILE - 1u)) * RPTR_BN_TILE;
        s.a = frame_id + frame_offset * 13u;
    } else if (variant == 2) { // SOBOL: per-pixel scrambling
        s.b = sample_index;
        s.a = ps_murmur_finalize(ps_murmur_mix(ps_murmur_mix(0u, px + py * width), frame_offset));
    } else if (variant == 3) { // Z_SBL: 65536 Sobol samples per tile, per-tile scrambling
        const uint32_t off = morton_sample_id(0u, px, py, RPTR_SOBOL_TILE, RPTR_SOBOL_TILE, true, false) & (RPTR_SOBOL_TILE * RPTR_SOBOL_TILE - 1u);
        s.b = sobol_shift_invert(t, off, RPTR_SOBOL_TILE * RPTR_SOBOL_TILE * sample_index);
        const uint32_t linear = (px >> RPTR_SOBOL_TILE_BITS) + (py >> RPTR_SOBOL_TILE_BITS) * (width >> RPTR_SOBOL_TILE_BITS);
        s.a = ps_murmur_finalize(ps_murmur_mix(ps_murmur_mix(0u, linear), frame_offset));
    } else { // UNIFORM (lcg_rng.glsl:28-39)
        s.b = 0u;
        s.a = ps_murmur_finalize(ps_murmur_mix(ps_murmur_mix(frame_offset, px + py * width), sample_index));
    }
    return s;
}
// RANDOM_FLOAT1(rng, d)
RPTR_HD float sampler_next(int variant, const PointsetTables &t, Sampler &s, int d) {
    if (variant == 2 || variant == 3) return sobol_point(t, s.b, (uint32_t)(s.dim + d), ps_lcg_next(s.a), variant == 3);
    if (variant == 1) return sample_bnd(t, s.b, s.a, (uint32_t)(s.dim + d));
    return ps_unorm32(ps_lcg_next(s.a));
}

} // namespace rp
