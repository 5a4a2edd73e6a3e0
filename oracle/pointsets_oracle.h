// oracle/pointsets_oracle.h -- TEST INFRASTRUCTURE: CPU restatement of the reference's low-discrepancy samplers
// (RenderBackendOptions::rng_variant = BN / SOBOL / Z_SBL), pinned by tests/golden/ref_pointsets.npz, which
// oracle/gen_golden.py generates by executing the reference's own rendering/pointsets/{sobol,sample_order,bn_rng}.glsl
// through oracle/_ref.  Written as a small class with one method per reference macro:
//   GET_RNG            -> QmcRng::seed          (sobol.glsl:156-194, bn_rng.glsl:91-101,112)
//   RANDOM_FLOAT1      -> QmcRng::draw          (sobol.glsl:196-205, bn_rng.glsl:31-82)
//   RANDOM_SET_DIM     -> QmcRng::set_dim       RANDOM_SHIFT_DIM -> QmcRng::shift_dim
#pragma once
#include <cstdint>

namespace oracle_ps {

enum { UNIFORM = 0, BN = 1, SOBOL = 2, Z_SBL = 3 };
enum { T_SOBOL_MATRIX = 0, T_SOBOL_INVERT = 1, T_BN_SOBOL = 2, T_BN_SCRAMBLE = 3 };

static inline uint32_t rotl(uint32_t v, int k) { return (v << k) | (v >> (32 - k)); }
static inline uint32_t mm3_mix(uint32_t h, uint32_t k) { // hashing.glsl:11-26
    k = rotl(k * 0xcc9e2d51u, 15) * 0x1b873593u;
    return rotl(h ^ k, 13) * 5u + 0xe6546b64u;
}
static inline uint32_t mm3_fin(uint32_t h) { // hashing.glsl:28-39
    h = (h ^ (h >> 16)) * 0x85ebca6bu;
    h = (h ^ (h >> 13)) * 0xc2b2ae35u;
    return h ^ (h >> 16);
}
static inline uint32_t hash_lcg(uint32_t index, uint32_t frame, uint32_t linear) { return mm3_fin(mm3_mix(mm3_mix(frame, linear), index)); }

static inline uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }
static inline int log2_exact(uint32_t p) { int k = 0; while ((1u << k) < p) ++k; return k; }
static inline uint32_t interleave_zero(uint32_t x) { // util.glsl:156-163, bit by bit
    uint32_t r = 0;
    for (int i = 0; i < 16; ++i) r |= ((x >> i) & 1u) << (2 * i);
    return r;
}

// sample_order.glsl:21-73
static inline uint32_t morton_sample_id(uint32_t sample_id, uint32_t px, uint32_t py, uint32_t tw, uint32_t th, bool hash_tile, bool hash_sample) {
    const uint32_t pw = next_pow2(tw), ph = next_pow2(th), n_tile = pw * ph;
    const uint32_t zx = interleave_zero(px), zy = interleave_zero(py);
    const uint32_t common = (pw - 1) & (ph - 1);
    const uint32_t side = common + 1, zmask = side * side - 1;
    uint32_t lin = ((zy << 1) + zx) & zmask;
    lin |= ((px | py) & ~common) * side;
    if (!hash_tile) lin &= n_tile - 1;
    uint32_t res = lin;
    uint32_t flip = zx ^ zy;
    flip |= flip << 1;
    const uint32_t h0 = hash_sample ? mm3_mix(0, sample_id) : 0;
    for (int level = log2_exact(side); level > 0; --level) { // bit pair [2*level-2, 2*level-1], hashed by everything above it
        const uint32_t perm = mm3_fin(mm3_mix(h0, lin >> (2 * level)));
        const int lo = 2 * level - 2;
        res ^= ((perm & 3u) << lo) & zmask;
        if (perm & 4u) {
            const uint32_t pair = 3u << lo;
            if ((zmask & pair) == pair) res ^= flip & pair;
        }
    }
    if (hash_tile) res &= n_tile - 1;
    return sample_id * n_tile + res;
}

struct QmcRng {
    int variant = UNIFORM;
    const uint32_t *const *tab = nullptr;
    uint32_t index = 0, scramble = 0; // Sobol
    uint32_t pixel = 0, sample = 0;   // BN
    uint32_t lcg = 0;                 // UNIFORM
    int dim = 0;

    uint32_t gen_matrix_xor(uint32_t d, uint32_t idx) const {
        uint32_t r = 0;
        for (int bit = 0; bit < 32; ++bit)
            if ((idx >> bit) & 1u) r ^= tab[T_SOBOL_MATRIX][d * 32 + bit];
        return r;
    }
    void seed(int v, const uint32_t *const *tables, uint32_t sample_index, uint32_t frame_id, uint32_t frame_offset, uint32_t px, uint32_t py,
              uint32_t w) {
        variant = v; tab = tables; dim = 0;
        if (v == UNIFORM) lcg = hash_lcg(sample_index, frame_offset, px + py * w);
        else if (v == BN) {
            pixel = (px % 128u) + (py % 128u) * 128u;
            sample = frame_id + frame_offset * 13u;
        } else if (v == SOBOL) {
            index = sample_index;
            scramble = hash_lcg(frame_offset, 0, px + py * w);
        } else {
            const uint32_t in_tile = morton_sample_id(0, px, py, 256, 256, true, false) % 65536u;
            const uint32_t shift = 65536u * sample_index;
            const uint32_t i = in_tile + shift;
            const uint32_t cx = gen_matrix_xor(0, i) >> 24, cy = gen_matrix_xor(1, i) >> 24;
            index = shift + tab[T_SOBOL_INVERT][cy * 256 + cx];
            scramble = hash_lcg(frame_offset, 0, (px / 256u) + (py / 256u) * (w / 256u));
        }
    }
    void set_dim(int d) { dim = d; }
    void shift_dim(int d) { dim += d; }
    float draw(int d) {
        if (variant == UNIFORM) {
            lcg = lcg * 1664525u + 1013904223u;
            return (float)lcg * 2.3283064365386963e-10f;
        }
        if (variant == BN) return bn((uint32_t)(dim + d));
        scramble = scramble * 1664525u + 1013904223u;
        const uint32_t dd = (uint32_t)(dim + d) % 1024u;
        uint32_t bits = scramble ^ gen_matrix_xor(dd, index);
        if (variant == Z_SBL && dd < 2) bits ^= bits << 8;
        return (float)bits * 2.3283064365386963e-10f;
    }
    float bn(uint32_t d) const {
        uint32_t pid = pixel;
        const uint32_t col_shift = d / 8u;
        auto shift_x = [](uint32_t p, uint32_t by) { return (p & ~127u) | ((p + by) & 127u); };
        auto shift_y = [](uint32_t p, uint32_t by) { return (p & ~(127u * 128u)) | ((p + by * 128u) & (127u * 128u)); };
        pid = shift_x(pid, col_shift);
        d = (d % 8u) + (col_shift / 128u) * 8u;
        d %= 256u;
        if (sample & 1u) pid ^= 127u;
        if (sample & 2u) pid ^= 127u * 128u;
        pid = shift_x(pid, sample * 73u);
        pid = shift_y(pid, sample * 97u);
        const uint32_t value = tab[T_BN_SOBOL][d] ^ tab[T_BN_SCRAMBLE][pid * 8u + (d % 8u)];
        return (0.5f + (float)value) / 256.0f;
    }
};

} // namespace oracle_ps
