// oracle/ref_shim/ref_lights.cpp -- TEST INFRASTRUCTURE.
// extern "C" wrapper around the reference's host light pre-pass, linked against librender/lights.cpp compiled
// from where it lies under REF (see oracle/Makefile).
#include <glm/glm.hpp>
#include <cstdint>
#include <vector>
#include "../../include/rptr_types.h"
#include "librender/lights.h"

static inline glm::vec3 V(const float *p) { return glm::vec3(p[0], p[1], p[2]); }
static inline void S(float *o, glm::vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

extern "C" {
// librender/lights.cpp:75-90 (update_light_sampling on a fresh BinnedLightSampling)
int32_t ref_bin_emitters(const rptr_tri_light_data *in, int32_t n, const rptr_light_sampling_config *ls, rptr_tri_light_data *out, int32_t max_out) {
    std::vector<TriLight> em(n);
    for (int i = 0; i < n; ++i) {
        em[i].v0 = V(in[i].v0); em[i].v1 = V(in[i].v1); em[i].v2 = V(in[i].v2); em[i].radiance = V(in[i].radiance);
    }
    BinnedLightSampling binned;
    LightSamplingConfig cfg;
    cfg.light_mis_angle = ls->light_mis_angle;
    cfg.bin_size = ls->bin_size;
    cfg.min_perceived_receiver_dist = ls->min_perceived_receiver_dist;
    cfg.min_radiance = ls->min_radiance;
    update_light_sampling(binned, em, cfg);
    int m = int(binned.emitters.size());
    if (m > max_out) return -m;
    for (int i = 0; i < m; ++i) {
        S(out[i].v0, binned.emitters[i].v0); S(out[i].v1, binned.emitters[i].v1); S(out[i].v2, binned.emitters[i].v2);
        S(out[i].radiance, binned.emitters[i].radiance);
    }
    return m;
}


} // extern "C"
