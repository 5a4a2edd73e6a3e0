// oracle/rptr_oracle.cpp -- TEST INFRASTRUCTURE. CPU restatement of the reference's per-pixel-sample path
// tracing loop (PT_MEGAKERNEL with the LCG sampler).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product (librptr_cuda.so) never does.
//
// PARITY STATUS: the reference ships no golden vectors, scenes or images for this path, and its
// traversal/intersection and GLSL built-ins live in the Vulkan driver.  Everything the reference can execute as C++ is pinned
// against oracle/_ref (the reference's own sources compiled from /root/reference, see oracle/Makefile and oracle/ref_shim/)
// through the fixtures of tests/golden/ (tests/test_oracle_golden.py, tests/test_pointsets.py): LCG / murmur, the Sobol /
// Z-order Sobol / blue-noise samplers and their tables, the Halton screen jitter, dequantisation, hit attributes, the glTF BSDF
// (with and without transmission), triangle-light solid angles / sampling / binned RIS, host light binning, the sky fit and
// skymodel_radiance, sun sampling, the material decode with texture handles, sample_direct_light (nee.glsl), the complete
// per-vertex shading function shade_base_material() with its LCG draw order, the miss shading compute_sky_illum(), and blocks
// cut out of the shader files at build time (oracle/ref_shim/ref_loop.cpp): the ray-generation head, the bounce prologue and
// the Russian-roulette step of main_spp, the alpha test of generate_candidate_hit, raytrace_test_visibility over scripted ray
// queries, geometry_scale_to_tmin, the running mean of process_samples.comp, the camera basis of update_view_parameters,
// the body of rt_intersect.comp.
// "PARITY UNPINNED" (restated only, no reference-executed check possible): the glue of pt_megakernel.glsl between those pieces
// (loop control, the closest-hit rayQueryEXT candidate loop whose candidate order is the driver's), accumulate.glsl, the
// texture unit (UNORM8 / sRGB decode of a texel), the ray/triangle routine, which the reference does not contain at all, and
// view_params.VP / VP_reference behind the motion / jitter AOV image: built with glm 0.9.9.8, a configure-time download of
// the reference (ext/CMakeLists.txt:18-21) that is not in its tree -- glm's published operator*, inverse and
// infinitePerspective are restated (oracle_view_projection) and checked against the camera model only.
//
// Structure follows the reference megakernel (one sequential loop per pixel sample), NOT the CUDA wavefront.
#include "shading_oracle.h"
#include "pointsets_oracle.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

// ------------------------------------------------------------------------------------------------------------
// scene: instances flattened to world-space triangles (the closest-hit contract of SURVEY 8a-4: Moeller-Trumbore
// on world-space (v0, e1, e2), no culling, hit iff tmin < t < tmax, ties -> lowest flattened triangle id)
// ------------------------------------------------------------------------------------------------------------
struct GeomInst { // one entry of the reference's instanced_geometry[] (rendering/rt/geometry.h.glsl:72-98)
    const uint64_t *qverts;
    const uint64_t *qnuv;
    float scale[3], offset[3];
    int n_tris;
    bool has_normals, has_uvs;
    int material_id;         // >= 0: constant; < 0: -1 - material_offset with per-triangle ids
    const uint8_t *tri_mat;  // per-triangle ids of this geometry (already offset by primOffset)
    uint32_t flags;
    int instance;
    float o2w[12];  // object-to-world rows
    V3 w2o_row[3];  // rows of inverse(mat3(o2w)): columns of normals_to_world = transpose(world_to_object)
    int64_t first_tri;
};
struct Tri {
    V3 v0, e1, e2;
    int geom_inst, prim;
};
struct Node {
    float bmin[3], bmax[3];
    int left, right; // inner: children; leaf: left = -1 - first, right = count
};

static inline V3 xfm_point(const float *m, V3 v) {
    return v3(fmaf(m[0], v.x, fmaf(m[1], v.y, fmaf(m[2], v.z, m[3]))), fmaf(m[4], v.x, fmaf(m[5], v.y, fmaf(m[6], v.z, m[7]))),
              fmaf(m[8], v.x, fmaf(m[9], v.y, fmaf(m[10], v.z, m[11]))));
}

struct Scene {
    std::vector<GeomInst> ginst;
    std::vector<Tri> tris;
    std::vector<int> tri_order; // bvh leaf order -> triangle id
    std::vector<Node> nodes;
    std::vector<rptr_base_material> materials;
    std::vector<rptr_tri_light_data> lights;
    bool any_non_opaque = false;
    // owned copies of the scene's textures (1 x 1 texel mode)
    std::vector<rptr_texture_desc> textures;
    std::vector<std::vector<uint8_t>> own_texels;
    TextureSet texset;
    // owned copies of the input streams (the caller's Scene dies after set_scene, app.cpp:151-175)
    std::vector<std::vector<uint64_t>> own_qv, own_qn;
    std::vector<std::vector<uint8_t>> own_tm;
};

// --- binned SAH BVH2 over padded triangle boxes ---------------------------------------------------------------
struct BuildPrim { float bmin[3], bmax[3], c[3]; int id; };

static float g_abs_pad = 0.0f; // 2^-18 * scene extent, set by build_bvh
static void tri_bounds(const Tri &t, float *mn, float *mx) {
    V3 a = t.v0, b = t.v0 + t.e1, c = t.v0 + t.e2;
    float pa[3][3] = {{a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}};
    for (int k = 0; k < 3; ++k) {
        mn[k] = fminf(pa[0][k], fminf(pa[1][k], pa[2][k]));
        mx[k] = fmaxf(pa[0][k], fmaxf(pa[1][k], pa[2][k]));
        // conservative padding so that box culling never rejects a hit the triangle routine accepts
        float pad = 1.52587890625e-05f * fmaxf(fabsf(mn[k]), fabsf(mx[k])) + g_abs_pad + 1e-30f;
        mn[k] -= pad;
        mx[k] += pad;
    }
}

static int build_rec(Scene &s, std::vector<BuildPrim> &p, int lo, int hi) {
    int idx = (int)s.nodes.size();
    s.nodes.push_back(Node());
    float bmin[3] = {1e30f, 1e30f, 1e30f}, bmax[3] = {-1e30f, -1e30f, -1e30f}, cmin[3] = {1e30f, 1e30f, 1e30f},
          cmax[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; ++i)
        for (int k = 0; k < 3; ++k) {
            bmin[k] = fminf(bmin[k], p[i].bmin[k]);
            bmax[k] = fmaxf(bmax[k], p[i].bmax[k]);
            cmin[k] = fminf(cmin[k], p[i].c[k]);
            cmax[k] = fmaxf(cmax[k], p[i].c[k]);
        }
    for (int k = 0; k < 3; ++k) {
        s.nodes[idx].bmin[k] = bmin[k];
        s.nodes[idx].bmax[k] = bmax[k];
    }
    int n = hi - lo;
    auto make_leaf = [&]() {
        s.nodes[idx].left = -1 - (int)s.tri_order.size();
        s.nodes[idx].right = n;
        for (int i = lo; i < hi; ++i) s.tri_order.push_back(p[i].id);
        return idx;
    };
    if (n <= 2) return make_leaf();
    const int NB = 16;
    int best_axis = -1, best_bin = -1;
    float best_cost = 1e30f;
    auto area = [](const float *mn, const float *mx) {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    };
    for (int ax = 0; ax < 3; ++ax) {
        float ext = cmax[ax] - cmin[ax];
        if (!(ext > 0.0f)) continue;
        float bmn[NB][3], bmx[NB][3];
        int cnt[NB];
        for (int b = 0; b < NB; ++b) {
            cnt[b] = 0;
            for (int k = 0; k < 3; ++k) { bmn[b][k] = 1e30f; bmx[b][k] = -1e30f; }
        }
        float sc = (float)NB / ext;
        for (int i = lo; i < hi; ++i) {
            int b = std::min(NB - 1, std::max(0, (int)((p[i].c[ax] - cmin[ax]) * sc)));
            cnt[b]++;
            for (int k = 0; k < 3; ++k) {
                bmn[b][k] = fminf(bmn[b][k], p[i].bmin[k]);
                bmx[b][k] = fmaxf(bmx[b][k], p[i].bmax[k]);
            }
        }
        float ra[NB];
        int rc[NB];
        float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
        int c = 0;
        for (int b = NB - 1; b > 0; --b) {
            for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], bmn[b][k]); mx[k] = fmaxf(mx[k], bmx[b][k]); }
            c += cnt[b];
            ra[b] = c ? area(mn, mx) : 0.0f;
            rc[b] = c;
        }
        for (int k = 0; k < 3; ++k) { mn[k] = 1e30f; mx[k] = -1e30f; }
        c = 0;
        for (int b = 0; b < NB - 1; ++b) {
            for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], bmn[b][k]); mx[k] = fmaxf(mx[k], bmx[b][k]); }
            c += cnt[b];
            if (c == 0 || rc[b + 1] == 0) continue;
            float cost = area(mn, mx) * (float)c + ra[b + 1] * (float)rc[b + 1];
            if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
        }
    }
    int mid;
    if (best_axis < 0) {
        if (n <= 8) return make_leaf();
        mid = lo + n / 2;
    } else {
        float leaf_cost = area(bmin, bmax) * (float)n;
        if (n <= 4 && leaf_cost <= best_cost + area(bmin, bmax)) return make_leaf();
        float ext = cmax[best_axis] - cmin[best_axis];
        float sc = (float)NB / ext;
        float cm = cmin[best_axis];
        int ax = best_axis, bb = best_bin;
        auto it = std::partition(p.begin() + lo, p.begin() + hi, [&](const BuildPrim &q) {
            int b = std::min(NB - 1, std::max(0, (int)((q.c[ax] - cm) * sc)));
            return b <= bb;
        });
        mid = (int)(it - p.begin());
        if (mid == lo || mid == hi) mid = lo + n / 2;
    }
    int l = build_rec(s, p, lo, mid);
    int r = build_rec(s, p, mid, hi);
    s.nodes[idx].left = l;
    s.nodes[idx].right = r;
    return idx;
}

static void build_bvh(Scene &s) {
    s.nodes.clear();
    s.tri_order.clear();
    std::vector<BuildPrim> p(s.tris.size());
    float extent = 0.0f;
    for (const Tri &t : s.tris) {
        V3 b = t.v0 + t.e1, c = t.v0 + t.e2;
        extent = fmaxf(extent, fmaxf(max3(vabs(t.v0)), fmaxf(max3(vabs(b)), max3(vabs(c)))));
    }
    g_abs_pad = 3.814697265625e-06f * extent;
    for (size_t i = 0; i < s.tris.size(); ++i) {
        tri_bounds(s.tris[i], p[i].bmin, p[i].bmax);
        for (int k = 0; k < 3; ++k) p[i].c[k] = 0.5f * (p[i].bmin[k] + p[i].bmax[k]);
        p[i].id = (int)i;
    }
    if (p.empty()) return;
    s.nodes.reserve(p.size());
    build_rec(s, p, 0, (int)p.size());
}

// --- ray/triangle: the shared closest-hit contract -------------------------------------------------------------
struct Hit { float t, u, v; int tri; };

static inline bool intersect_tri(const Tri &tr, V3 o, V3 d, float &t, float &u, float &v) {
    V3 p = cross(d, tr.e2);
    float det = dot(tr.e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    V3 s = o - tr.v0;
    u = dot(s, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    V3 q = cross(s, tr.e1);
    v = dot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(tr.e2, q) * inv;
    return true;
}

static inline bool slab(const Node &n, V3 o, V3 inv, float tmin, float tmax) {
    float t0 = (n.bmin[0] - o.x) * inv.x, t1 = (n.bmax[0] - o.x) * inv.x;
    float tn = fminf(t0, t1), tf = fmaxf(t0, t1);
    t0 = (n.bmin[1] - o.y) * inv.y; t1 = (n.bmax[1] - o.y) * inv.y;
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    t0 = (n.bmin[2] - o.z) * inv.z; t1 = (n.bmax[2] - o.z) * inv.z;
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    tf *= 1.0000004f; // keep the slab test conservative against rounding
    return tn <= tf && tf >= tmin && tn <= tmax;
}

// after = (t0, id0): only hits strictly after that key in (t, id) order qualify (used to continue past
// alpha-rejected candidates front to back).  Pass t0 = tmin, id0 = INT_MAX for a plain query.
static bool closest_hit(const Scene &s, V3 o, V3 d, float tmin, float tmax, float after_t, int after_id, Hit &best) {
    best.tri = -1;
    best.t = tmax;
    if (s.nodes.empty()) return false;
    V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Node &n = s.nodes[stack[--sp]];
        if (!slab(n, o, inv, tmin, best.t)) continue;
        if (n.left < 0) {
            int first = -1 - n.left;
            for (int i = 0; i < n.right; ++i) {
                int id = s.tri_order[first + i];
                float t, u, v;
                if (!intersect_tri(s.tris[id], o, d, t, u, v)) continue;
                if (!(t > tmin && t < tmax)) continue;
                if (!(t > after_t || (t == after_t && id > after_id))) continue;
                if (best.tri < 0 || t < best.t || (t == best.t && id < best.tri)) {
                    best.t = t; best.u = u; best.v = v; best.tri = id;
                }
            }
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.right;
        }
    }
    return best.tri >= 0;
}

// calls f(tri id, t, u, v) for every triangle hit in (tmin, tmax) until it returns true (= occluded)
template <class F>
static bool any_hit(const Scene &s, V3 o, V3 d, float tmin, float tmax, F &&f) {
    if (s.nodes.empty()) return false;
    V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Node &n = s.nodes[stack[--sp]];
        if (!slab(n, o, inv, tmin, tmax)) continue;
        if (n.left < 0) {
            int first = -1 - n.left;
            for (int i = 0; i < n.right; ++i) {
                int id = s.tri_order[first + i];
                float t, u, v;
                if (!intersect_tri(s.tris[id], o, d, t, u, v)) continue;
                if (!(t > tmin && t < tmax)) continue;
                if (f(id, t, u, v)) return true;
            }
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.right;
        }
    }
    return false;
}

// --- hit attributes: rendering/rt/hit.glsl:49-128,162-203 -------------------------------------------------------
struct RTHit {
    V3 normal; float dist; V3 geo_normal; int material_id; V3 tangent; float bitangent_l; V2 uv;
};

static inline int calc_hit_material_id(const GeomInst &g, uint32_t prim) { // hit.glsl:49-56
    if (g.material_id < 0) return (int)g.tri_mat[prim] - g.material_id - 1;
    return g.material_id;
}

static RTHit calc_hit_attributes(const GeomInst &g, float ray_t, uint32_t prim, float ax, float ay) {
    RTHit h;
    h.dist = ray_t;
    V3 p0 = dequantize_position(g.qverts[3 * (size_t)prim + 0], g.scale, g.offset);
    V3 p1 = dequantize_position(g.qverts[3 * (size_t)prim + 1], g.scale, g.offset);
    V3 p2 = dequantize_position(g.qverts[3 * (size_t)prim + 2], g.scale, g.offset);
    V3 gn = cross(p1 - p0, p2 - p0);
    V3 bary = v3(1.0f - ax - ay, ax, ay);
    V3 n = gn;
    uint64_t qa = 0, qb = 0, qc = 0;
    if (g.has_normals || g.has_uvs) {
        qa = g.qnuv[3 * (size_t)prim + 0];
        qb = g.qnuv[3 * (size_t)prim + 1];
        qc = g.qnuv[3 * (size_t)prim + 2];
    }
    if (g.has_normals) {
        n = mat_mul(dequantize_normal((uint32_t)qa), dequantize_normal((uint32_t)qb), dequantize_normal((uint32_t)qc), bary);
        if (dot(n, gn) < 0.0f) gn = -gn;
    }
    h.geo_normal = gn * 0.5f;
    h.normal = n;
    V2 uva{0, 0}, uvb{0, 0}, uvc{0, 0};
    h.uv = V2{0.0f, 0.0f};
    if (g.has_uvs) {
        uva = dequantize_uv((uint32_t)(qa >> 32));
        uvb = dequantize_uv((uint32_t)(qb >> 32));
        uvc = dequantize_uv((uint32_t)(qc >> 32));
        h.uv = V2{fmaf(uvc.x, bary.z, fmaf(uvb.x, bary.y, uva.x * bary.x)), fmaf(uvc.y, bary.z, fmaf(uvb.y, bary.y, uva.y * bary.x))};
    }
    h.material_id = calc_hit_material_id(g, prim);
    const V3 *r = g.w2o_row;
    h.geo_normal = mat_mul(r[0], r[1], r[2], h.geo_normal);
    h.normal = normalize(mat_mul(r[0], r[1], r[2], h.normal));
    bool requires_tangent = true;
    if (g.has_uvs) {
        float det = length(gn);
        V3 frame_n = gn / (det * det);
        V3 dp2perp = cross(p2 - p0, frame_n);
        V3 dp1perp = cross(frame_n, p1 - p0);
        V2 duv1{uvb.x - uva.x, uvb.y - uva.y}, duv2{uvc.x - uva.x, uvc.y - uva.y};
        V3 T = dp2perp * duv1.x + dp1perp * duv2.x;
        V3 B = dp2perp * duv1.y + dp1perp * duv2.y;
        T = mat_mul(r[0], r[1], r[2], T);
        B = mat_mul(r[0], r[1], r[2], B);
        float Tlen = length(T);
        if (Tlen > 0.0f && !std::isinf(Tlen) && !std::isnan(Tlen)) {
            h.tangent = T;
            h.bitangent_l = dot(normalize(cross(h.geo_normal, T)), B);
            requires_tangent = false;
        }
    }
    if (requires_tangent) {
        h.tangent = normalize(mat_mul(r[0], r[1], r[2], cross(p2 - p0, gn)));
        h.bitangent_l = 1.0f;
    }
    return h;
}

// --- host pre-passes --------------------------------------------------------------------------------------------
static inline float halton2(unsigned index) { // util/compute_util.h:19-33
    index = (index << 16) | (index >> 16);
    index = ((index & 0x00ff00ffu) << 8) | ((index & 0xff00ff00u) >> 8);
    index = ((index & 0x0f0f0f0fu) << 4) | ((index & 0xf0f0f0f0u) >> 4);
    index = ((index & 0x33333333u) << 2) | ((index & 0xccccccccu) >> 2);
    index = ((index & 0x55555555u) << 1) | ((index & 0xaaaaaaaau) >> 1);
    return u2f(0x3f800000u | (index >> 9)) - 1.0f;
}

struct Emitter { V3 v0, v1, v2, radiance; };

static inline float host_luminance(V3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; } // util/util.cpp:293-296

// librender/lights.cpp:131-166 (host variant: libm atan, not the shader polynomial)
static float host_triangle_solid_angle(V3 v0, V3 v1, V3 v2) {
    V3 prm;
    float tangent = half_tri_solid_angle_tan(v0, v1, v2, prm);
    float off = (tangent < 0.0f) ? (float)M_PI : 0.0f;
    return 2.0f * (atanf(tangent) + off);
}

// librender/lights.cpp:169-203
static std::vector<float> estimate_normalized_radiance(const std::vector<Emitter> &em, float min_dist) {
    std::vector<float> rad(em.size());
    for (size_t i = 0; i < em.size(); ++i) {
        const Emitter &l = em[i];
        V3 n = normalize(cross(l.v1 - l.v0, l.v2 - l.v0));
        if (!(fabsf(length(n) - 1.0f) < 0.05f)) { rad[i] = 0.0f; continue; }
        V3 c = (l.v0 + l.v1 + l.v2) / 3.0f;
        V3 o = n * min_dist;
        float sa = host_triangle_solid_angle(normalize(l.v0 - c - o), normalize(l.v1 - c - o), normalize(l.v2 - c - o));
        rad[i] = (float)((double)host_luminance(l.radiance) * ((double)sa / M_2_PI)); // sic: M_2_PI = 2/pi
    }
    return rad;
}

// librender/lights.cpp:220-349
static void equalize_emitter_bins(std::vector<Emitter> &emitters, std::vector<float> &radiances, int bin_size) {
    if (bin_size <= 1 || radiances.empty()) return;
    int n0 = (int)radiances.size();
    int original_bin_count = (n0 + (bin_size - 1)) / bin_size;
    float average_weight = 0.0f;
    for (int i = 0; i < n0; ++i) average_weight += radiances[i];
    average_weight /= (float)radiances.size();
    struct Bin { float radiance; int source_idx; int split_count; };
    std::vector<Bin> bins;
    bins.reserve(2 * radiances.size());
    for (int i = 0; i < n0; ++i) {
        float w = radiances[i];
        int clones = (int)std::max((unsigned)std::min(w / average_weight, (float)original_bin_count), 1u);
        for (int j = 0; j < clones; ++j) bins.push_back(Bin{radiances[i] / (float)clones, i, clones});
    }
    std::vector<Bin> shuffled;
    auto reshuffle = [&]() {
        shuffled.resize(bins.size());
        for (int index = 0, count = (int)bins.size(); index < count; ++index) {
            int src = (int)(unsigned)(halton2((unsigned)index) * (float)(unsigned)count);
            for (;;) {
                if (src >= count) src = 0;
                if (bins[src].source_idx == ~0) ++src;
                else break;
            }
            shuffled[index] = bins[src];
            bins[src].source_idx = ~0;
        }
        bins = std::move(shuffled);
        shuffled.clear();
    };
    reshuffle();
    auto measure = [bin_size](const std::vector<Bin> &b) {
        float mn = 2.0e32f, mx = 0.0f;
        for (int i = 0, ie = (int)b.size(); i < ie;) {
            float tot = 0.0f;
            for (int j = 0; j < bin_size && i < ie; ++i, ++j) tot += b[i].radiance;
            mn = std::min(tot, mn);
            mx = std::max(tot, mx);
        }
        return std::min(mn / mx, 1.0f);
    };
    float equality = measure(bins);
    std::vector<Bin> postfix;
    for (int retries = 0; equality < 0.6f && retries < 2; ++retries) {
        postfix.resize(bins.size());
        Bin acc = bins[0];
        postfix[0] = acc;
        for (size_t i = 1; i < bins.size(); ++i) {
            acc = Bin{acc.radiance + bins[i].radiance, bins[i].source_idx, 1};
            postfix[i] = acc;
        }
        postfix.front().split_count = 1;
        float sum = postfix.back().radiance;
        for (auto &it : postfix) it.radiance /= sum;
        int prev_elements = (int)bins.size();
        int prev_bin_count = (prev_elements + (bin_size - 1)) / bin_size;
        int padded = (prev_bin_count + 1) * bin_size;
        unsigned hidx = 0;
        while ((int)bins.size() < padded) {
            float u = halton2(hidx++);
            auto it = std::upper_bound(postfix.begin(), postfix.end(), u, [](float bound, const Bin &b) { return bound < b.radiance; });
            if (it == postfix.end()) it = postfix.end() - 1;
            ++it->split_count;
            bins.push_back(Bin{it->radiance, (int)(it - postfix.begin()), 0});
        }
        for (int i = prev_elements; i < padded; ++i) {
            Bin &clone = bins[i];
            Bin &orig = bins[clone.source_idx];
            int &counter = postfix[clone.source_idx].split_count;
            if (counter > 1) {
                orig.radiance /= (float)counter;
                orig.split_count *= counter;
                counter = 1;
            }
            clone.radiance = orig.radiance;
            clone.source_idx = orig.source_idx;
            clone.split_count = orig.split_count;
        }
        reshuffle();
        equality = measure(bins);
    }
    std::vector<Emitter> out(bins.size());
    radiances.resize(bins.size());
    for (size_t i = 0; i < bins.size(); ++i) {
        radiances[i] = bins[i].radiance;
        out[i] = emitters[bins[i].source_idx];
        out[i].radiance = out[i].radiance / (float)bins[i].split_count;
    }
    emitters = std::move(out);
}

} // namespace

// ================================================================================================================
// C API (ctypes-callable)
// ================================================================================================================
// ---- texture ingestion (util/image.h:10-27: all mip levels of an image back to back; block formats of librender/scene.cpp:836-930) ----
// S3TC / RGTC blocks after the Khronos Data Format Specification; interpolated palette entries are the specified rationals rounded
// to the nearest 8-bit code.  Written independently of the product's decoder (csrc/rptr_host.cpp).
static void palette_bc1(const uint8_t *blk, bool four_colour_always, bool punch_through, uint8_t pal[4][4]) {
    const unsigned e0 = blk[0] + 256u * blk[1], e1 = blk[2] + 256u * blk[3];
    const unsigned ends[2] = {e0, e1};
    for (int k = 0; k < 2; ++k) {
        const unsigned r5 = ends[k] >> 11, g6 = (ends[k] >> 5) & 0x3fu, b5 = ends[k] & 0x1fu;
        pal[k][0] = (uint8_t)(r5 * 8 + r5 / 4); pal[k][1] = (uint8_t)(g6 * 4 + g6 / 16); pal[k][2] = (uint8_t)(b5 * 8 + b5 / 4); pal[k][3] = 255;
    }
    pal[2][3] = pal[3][3] = 255;
    for (int c = 0; c < 3; ++c) {
        const int a = pal[0][c], b = pal[1][c];
        if (e0 > e1 || four_colour_always) {
            pal[2][c] = (uint8_t)((2 * a + b + 1) / 3); // round(2a/3 + b/3)
            pal[3][c] = (uint8_t)((a + 2 * b + 1) / 3);
        } else {
            pal[2][c] = (uint8_t)((a + b + 1) / 2);
            pal[3][c] = 0;
        }
    }
    if (!(e0 > e1 || four_colour_always) && punch_through) pal[3][3] = 0;
}
static void palette_bc4(const uint8_t *blk, uint8_t pal[8]) {
    const int a = blk[0], b = blk[1];
    pal[0] = (uint8_t)a; pal[1] = (uint8_t)b;
    if (a > b) {
        for (int k = 2; k < 8; ++k) pal[k] = (uint8_t)(((8 - k) * a + (k - 1) * b + 3) / 7);
    } else {
        for (int k = 2; k < 6; ++k) pal[k] = (uint8_t)(((6 - k) * a + (k - 1) * b + 2) / 5);
        pal[6] = 0; pal[7] = 255;
    }
}
static unsigned bc4_index(const uint8_t *blk, int texel) { // 3 bits per texel, little endian, starting at byte 2
    const int bit = 3 * texel;
    const unsigned lo = blk[2 + bit / 8], hi = bit / 8 + 1 < 6 ? blk[3 + bit / 8] : 0u;
    return ((lo | (hi << 8)) >> (bit % 8)) & 7u;
}
static bool texture_levels_to_rgba8(const rptr_texture_desc &td, std::vector<uint8_t> &out) {
    const int fmt = td.bc_format;
    if (fmt != 0 && fmt != 1 && fmt != -1 && fmt != 3 && fmt != 5) return false;
    const int levels = td.mip_levels > 0 ? td.mip_levels : 1;
    out.clear();
    const uint8_t *in = td.texels;
    int w = td.width, h = td.height;
    for (int level = 0; level < levels; ++level) {
        if (fmt == 0) { // kept in its own channel count (the texture set reads `channels` bytes per texel)
            out.insert(out.end(), in, in + (size_t)w * h * td.channels);
            in += (size_t)w * h * td.channels;
        } else {
            const size_t at = out.size();
            out.resize(at + (size_t)4 * w * h);
            const int blocks_x = (w + 3) / 4, blocks_y = (h + 3) / 4, bytes = (fmt == 3 || fmt == 5) ? 16 : 8;
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x) {
                    const uint8_t *blk = in + ((size_t)(y / 4) * blocks_x + x / 4) * bytes;
                    const int texel = 4 * (y % 4) + x % 4;
                    uint8_t *px = out.data() + at + 4 * ((size_t)y * w + x);
                    if (fmt == 5) {
                        uint8_t pr[8], pg[8];
                        palette_bc4(blk, pr); palette_bc4(blk + 8, pg);
                        px[0] = pr[bc4_index(blk, texel)]; px[1] = pg[bc4_index(blk + 8, texel)]; px[2] = 0; px[3] = 255;
                    } else {
                        const uint8_t *colour = fmt == 3 ? blk + 8 : blk;
                        uint8_t pal[4][4];
                        palette_bc1(colour, fmt == 3, fmt == -1, pal);
                        const unsigned code = (colour[4 + texel / 4] >> (2 * (texel % 4))) & 3u;
                        px[0] = pal[code][0]; px[1] = pal[code][1]; px[2] = pal[code][2]; px[3] = pal[code][3];
                        if (fmt == 3) { uint8_t pa[8]; palette_bc4(blk, pa); px[3] = pa[bc4_index(blk, texel)]; }
                    }
                }
            in += (size_t)blocks_x * blocks_y * bytes;
        }
        w = w > 1 ? w / 2 : 1;
        h = h > 1 ? h / 2 : 1;
    }
    return true;
}

extern "C" {

typedef struct oracle_render_args {
    int32_t width, height;
    rptr_camera_params camera;
    rptr_render_params params;
    rptr_light_sampling_config lighting;
    rptr_scene_params scene_params; // sun_radiance[3] as produced by the sky fit BEFORE the light-count rule
    uint32_t frame_offset;
    uint32_t first_sample;  // frame_id of the first frame
    int32_t n_samples;      // number of frames (batch_spp = 1 each)
    int32_t x0, y0, x1, y1; // pixel region to render (x1/y1 exclusive)
    int32_t transmission;   // 0 = megakernel build (no GLTF_SUPPORT_TRANSMISSION)
    int32_t n_threads;      // 0 = all
    int32_t rng_variant;    // RenderBackendOptions::rng_variant: 0 UNIFORM, 1 BN, 2 SOBOL, 3 Z_SBL
    int32_t batch_spp;      // >= 1: view_params.frame_id of sample k is first_sample + (k / batch_spp) * batch_spp (the BN
                            // sampler seeds from frame_id, not from the sample index: bn_rng.glsl:112)
    const uint32_t *pointset_tables[4]; // SobolMatrix, SobolInversion_1_0, sobol_256spp_256d, scramblingTile_yx_d_1spp
    float vp_reference[16];             // view_params.VP_reference (column-major): the VP of the previous begin_frame; all zero
                                        // before the first one (the value-initialised ParameterCache, render_vulkan.cpp:103)
} oracle_render_args;

struct oracle_scene { Scene s; };

// Host inverse of the 3x3 part, rows of the inverse = cross products of columns / det.
static void inverse_rows(const float *m, V3 *rows) {
    V3 c0 = v3(m[0], m[4], m[8]), c1 = v3(m[1], m[5], m[9]), c2 = v3(m[2], m[6], m[10]);
    V3 r0 = cross(c1, c2), r1 = cross(c2, c0), r2 = cross(c0, c1);
    float det = dot(c0, r0);
    float inv = 1.0f / det;
    rows[0] = r0 * inv;
    rows[1] = r1 * inv;
    rows[2] = r2 * inv;
}

oracle_scene *oracle_scene_create(const rptr_scene_desc *d, const rptr_light_sampling_config *ls) {
    oracle_scene *os = new oracle_scene();
    Scene &s = os->s;
    s.materials.assign(d->materials, d->materials + d->n_materials);
    s.own_texels.resize(d->textures ? d->n_textures : 0);
    for (int t = 0; t < (int)s.own_texels.size(); ++t) {
        rptr_texture_desc td = d->textures[t];
        if (td.width < 1 || td.height < 1 || !td.texels || (td.bc_format == 0 && (td.channels < 1 || td.channels > 4))) { delete os; return nullptr; }
        if (!texture_levels_to_rgba8(td, s.own_texels[t])) { delete os; return nullptr; }
        if (td.bc_format != 0) { td.bc_format = 0; td.channels = 4; } // decoded: RGBA8 levels back to back
        td.texels = s.own_texels[t].data();
        s.textures.push_back(td);
    }
    s.texset.tex = s.textures.data();
    s.texset.n = (int)s.textures.size();
    // copy the streams
    s.own_qv.resize(d->n_geometries);
    s.own_qn.resize(d->n_geometries);
    for (int g = 0; g < d->n_geometries; ++g) {
        const rptr_geometry_desc &gd = d->geometries[g];
        s.own_qv[g].assign(gd.qverts, gd.qverts + 3 * (size_t)gd.n_tris);
        if (gd.qnormal_uv && (gd.has_normals || gd.has_uvs)) s.own_qn[g].assign(gd.qnormal_uv, gd.qnormal_uv + 3 * (size_t)gd.n_tris);
    }
    s.own_tm.resize(d->n_pmeshes);
    for (int p = 0; p < d->n_pmeshes; ++p)
        if (d->pmeshes[p].tri_material_ids)
            s.own_tm[p].assign(d->pmeshes[p].tri_material_ids, d->pmeshes[p].tri_material_ids + d->pmeshes[p].n_tri_material_ids);

    std::vector<Emitter> emitters;
    std::vector<char> pmesh_nonemissive(d->n_pmeshes, 0);
    for (int i = 0; i < d->n_instances; ++i) {
        const rptr_instance_desc &inst = d->instances[i];
        const rptr_pmesh_desc &pm = d->pmeshes[inst.pmesh_id];
        const rptr_mesh_desc &mesh = d->meshes[pm.mesh_id];
        bool per_tri = pm.tri_material_ids != nullptr;
        // vulkan/render_vulkan.cpp:1116-1131: pmesh no_alpha = all its materials are NOALPHA
        int64_t prim_offset = 0;
        std::vector<Emitter> next;
        for (int j = 0; j < mesh.n_geometries; ++j) {
            int g = mesh.first_geometry + j;
            const rptr_geometry_desc &gd = d->geometries[g];
            GeomInst gi;
            gi.qverts = s.own_qv[g].data();
            gi.qnuv = s.own_qn[g].empty() ? nullptr : s.own_qn[g].data();
            for (int k = 0; k < 3; ++k) { gi.scale[k] = gd.quantized_scaling[k]; gi.offset[k] = gd.quantized_offset[k]; }
            gi.n_tris = gd.n_tris;
            gi.has_normals = gd.has_normals && gi.qnuv;
            gi.has_uvs = gd.has_uvs && gi.qnuv;
            int mat_off = pm.n_material_offsets ? pm.material_offsets[j] : 0;
            gi.flags = RPTR_GEOMETRY_FLAGS_IMPLICIT_INDICES;
            gi.tri_mat = nullptr;
            bool no_alpha;
            if (per_tri) { // render_vulkan.cpp:2812-2818
                gi.material_id = -1 - mat_off;
                gi.tri_mat = s.own_tm[inst.pmesh_id].data() + prim_offset;
                gi.flags |= RPTR_GEOMETRY_FLAGS_EXTENDED_SHADER;
                no_alpha = true;
                for (int t = 0; t < gd.n_tris; ++t)
                    if (!(s.materials[mat_off + gi.tri_mat[t]].flags & RPTR_BASE_MATERIAL_NOALPHA)) { no_alpha = false; break; }
            } else {
                gi.material_id = mat_off;
                no_alpha = (s.materials[mat_off].flags & RPTR_BASE_MATERIAL_NOALPHA) != 0;
                if (s.materials[mat_off].flags & RPTR_BASE_MATERIAL_EXTENDED) gi.flags |= RPTR_GEOMETRY_FLAGS_EXTENDED_SHADER;
                if (!(s.materials[mat_off].flags & RPTR_BASE_MATERIAL_ONESIDED)) gi.flags |= RPTR_GEOMETRY_FLAGS_THIN;
            }
            if (no_alpha) gi.flags |= RPTR_GEOMETRY_FLAGS_NOALPHA;
            else s.any_non_opaque = true;
            gi.instance = i;
            std::memcpy(gi.o2w, inst.transform, sizeof(gi.o2w));
            inverse_rows(gi.o2w, gi.w2o_row);
            gi.first_tri = (int64_t)s.tris.size();
            int gidx = (int)s.ginst.size();
            for (int t = 0; t < gd.n_tris; ++t) {
                V3 a = xfm_point(gi.o2w, dequantize_position(gi.qverts[3 * (size_t)t + 0], gi.scale, gi.offset));
                V3 b = xfm_point(gi.o2w, dequantize_position(gi.qverts[3 * (size_t)t + 1], gi.scale, gi.offset));
                V3 c = xfm_point(gi.o2w, dequantize_position(gi.qverts[3 * (size_t)t + 2], gi.scale, gi.offset));
                s.tris.push_back(Tri{a, b - a, c - a, gidx, t});
                // collect_emitters (librender/lights.cpp:33-73)
                if (!pmesh_nonemissive[inst.pmesh_id]) {
                    int mid = per_tri ? mat_off + gi.tri_mat[t] : mat_off;
                    const rptr_base_material &m = s.materials[mid];
                    if (m.emission_intensity > 0.0f) {
                        V3 rad = v3(m.base_color[0], m.base_color[1], m.base_color[2]) * m.emission_intensity;
                        next.push_back(Emitter{a, b, c, rad});
                    }
                }
            }
            s.ginst.push_back(gi);
            prim_offset += gd.n_tris;
        }
        if (!pmesh_nonemissive[inst.pmesh_id]) { // lights.cpp:17-30: new emitters are PREPENDED
            if (!next.empty()) emitters.insert(emitters.begin(), next.begin(), next.end());
            else pmesh_nonemissive[inst.pmesh_id] = 1;
        }
    }
    if (d->binned_lights) {
        s.lights.assign(d->binned_lights, d->binned_lights + d->n_binned_lights);
    } else if (!emitters.empty()) { // update_light_sampling, lights.cpp:75-90
        std::vector<float> rad = estimate_normalized_radiance(emitters, ls->min_perceived_receiver_dist);
        if (ls->min_radiance > 0.0f) {
            size_t n = 0;
            for (size_t i = 0; i < emitters.size(); ++i)
                if (rad[i] >= ls->min_radiance) { emitters[n] = emitters[i]; rad[n] = rad[i]; ++n; }
            emitters.resize(n);
            rad.resize(n);
        }
        equalize_emitter_bins(emitters, rad, ls->bin_size);
        for (const Emitter &e : emitters) {
            rptr_tri_light_data t;
            t.v0[0] = e.v0.x; t.v0[1] = e.v0.y; t.v0[2] = e.v0.z;
            t.v1[0] = e.v1.x; t.v1[1] = e.v1.y; t.v1[2] = e.v1.z;
            t.v2[0] = e.v2.x; t.v2[1] = e.v2.y; t.v2[2] = e.v2.z;
            t.radiance[0] = e.radiance.x; t.radiance[1] = e.radiance.y; t.radiance[2] = e.radiance.z;
            s.lights.push_back(t);
        }
    }
    build_bvh(s);
    return os;
}

void oracle_scene_destroy(oracle_scene *s) { delete s; }
int64_t oracle_scene_num_tris(const oracle_scene *s) { return (int64_t)s->s.tris.size(); }
int32_t oracle_scene_num_lights(const oracle_scene *s) { return (int32_t)s->s.lights.size(); }
void oracle_scene_get_lights(const oracle_scene *s, rptr_tri_light_data *out) {
    std::memcpy(out, s->s.lights.data(), s->s.lights.size() * sizeof(rptr_tri_light_data));
}

// vulkan/render_vulkan.cpp:2880-2894 -> out = du(3), dv(3), top_left(3)
void oracle_view_params(const rptr_camera_params *cam, int32_t w, int32_t h, float *out) {
    V3 dir = v3(cam->dir[0], cam->dir[1], cam->dir[2]), up = v3(cam->up[0], cam->up[1], cam->up[2]);
    float py = 2.0f * tanf(0.5f * cam->fovy * 0.01745329251994329576923690768489f);
    float aspect = (float)w / (float)h;
    float px = py * aspect;
    V3 du = normalize(cross(dir, up)) * px;
    V3 dv = -normalize(cross(du, dir)) * py;
    V3 tl = dir - du * 0.5f - dv * 0.5f;
    out[0] = du.x; out[1] = du.y; out[2] = du.z;
    out[3] = dv.x; out[4] = dv.y; out[5] = dv.z;
    out[6] = tl.x; out[7] = tl.y; out[8] = tl.z;
}

// view_params.VP (vulkan/render_vulkan.cpp:2926-2930) = GLToVulkan * infinitePerspective(radians(fovy), aspect, 0.5) *
// inverse(mat4(mat4x3(cross(dir, up), up, -dir, cam_pos))), out = 16 floats, column-major.  glm (0.9.9.8, fetched at
// configure time by ext/CMakeLists.txt:18-21, absent from the reference tree) is restated from its published sources:
// type_mat4x4.inl (operator*), func_matrix.inl (compute_inverse<4, 4>), ext/matrix_clip_space.inl (infinitePerspectiveRH).
} // extern "C"
namespace {
struct M4 { V4 col[4]; };
static inline V4 operator*(V4 a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }
static inline V4 operator*(V4 a, V4 b) { return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
static inline V4 operator+(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
static inline V4 operator-(V4 a, V4 b) { return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
static inline float comp(V4 v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
static M4 glm_mul(const M4 &a, const M4 &b) {
    M4 r;
    for (int j = 0; j < 4; ++j) r.col[j] = a.col[0] * b.col[j].x + a.col[1] * b.col[j].y + a.col[2] * b.col[j].z + a.col[3] * b.col[j].w;
    return r;
}
static M4 glm_inverse(const M4 &m) {
#define E(c, r) comp(m.col[c], r)
    float Coef00 = E(2, 2) * E(3, 3) - E(3, 2) * E(2, 3), Coef02 = E(1, 2) * E(3, 3) - E(3, 2) * E(1, 3), Coef03 = E(1, 2) * E(2, 3) - E(2, 2) * E(1, 3);
    float Coef04 = E(2, 1) * E(3, 3) - E(3, 1) * E(2, 3), Coef06 = E(1, 1) * E(3, 3) - E(3, 1) * E(1, 3), Coef07 = E(1, 1) * E(2, 3) - E(2, 1) * E(1, 3);
    float Coef08 = E(2, 1) * E(3, 2) - E(3, 1) * E(2, 2), Coef10 = E(1, 1) * E(3, 2) - E(3, 1) * E(1, 2), Coef11 = E(1, 1) * E(2, 2) - E(2, 1) * E(1, 2);
    float Coef12 = E(2, 0) * E(3, 3) - E(3, 0) * E(2, 3), Coef14 = E(1, 0) * E(3, 3) - E(3, 0) * E(1, 3), Coef15 = E(1, 0) * E(2, 3) - E(2, 0) * E(1, 3);
    float Coef16 = E(2, 0) * E(3, 2) - E(3, 0) * E(2, 2), Coef18 = E(1, 0) * E(3, 2) - E(3, 0) * E(1, 2), Coef19 = E(1, 0) * E(2, 2) - E(2, 0) * E(1, 2);
    float Coef20 = E(2, 0) * E(3, 1) - E(3, 0) * E(2, 1), Coef22 = E(1, 0) * E(3, 1) - E(3, 0) * E(1, 1), Coef23 = E(1, 0) * E(2, 1) - E(2, 0) * E(1, 1);
    V4 Fac0{Coef00, Coef00, Coef02, Coef03}, Fac1{Coef04, Coef04, Coef06, Coef07}, Fac2{Coef08, Coef08, Coef10, Coef11};
    V4 Fac3{Coef12, Coef12, Coef14, Coef15}, Fac4{Coef16, Coef16, Coef18, Coef19}, Fac5{Coef20, Coef20, Coef22, Coef23};
    V4 Vec0{E(1, 0), E(0, 0), E(0, 0), E(0, 0)}, Vec1{E(1, 1), E(0, 1), E(0, 1), E(0, 1)};
    V4 Vec2{E(1, 2), E(0, 2), E(0, 2), E(0, 2)}, Vec3{E(1, 3), E(0, 3), E(0, 3), E(0, 3)};
#undef E
    V4 Inv0 = Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2;
    V4 Inv1 = Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4;
    V4 Inv2 = Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5;
    V4 Inv3 = Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5;
    V4 SignA{+1.0f, -1.0f, +1.0f, -1.0f}, SignB{-1.0f, +1.0f, -1.0f, +1.0f};
    M4 Inverse{{Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB}};
    V4 Row0{Inverse.col[0].x, Inverse.col[1].x, Inverse.col[2].x, Inverse.col[3].x};
    V4 Dot0 = m.col[0] * Row0;
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.0f / Dot1;
    for (int j = 0; j < 4; ++j) Inverse.col[j] = Inverse.col[j] * OneOverDeterminant;
    return Inverse;
}
} // namespace
extern "C" {
void oracle_view_projection(const rptr_camera_params *cam, int32_t w, int32_t h, float *out) {
    V3 dir = v3(cam->dir[0], cam->dir[1], cam->dir[2]), up = v3(cam->up[0], cam->up[1], cam->up[2]);
    V3 right_ = cross(dir, up);
    M4 frame{{V4{right_.x, right_.y, right_.z, 0.0f}, V4{up.x, up.y, up.z, 0.0f}, V4{-dir.x, -dir.y, -dir.z, 0.0f},
              V4{cam->pos[0], cam->pos[1], cam->pos[2], 1.0f}}};
    float fovy = cam->fovy * 0.01745329251994329576923690768489f; // glm::radians
    float aspect = (float)w / (float)h;
    float zNear = 0.5f;
    float range = tanf(fovy / 2.0f) * zNear;
    float left = -range * aspect, right = range * aspect, bottom = -range, top = range;
    M4 P{{V4{(2.0f * zNear) / (right - left), 0.0f, 0.0f, 0.0f}, V4{0.0f, (2.0f * zNear) / (top - bottom), 0.0f, 0.0f}, V4{0.0f, 0.0f, -1.0f, -1.0f},
          V4{0.0f, 0.0f, -2.0f * zNear, 0.0f}}};
    M4 GLToVulkan{{V4{1.0f, 0.0f, 0.0f, 0.0f}, V4{0.0f, -1.0f, 0.0f, 0.0f}, V4{0.0f, 0.0f, 0.5f, 0.0f}, V4{0.0f, 0.0f, 0.5f, 1.0f}}};
    M4 VP = glm_mul(glm_mul(GLToVulkan, P), glm_inverse(frame));
    for (int j = 0; j < 4; ++j) { out[4 * j] = VP.col[j].x; out[4 * j + 1] = VP.col[j].y; out[4 * j + 2] = VP.col[j].z; out[4 * j + 3] = VP.col[j].w; }
}

} // extern "C"

namespace {

// Raster-TAA screen jitter (vulkan/render_vulkan.cpp:2917-2926): entry (frame_offset + frame_id) % RASTER_TAA_NUM_SAMPLES (16,
// CMakeLists.txt:30) of the 2-3 Halton table of librender/halton.h, whose entries are 6-decimal literals:
// halton_23[k] = (phi_2(k + 1), phi_3(k + 1)) rounded to 6 decimals, then to float.
static float halton_literal(int index, int base) {
    double f = 1.0, r = 0.0;
    for (int i = index; i > 0; i /= base) {
        f /= base;
        r += f * (i % base);
    }
    // the table spells the value the way printf("%.6f") does (exact ties to even: phi_2(64) = 0.0078125 is listed as 0.007812)
    char buf[32];
    std::snprintf(buf, sizeof(buf), "%.6f", r);
    return std::strtof(buf, nullptr);
}
static void screen_jitter(uint32_t frame_offset, uint32_t frame_id, int w, int h, float *out) {
    const uint32_t idx = (frame_offset + frame_id) % 16u;
    const float hx = halton_literal((int)idx + 1, 2), hy = halton_literal((int)idx + 1, 3);
    out[0] = hx * 2.0f / (float)w - 1.0f / (float)w;
    out[1] = hy * 2.0f / (float)h - 1.0f / (float)h;
}

struct Frame {
    const Scene *s;
    const oracle_render_args *a;
    rptr_scene_params sp; // with the light-count rule applied to sun_radiance.w
    V3 cam_pos, du, dv, tl;
    float VP[16]; // view_params.VP of this frame, column-major
    int n_lights, n_bins;
    bool tr;
};

// counters for the roofline bookkeeping the bench reports (path vertices / shadow rays per sample)
struct Counters { uint64_t closest_rays = 0, shadow_rays = 0, vertices = 0; };

// sample_tri_lights: rendering/mc/lights_linear.glsl:19-127 (BINNED_LIGHTS_BIN_MAX_SIZE = 16)
static V3 sample_tri_lights(const Frame &f, V3 hit_p, V3 hit_n, V2 dir_sample, V2 sel, V3 &light_dir, float &light_dist,
                            float &pdf, float &mis_wpdf) {
    int num_lights = f.n_lights;
    int num_bins = f.n_bins;
    int bin_size = f.a->lighting.bin_size;
    sel.x *= (float)num_bins;
    int bin_id = (int)(uint32_t)sel.x;
    bin_id = std::min(bin_id, num_bins - 1);
    float sel_p = 1.0f / (float)num_bins;
    sel.x -= (float)bin_id;
    float contribs[RPTR_BINNED_LIGHTS_BIN_MAX_SIZE];
    float total = 0.0f;
    const float MIN_IRRADIANCE = 6.2e-4f * 0.001f;
    int bin_end = std::min(bin_size * (bin_id + 1), num_lights);
    for (int i = 0; i < RPTR_BINNED_LIGHTS_BIN_MAX_SIZE; ++i) {
        int light_id = bin_size * bin_id + i;
        if (!(light_id < bin_end)) break;
        const rptr_tri_light_data &L = f.s->lights[light_id];
        V3 a = v3(L.v0[0], L.v0[1], L.v0[2]) - hit_p, b = v3(L.v1[0], L.v1[1], L.v1[2]) - hit_p, c = v3(L.v2[0], L.v2[1], L.v2[2]) - hit_p;
        bool front = dot(cross(a, b), c) < 0.0f; // is_tri_facing_forward, lights/tri.glsl:23-25
        float contrib = luminance(v3(L.radiance[0], L.radiance[1], L.radiance[2]));
        if ((dot(a, hit_n) > 0.0f || dot(b, hit_n) > 0.0f || dot(c, hit_n) > 0.0f) && front) {
            V3 prm;
            contrib *= triangle_solid_angle(normalize(a), normalize(b), normalize(c), prm);
        } else
            contrib = 0.0f;
        contrib += MIN_IRRADIANCE;
        contribs[i] = contrib;
        total += contrib;
    }
    float p = 0.0f, t = 0.0f;
    int light_id = 0;
    for (int i = 0; i < RPTR_BINNED_LIGHTS_BIN_MAX_SIZE; ++i) {
        light_id = bin_size * bin_id + i;
        if (!(light_id < bin_end)) break;
        p = contribs[i] / total;
        t += p;
        if (sel.y < t) break;
    }
    // note: when the loop runs off the end of the bin, the reference indexes one light past it; with the bin
    // padded by equalize_emitter_bins this cannot go out of bounds except in the last bin -> clamp there.
    light_id = std::min(light_id, num_lights - 1);
    sel_p *= p;
    const rptr_tri_light_data &L = f.s->lights[light_id];
    V3 v0 = v3(L.v0[0], L.v0[1], L.v0[2]), v1 = v3(L.v1[0], L.v1[1], L.v1[2]), v2 = v3(L.v2[0], L.v2[1], L.v2[2]);
    V3 d0 = normalize(v0 - hit_p), d1 = normalize(v1 - hit_p), d2 = normalize(v2 - hit_p);
    V3 prm;
    float omega = triangle_solid_angle(d0, d1, d2, prm);
    light_dir = sample_solid_angle_polygon(d0, d1, d2, omega, prm, dir_sample);
    pdf = 1.0f / omega;
    V3 e_n = cross(v1 - v0, v2 - v0);
    light_dist = dot(v0 - hit_p, e_n) / dot(light_dir, e_n);
    mis_wpdf = 2.0f * light_dist * light_dist / fabsf(dot(light_dir, e_n));
    pdf *= sel_p;
    mis_wpdf /= (float)num_bins;
    return v3(L.radiance[0], L.radiance[1], L.radiance[2]) / pdf;
}

static inline float geometry_scale_to_tmin(V3 orig, float scale) { return (length(orig) + scale) * 0.000005f; } // vulkan/geometry.glsl:76-78


// pieces of raytrace_test_visibility / generate_candidate_hit (vulkan/pt_megakernel.glsl:216-272, 202-210)
// range of a shadow ray: (eps, dist - eps) with eps = geometry_scale_to_tmin; not cast at all (= visible) when dist - 2 eps <= 0
static bool shadow_ray_range(V3 from, float dist, float geometry_scale, float &tmin, float &tmax) {
    float eps = geometry_scale_to_tmin(from, geometry_scale);
    if (dist - 2.0f * eps > 0.0f) {
        tmin = eps;
        tmax = dist - eps;
        return true;
    }
    return false;
}
// every candidate of a shadow ray gets its own LCG (:252-254)
static Lcg shadow_alpha_rng(uint32_t prim, uint32_t instance, uint32_t frame_id, uint32_t frame_offset, uint32_t pixel_linear) {
    return lcg_seed(prim ^ frame_id, instance ^ frame_offset, pixel_linear);
}
// stochastic alpha test of a candidate (:205-207): true = rejected, the ray passes
static bool alpha_test_rejects(float alpha, Lcg &rng) { return !(alpha > 0.0f) || (alpha < 1.0f && lcg_randomf(rng) > alpha); }

// raytrace_test_visibility: vulkan/pt_megakernel.glsl:216-272
static bool test_visibility(const Frame &f, V3 from, V3 dir, float dist, float geometry_scale, uint32_t pixel_linear,
                            uint32_t frame_id, Counters &cnt) {
    float tmin, tmax;
    if (shadow_ray_range(from, dist, geometry_scale, tmin, tmax)) {
        cnt.shadow_rays++;
        const Scene &s = *f.s;
        bool occluded = any_hit(s, from, dir, tmin, tmax, [&](int id, float t, float bu, float bv) {
            const Tri &tr = s.tris[id];
            const GeomInst &g = s.ginst[tr.geom_inst];
            if (g.flags & RPTR_GEOMETRY_FLAGS_NOALPHA) return true;
            const rptr_base_material &m = s.materials[calc_hit_material_id(g, (uint32_t)tr.prim)];
            if (m.flags & RPTR_BASE_MATERIAL_NOALPHA) return true;
            Lcg arng = shadow_alpha_rng((uint32_t)tr.prim, (uint32_t)g.instance, frame_id, f.a->frame_offset, pixel_linear);
            // generate_candidate_hit (:153-211): the candidate's own hit attributes give the uv of the alpha lookup
            const V2 uv = TextureSet::is_handle(m.base_color[0]) ? calc_hit_attributes(g, t, (uint32_t)tr.prim, bu, bv).uv : V2{0.0f, 0.0f};
            return !alpha_test_rejects(material_alpha(s.texset, m, uv), arng);
        });
        return !occluded;
    }
    return true;
}

// main_spp: vulkan/pt_megakernel.glsl:310-737, shade_base_material (rendering/mc/shade_base_material.glsl:14-96),
// sample_direct_light (rendering/mc/nee.glsl:32-90).  Returns vec4(illum, bounce == 0 ? 0 : 1).
// sample_direct_light: rendering/mc/nee.glsl:32-90 with sample_sun_light (mc/lights_sun.glsl:8-17, lights/sun.glsl:9-20) and
// sample_tri_lights; `visible(from, dir, dist)` is raytrace_test_visibility.  contrib excludes the path throughput.
struct NeeSample { V3 contrib, light_dir; float light_dist, mis_pdf; };
template <class Vis>
static NeeSample sample_direct_light(const Frame &f, const GltfMat &mat, V3 ip, V3 ign, V3 in_, V3 w_o, V2 dir_sample, V2 sel_sample, Vis &&visible) {
    const rptr_scene_params &sp = f.sp;
    const float p_sun = sp.sun_radiance[3];
    V3 sun_dir = v3(sp.sun_dir[0], sp.sun_dir[1], sp.sun_dir[2]);
    V3 sun_rgb = v3(sp.sun_radiance[0], sp.sun_radiance[1], sp.sun_radiance[2]);
    V3 li = v3(0.0f), light_dir = v3(0.0f);
    float light_dist = 2.e16f, light_pdf = 0.0f, mis_pdf = 0.0f;
    if (sel_sample.x <= p_sun) {
        sel_sample.x /= p_sun;
        float sn, cs;
        sincos_pos(TWO_PI_F * dir_sample.x, sn, cs);
        float cosT = mix(1.0f, sp.sun_cos_angle, dir_sample.y);
        float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
        V3 fx, fy;
        ortho_basis(fx, fy, sun_dir);
        light_dir = mat_mul(fx, fy, sun_dir, v3(sinT * cs, sinT * sn, cosT));
        float pdf = 1.0f / (TWO_PI_F * (1.0f - sp.sun_cos_angle));
        li = li + (v3(1.0f) / pdf) * (sun_rgb / p_sun);
        light_pdf = pdf * p_sun;
        mis_pdf = light_pdf;
    } else {
        sel_sample.x = (sel_sample.x - p_sun) / (1.0f - p_sun);
        float tri_mis = 0.0f;
        li = li + sample_tri_lights(f, ip, in_, dir_sample, sel_sample, light_dir, light_dist, light_pdf, tri_mis) / (1.0f - p_sun);
        light_pdf *= 1.0f - p_sun;
        mis_pdf = tri_mis * (1.0f - p_sun);
    }
    NeeSample r;
    r.contrib = v3(0.0f); r.light_dir = light_dir; r.light_dist = light_dist; r.mis_pdf = 0.0f;
    if (light_pdf > 0.0f && dot(light_dir, ign) * dot(light_dir, in_) > 0.0f) {
        bool vis = visible(ip, light_dir, light_dist);
        float bsdf_pdf = gltf_wpdf(mat, in_, w_o, light_dir, f.tr);
        if (bsdf_pdf >= 0.0f && vis) {
            V3 bsdf = gltf_bsdf(mat, in_, w_o, light_dir, f.tr);
            float w = nee_mis_heuristic(1.0f, mis_pdf, 1.0f, bsdf_pdf);
            r.contrib = li * (bsdf * (w * fabsf(dot(light_dir, in_))));
            r.mis_pdf = mis_pdf;
        }
    }
    return r;
}

// RANDOM_STATE of the selected pointset (rendering/pointsets/selected_rng.glsl) + the separate LCG the megakernel keeps
// for stochastic alpha when the pointset is not the LCG (pt_megakernel.glsl:354-358)
struct PathRng {
    oracle_ps::QmcRng q;
    Lcg lcg, alpha;
    bool qmc = false;
    float draw(int d) { return qmc ? q.draw(d) : lcg_randomf(lcg); }
    float draw_alpha() { return qmc ? lcg_randomf(alpha) : lcg_randomf(lcg); }
    void set_dim(int d) { q.set_dim(d); }
    void shift_dim(int d) { q.shift_dim(d); }
};

// what the megakernel imageStore()s into aov_albedo_roughness_buffer / aov_normal_depth_buffer for the first path vertex
// (vulkan/accumulate.glsl:77-103), as floats: [0..3] = albedo.rgb, roughness; [4..7] = normal.xyz, depth;
// [8..11] = motion.xy, screen_jitter.xy (aov_motion_jitter_buffer)
struct AovOut { float v[12]; uint32_t view_frame_id; };
// store_motion_jitter_aovs: vulkan/accumulate.glsl:77-87.  mat4 * vec4 sums the columns left to right (RPTR-FP).
static V4 mat_vec(const float *M, V4 v) {
    V4 r;
    r.x = ((M[0] * v.x + M[4] * v.y) + M[8] * v.z) + M[12] * v.w;
    r.y = ((M[1] * v.x + M[5] * v.y) + M[9] * v.z) + M[13] * v.w;
    r.z = ((M[2] * v.x + M[6] * v.y) + M[10] * v.z) + M[14] * v.w;
    r.w = ((M[3] * v.x + M[7] * v.y) + M[11] * v.z) + M[15] * v.w;
    return r;
}
static float glsl_max(float x, float y) { return x < y ? y : x; }
static void store_motion_jitter_aovs(const Frame &f, V3 position, V3 motion_vector, AovOut *aov) {
    const oracle_render_args &a = *f.a;
    V3 moved = position + motion_vector;
    V4 ref_proj = mat_vec(a.vp_reference, V4{moved.x, moved.y, moved.z, 1.0f});
    float ref_w = glsl_max(ref_proj.w, 0.0f);
    V4 cur_proj = mat_vec(f.VP, V4{position.x, position.y, position.z, 1.0f});
    float cur_w = glsl_max(cur_proj.w, 0.0f);
    float sj[2] = {0.0f, 0.0f}; // render_vulkan.cpp:2917-2926
    if (a.params.enable_raster_taa > 0) screen_jitter(a.frame_offset, aov->view_frame_id, a.width, a.height, sj);
    aov->v[8] = ref_proj.x / ref_w - cur_proj.x / cur_w;
    aov->v[9] = ref_proj.y / ref_w - cur_proj.y / cur_w;
    aov->v[10] = sj[0];
    aov->v[11] = sj[1];
}

// shade_base_material: rendering/mc/shade_base_material.glsl:14-96 (material unpack, emitter MIS, AOV channels, path-length
// cut, NEE, glossy-only cut, BSDF sampling with its draw order, bounce counting).  Returns SHADING_RESULT_*.
enum { SHADING_RESULT_TERMINATE = -1, SHADING_RESULT_BOUNCE = 1 };
template <class Vis>
static int shade_base_material(const Frame &f, int &bounce, float &prev_bounce_pdf, V3 &illum, V3 &throughput, const rptr_base_material &mp,
                               float approx_sa, V3 w_o, V3 ip, V3 ign, V3 in_, V3 v_x, V3 v_y, PathRng &rng, V3 &w_i, AovOut *aov, V2 uv, V2 duvdx, V2 duvdy,
                               Vis &&visible) {
    const oracle_render_args &a = *f.a;
    const float p_sun = f.sp.sun_radiance[3];
    GltfMat mat;
    V3 emit;
    unpack_material(mat, emit, mp, f.tr, f.s->texset, uv, duvdx, duvdy);
    if (aov && bounce == 0) { // pt_megakernel.glsl:670-672 + shade_base_material.glsl:28-31
        const V3 alb = throughput * mat.base_color;
        const float m[8] = {alb.x, alb.y, alb.z, mat.ior != 1.0f ? mat.roughness : 1.0f, in_.x, in_.y, in_.z, length(ip - f.cam_pos)};
        std::memcpy(aov->v, m, sizeof(m));
        store_motion_jitter_aovs(f, ip, v3(0.0f), aov); // motion_vector = 0: static geometry (pt_megakernel.glsl:426)
    }
    if (a.params.output_channel == 0 && !is_zero(emit)) { // :33-39
        float light_pdf = (1.0f - p_sun) * (1.0f / ((float)f.n_bins * approx_sa));
        float w = nee_mis_heuristic(1.0f, prev_bounce_pdf, 1.0f, light_pdf);
        illum = illum + throughput * w * emit;
    }
    if (a.params.output_channel != 0) { // AOV channels, :42-53; pow(0.25, bounce) is an exact power of two
        float reliability = u2f((uint32_t)(127 - 2 * bounce) << 23);
        if (a.params.output_channel == 1) illum = illum + throughput * mat.base_color * reliability;
        else if (a.params.output_channel == 2) illum = illum + in_ * reliability;
        else if (a.params.output_channel == 3) illum = illum + ip * reliability;
    }
    if (bounce + 1 >= a.params.max_path_depth) return SHADING_RESULT_TERMINATE; // :56-57
    if (a.params.output_channel == 0) { // :59-65, nee.glsl:32-90
        V2 dir_sample, sel_sample;
        dir_sample.x = rng.draw(2); // DIM_POSITION_X
        dir_sample.y = rng.draw(3);
        sel_sample.x = rng.draw(0); // DIM_LIGHT_SEL_1
        sel_sample.y = rng.draw(1);
        NeeSample ns = sample_direct_light(f, mat, ip, ign, in_, w_o, dir_sample, sel_sample, visible);
        illum = illum + throughput * ns.contrib;
    }
    rng.shift_dim(4); // RANDOM_SHIFT_DIM(rng, DIM_LIGHT_END), :66
    if (a.params.glossy_only_mode != 0 && !(mat.roughness < RPTR_GLOSSY_MODE_ROUGHNESS_THRESHOLD && mat.ior != 1.0f)) return SHADING_RESULT_TERMINATE;
    V2 lobe, dirs;
    lobe.x = rng.draw(2); // DIM_LOBE
    lobe.y = rng.draw(3);
    dirs.x = rng.draw(0); // DIM_DIRECTION_X
    dirs.y = rng.draw(1);
    float sampling_pdf = 0.0f, mis_wpdf = 0.0f;
    V3 bsdf = sample_gltf_brdf(mat, in_, w_o, w_i, sampling_pdf, mis_wpdf, dirs, lobe, v_x, v_y, f.tr);
    rng.shift_dim(4); // DIM_VERTEX_END, :82
    ++bounce;
    if (mis_wpdf == 0.0f || is_zero(bsdf) || !(dot(w_i, in_) * dot(w_i, ign) > 0.0f)) return SHADING_RESULT_TERMINATE;
    throughput = throughput * bsdf;
    prev_bounce_pdf = mis_wpdf;
    return SHADING_RESULT_BOUNCE;
}

// ---- rendering/rt/footprint.glsl, with GLSL's matrices spelled out: Mat2 / Mat23 hold columns, (A * B)[c][r] = sum_k A[k][r] * B[c][k];
// every sum of products is the contract's dot product (fma chain) ----
struct Mat2 { V2 c[2]; };
struct Mat23 { V3 c[2]; }; // mat2x3: two columns of three rows
static inline Mat2 mul_t23_23(const Mat23 &a, const Mat23 &b) { // transpose(a) * b
    Mat2 m;
    for (int col = 0; col < 2; ++col) m.c[col] = V2{dot(a.c[0], b.c[col]), dot(a.c[1], b.c[col])};
    return m;
}
static inline Mat2 transpose2(const Mat2 &a) { return Mat2{{V2{a.c[0].x, a.c[1].x}, V2{a.c[0].y, a.c[1].y}}}; }
static inline Mat2 mul22(const Mat2 &a, const Mat2 &b) {
    Mat2 m;
    for (int col = 0; col < 2; ++col) {
        const V2 row0{a.c[0].x, a.c[1].x}, row1{a.c[0].y, a.c[1].y};
        m.c[col] = V2{dot(row0, b.c[col]), dot(row1, b.c[col])};
    }
    return m;
}
static inline Mat2 dpdxy_to_footprint(V3 ray_dir, V3 dpdx, V3 dpdy) { // footprint.glsl:10-15
    V3 t, b;
    ortho_basis(t, b, ray_dir);
    const Mat2 F = mul_t23_23(Mat23{{t, b}}, Mat23{{dpdx, dpdy}});
    return mul22(F, transpose2(F));
}
static inline Mat2 transform_footprint(V3 dst_ray_dir, V3 T0, V3 T1, V3 T2col, V3 src_ray_dir, const Mat2 &F) { // :28-35, T by columns
    V3 t, b;
    ortho_basis(t, b, src_ray_dir);
    const Mat23 T2{{mat_mul(T0, T1, T2col, t), mat_mul(T0, T1, T2col, b)}};
    ortho_basis(t, b, dst_ray_dir);
    const Mat2 T3 = mul_t23_23(Mat23{{t, b}}, T2);
    return mul22(mul22(T3, F), transpose2(T3));
}
static inline Mat2 reflect_footprint(V3 dst_ray_dir, V3 src_ray_dir, const Mat2 &F) { // :38-42
    const V3 n = normalize(dst_ray_dir - src_ray_dir);
    V3 R[3]; // mat3(1.0f) - 2.0f * outerProduct(n, n), column by column
    const float nn[3] = {n.x, n.y, n.z};
    for (int col = 0; col < 3; ++col) {
        R[col] = v3((col == 0 ? 1.0f : 0.0f) - 2.0f * (n.x * nn[col]), (col == 1 ? 1.0f : 0.0f) - 2.0f * (n.y * nn[col]),
                    (col == 2 ? 1.0f : 0.0f) - 2.0f * (n.z * nn[col]));
    }
    return transform_footprint(dst_ray_dir, R[0], R[1], R[2], src_ray_dir, F);
}
static inline void footprint_to_dpdxy(V3 &dpdx, V3 &dpdy, V3 ray_dir, const Mat2 &F) { // :44-61
    const float B = F.c[0].x + F.c[1].y;
    const float C = F.c[0].x * F.c[1].y - F.c[0].y * F.c[1].x;
    const float D = sqrtf(B * B * 0.25f - C);
    const V2 ev{0.5f * B - D, 0.5f * B + D};
    Mat2 X{{V2{1.0f, 0.0f}, V2{0.0f, 1.0f}}};
    if (fabsf(F.c[0].y) > 3.0e-39f) {
        X.c[0] = V2{F.c[1].x, ev.x - F.c[0].x};
        X.c[1] = V2{ev.y - F.c[1].y, F.c[0].y};
    }
    V3 t, b;
    ortho_basis(t, b, ray_dir);
    const V2 x0 = V2{X.c[0].x * (1.0f / sqrtf(dot(X.c[0], X.c[0]))), X.c[0].y * (1.0f / sqrtf(dot(X.c[0], X.c[0])))};
    const V2 x1 = V2{X.c[1].x * (1.0f / sqrtf(dot(X.c[1], X.c[1]))), X.c[1].y * (1.0f / sqrtf(dot(X.c[1], X.c[1])))};
    // mat2x3(t, b) * v = t * v.x + b * v.y, as the contract's matrix-vector product: fma(b, v.y, t * v.x)
    const V3 wx = v3(fmaf(b.x, x0.y, t.x * x0.x), fmaf(b.y, x0.y, t.y * x0.x), fmaf(b.z, x0.y, t.z * x0.x));
    const V3 wy = v3(fmaf(b.x, x1.y, t.x * x1.x), fmaf(b.y, x1.y, t.y * x1.x), fmaf(b.z, x1.y, t.z * x1.x));
    dpdx = wx * sqrtf(ev.x);
    dpdy = wy * sqrtf(ev.y);
}

// bounce prologue, pt_megakernel.glsl:578-580 and :609-678: approximate solid angle of the hit triangle, shading point,
// face-forwarding (unless ONESIDED / VOLUME), one-texel normal map, "fix incident direction" blend, tangent frame
static void bounce_prologue(RTHit &h, const rptr_base_material &mp, const TextureSet &texset, float normal_z_scale, V3 ray_origin, V3 ray_dir,
                            float &approx_sa, V3 &ip, V3 &ign, V3 &in_, V3 &v_x, V3 &v_y, int bounce = 0) {
    approx_sa = length(h.geo_normal); // :578-580
    h.geo_normal = h.geo_normal / approx_sa;
    approx_sa *= fabsf(dot(h.geo_normal, ray_dir)) / (h.dist * h.dist);

    V3 w_o = -ray_dir;
    ip = ray_origin + ray_dir * h.dist; // :613
    ign = h.geo_normal;
    in_ = h.normal;
    if (dot(w_o, ign) < 0.0f) { // :622-633
        if (mp.flags & RPTR_BASE_MATERIAL_VOLUME) {
            ip = ray_origin;
            h.dist = 0.0f;
        } else if (!(mp.flags & RPTR_BASE_MATERIAL_ONESIDED)) {
            in_ = -in_;
            ign = -ign;
        }
    }
    if (mp.normal_map != -1) { // :634-654, the normal map read through the texture set (1 x 1: uv and LOD are irrelevant)
        V3 t_y = normalize(cross(h.normal, h.tangent));
        V3 t_x = cross(t_y, h.normal);
        t_x = t_x * length(h.tangent);
        t_y = t_y * h.bitangent_l;
        TextureSet::RGBA tx = texset.sample_lod((uint32_t)mp.normal_map, h.uv, bounce); // textureLod(.., hit.uv, float(shading_state.bounce)), :641-647
        V3 map_nrm = v3(2.0f * tx.r - 1.0f, 2.0f * tx.g - 1.0f, 1.0f * tx.b - 0.0f);
        map_nrm.z = sqrtf(fmaxf(1.0f - map_nrm.x * map_nrm.x - map_nrm.y * map_nrm.y, 0.0f));
        in_ = normalize(mat_mul(t_x, t_y, in_ * normal_z_scale, map_nrm));
    }
    { // :657-668
        float nw = dot(w_o, in_), gnw = dot(w_o, ign);
        if (nw * gnw <= 0.0f) {
            float blend = gnw / (gnw - nw);
            in_ = normalize(mix(ign, in_, blend - 0.0001f));
        }
    }
    v_y = normalize(cross(in_, h.tangent)); // :677-678
    v_x = cross(v_y, in_);
}

// Russian roulette, pt_megakernel.glsl:715-729 (the caller draws rr_sample once bounce >= rr_path_depth); false = terminate
static bool russian_roulette(int bounce, V3 &throughput, float rr_sample) {
    float prefix = fmaxf(throughput.x, fmaxf(throughput.y, throughput.z));
    float rr_prob = prefix;
    if (bounce > 6) rr_prob = fminf(0.95f, rr_prob);
    else rr_prob = fminf(1.0f, rr_prob);
    if (rr_sample < rr_prob) {
        throughput = throughput / rr_prob;
        return true;
    }
    return false;
}

// ray generation, head of main_spp (pt_megakernel.glsl:311-325): box pixel filter (two draws) unless raster TAA supplies
// the frame's screen jitter
static void camera_ray(const Frame &f, int px, int py, PathRng &rng, uint32_t view_frame_id, V3 &ray_origin, V3 &ray_dir) {
    const oracle_render_args &a = *f.a;
    float ptx = (float)px + 0.5f, pty = (float)py + 0.5f;
    if (a.params.enable_raster_taa == 0) {
        float ux = rng.draw(0); // DIM_PIXEL_X, pathspace.h:13-14
        float uy = rng.draw(1);
        ptx += ux - 0.5f;
        pty += uy - 0.5f;
    }
    ptx /= (float)a.width;
    pty /= (float)a.height;
    if (a.params.enable_raster_taa != 0) { // pt_megakernel.glsl:319-320; screen_jitter belongs to the frame (view_params)
        float sj[2];
        screen_jitter(a.frame_offset, view_frame_id, a.width, a.height, sj);
        ptx += 0.5f * sj[0];
        pty += 0.5f * sj[1];
    }
    ray_origin = f.cam_pos;
    ray_dir = normalize(f.du * ptx + f.dv * pty + f.tl);
}

// the running mean of the resolve pass (process_samples.comp:121-127): history += (x - history) / float(base + batch)
static void fold_sample(float *history, const float *x, uint32_t sample_base_index, uint32_t sample_batch_size) {
    float denom = (float)(sample_base_index + sample_batch_size);
    for (int j = 0; j < 4; ++j) {
        float m = history[j];
        m += (x[j] - m) / denom;
        history[j] = m;
    }
}

static V4 main_spp(const Frame &f, int px, int py, uint32_t sample_index, uint32_t view_frame_id, Counters &cnt, AovOut *aov = nullptr,
                   const rptr_render_ray_query *query = nullptr) {
    const oracle_render_args &a = *f.a;
    const Scene &s = *f.s;
    const rptr_scene_params &sp = f.sp;
    uint32_t linear = (uint32_t)px + (uint32_t)py * (uint32_t)a.width;
    PathRng rng;
    rng.lcg = lcg_seed(sample_index, a.frame_offset, linear);
    rng.alpha = rng.lcg;
    rng.qmc = a.rng_variant != 0;
    if (rng.qmc) rng.q.seed(a.rng_variant, a.pointset_tables, sample_index, view_frame_id, a.frame_offset, (uint32_t)px, (uint32_t)py, (uint32_t)a.width);
    V3 ray_origin, ray_dir;
    camera_ray(f, px, py, rng, view_frame_id, ray_origin, ray_dir);
    float t_min = 0.0f, t_max = 2.e32f;
    if (query) { // pt_megakernel.glsl:327-334: the sampler is seeded and the pixel-filter draws are consumed as for a pixel
        ray_origin = v3(query->origin[0], query->origin[1], query->origin[2]);
        ray_dir = v3(query->dir[0], query->dir[1], query->dir[2]);
        t_max = query->t_max;
    }
    float total_t = 0.0f;
    // USE_MIPMAPPING (librender/render_params.glsl.h:8), pt_megakernel.glsl:338-351: the texture footprint of the pixel
    Mat2 texture_footprint;
    {
        V3 dpdx = f.du / (float)a.width, dpdy = f.dv / (float)a.height;
        dpdx = dpdx * a.params.pixel_radius;
        dpdy = dpdy * a.params.pixel_radius;
        texture_footprint = dpdxy_to_footprint(ray_dir, dpdx, dpdy);
    }
    V3 illum = v3(0.0f), throughput = v3(1.0f);
    int bounce = 0;
    float prev_bounce_pdf = 2.e16f;
    const float p_sun = sp.sun_radiance[3];
    V3 sun_dir = v3(sp.sun_dir[0], sp.sun_dir[1], sp.sun_dir[2]);
    V3 sun_rgb = v3(sp.sun_radiance[0], sp.sun_radiance[1], sp.sun_radiance[2]);

    for (int v = 0; v < a.params.max_path_depth; ++v) {
        rng.set_dim(6 + v * (4 + 4)); // RANDOM_SET_DIM(rng, DIM_CAMERA_END + unrollBounceIdx * (DIM_VERTEX_END + DIM_LIGHT_END)), :423
        // closest hit with front-to-back stochastic alpha (pt_megakernel.glsl:440-478,153-211)
        Hit hit;
        cnt.closest_rays++;
        float after_t = t_min;
        int after_id = 0x7fffffff;
        bool found;
        for (;;) {
            found = closest_hit(s, ray_origin, ray_dir, t_min, t_max, after_t, after_id, hit);
            if (!found) break;
            const Tri &tr = s.tris[hit.tri];
            const GeomInst &g = s.ginst[tr.geom_inst];
            if (g.flags & RPTR_GEOMETRY_FLAGS_NOALPHA) break;
            const rptr_base_material &m = s.materials[calc_hit_material_id(g, (uint32_t)tr.prim)];
            if (m.flags & RPTR_BASE_MATERIAL_NOALPHA) break;
            const V2 cuv = TextureSet::is_handle(m.base_color[0]) ? calc_hit_attributes(g, hit.t, (uint32_t)tr.prim, hit.u, hit.v).uv : V2{0.0f, 0.0f};
            float alpha = material_alpha(s.texset, m, cuv);
            if (!(alpha > 0.0f) || (alpha < 1.0f && rng.draw_alpha() > alpha)) {
                after_t = hit.t;
                after_id = hit.tri;
                continue;
            }
            break;
        }
        if (!found) { // :480-489
            illum = illum + throughput * compute_sky_illum(sp, ray_dir, prev_bounce_pdf);
            if (aov && bounce == 0) { // :482-486: store_geometry_aovs(vec3(0), vec3(2e32), vec3(0)); store_material_aovs(vec3(0), 1, 1)
                const float far_depth = length(v3(2.e32f) - f.cam_pos); // accumulate.glsl:92
                const float m[8] = {0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, far_depth};
                std::memcpy(aov->v, m, sizeof(m));
                store_motion_jitter_aovs(f, v3(2.e32f), v3(0.0f), aov);
            }
            break;
        }
        cnt.vertices++;
        const Tri &tr = s.tris[hit.tri];
        const GeomInst &g = s.ginst[tr.geom_inst];
        RTHit h = calc_hit_attributes(g, hit.t, (uint32_t)tr.prim, hit.u, hit.v);

        total_t += h.dist; // :585 (before a VOLUME hit zeroes hit.dist)
        float geometry_scale = total_t;
        const rptr_base_material &mp = s.materials[h.material_id];
        float approx_sa;
        V3 ip, ign, in_, v_x, v_y;
        const V3 tex_geo_normal = h.geo_normal / length(h.geo_normal); // hit.geo_normal as :579 leaves it (same operations as the prologue)
        bounce_prologue(h, mp, s.texset, sp.normal_z_scale, ray_origin, ray_dir, approx_sa, ip, ign, in_, v_x, v_y, bounce);
        V3 w_o = -ray_dir;
        V2 duvdx{0.0f, 0.0f}, duvdy{0.0f, 0.0f}; // hit.duvdxy, :583-605
        {
            V3 dpdx, dpdy;
            footprint_to_dpdxy(dpdx, dpdy, ray_dir, texture_footprint);
            const V3 dir_tangent_un = ray_dir - tex_geo_normal * dot(ray_dir, tex_geo_normal);
            const float cosTheta2 = fmaxf(1.0f - dot(dir_tangent_un, dir_tangent_un), 0.0f);
            const V3 dir_tangent_elong = dir_tangent_un / (sqrtf(cosTheta2) + cosTheta2);
            const V3 dpdx_ = dpdx + dir_tangent_elong * dot(dpdx, dir_tangent_un);
            const V3 dpdy_ = dpdy + dir_tangent_elong * dot(dpdy, dir_tangent_un);
            const V3 bitangent = cross(tex_geo_normal, normalize(h.tangent)) * h.bitangent_l;
            duvdx = V2{dot(h.tangent, dpdx_) * total_t, dot(bitangent, dpdx_) * total_t};
            duvdy = V2{dot(h.tangent, dpdy_) * total_t, dot(bitangent, dpdy_) * total_t};
        }

        // ---- shade_base_material ----
        V3 w_i;
        {
            const int result = shade_base_material(f, bounce, prev_bounce_pdf, illum, throughput, mp, approx_sa, w_o, ip, ign, in_, v_x, v_y, rng, w_i, aov, h.uv,
                                                   duvdx, duvdy,
                                                   [&](V3 from, V3 dir, float dist) {
                                                       return test_visibility(f, from, dir, dist, geometry_scale, linear, view_frame_id, cnt);
                                                   });
            if (result != SHADING_RESULT_BOUNCE) break;
        }
        if (dot(w_i, in_) * dot(w_o, in_) > -0.999f) texture_footprint = reflect_footprint(w_i, ray_dir, texture_footprint); // :698-702
        // next ray: pt_megakernel.glsl:703-709
        ray_dir = w_i;
        ray_origin = ip;
        t_min = geometry_scale_to_tmin(ray_origin, total_t);
        t_max = 1e20f;
        // Russian roulette: :715-729
        if (bounce >= a.params.rr_path_depth) {
            float rr_sample = rng.draw(-1); // DIM_RR = DIM_FREE_PATH - DIM_VERTEX_END
            if (!russian_roulette(bounce, throughput, rr_sample)) break;
        }
    }
    return V4{illum.x, illum.y, illum.z, bounce == 0 ? 0.0f : 1.0f};
}

static Frame make_frame(const oracle_scene *os, const oracle_render_args *a) {
    Frame f;
    f.s = &os->s;
    f.a = a;
    f.sp = a->scene_params;
    f.n_lights = (int)os->s.lights.size();
    int bs = a->lighting.bin_size;
    f.n_bins = bs > 0 ? (f.n_lights + (bs - 1)) / bs : 0;
    // vulkan/render_sky.cpp:67-70
    if (f.n_lights > 0) f.sp.sun_radiance[3] *= 0.5f;
    else f.sp.sun_radiance[3] = 1.0f;
    float vp[9];
    oracle_view_params(&a->camera, a->width, a->height, vp);
    f.cam_pos = v3(a->camera.pos[0], a->camera.pos[1], a->camera.pos[2]);
    f.du = v3(vp[0], vp[1], vp[2]);
    f.dv = v3(vp[3], vp[4], vp[5]);
    f.tl = v3(vp[6], vp[7], vp[8]);
    oracle_view_projection(&a->camera, a->width, a->height, f.VP);
    f.tr = a->transmission != 0;
    return f;
}

} // namespace

extern "C" {

// Renders n_samples frames of batch_spp = 1 into rgba (W*H*4 floats, row-major, top row first), replaying
// accumulate.glsl:68-73 + process_samples.comp:116-129: frame 0 stores x, frame k folds m += (x - m)/float(k+1).
// If first_sample > 0 the buffer must hold the running mean of the previous frames.
// stats (optional, 3 x uint64): closest rays, shadow rays, path vertices.
static int g_last_threads = 1;
// number of OpenMP threads the last oracle_render() call really ran on (bench.py reports it as cpu_baseline.cores)
int oracle_last_threads(void) { return g_last_threads; }
int oracle_render(const oracle_scene *os, const oracle_render_args *a, float *rgba, uint64_t *stats) {
    Frame f = make_frame(os, a);
    int nt = a->n_threads;
#ifdef _OPENMP
    if (nt <= 0) nt = omp_get_max_threads();
#else
    nt = 1;
#endif
    uint64_t c0 = 0, c1 = 0, c2 = 0;
    // work items are tiles of 16 x 4 pixels over the region (not scanlines: a band of 8 rows would keep 8 threads busy at most)
    const int TW = 16, TH = 4;
    const int tiles_x = (a->x1 - a->x0 + TW - 1) / TW, tiles_y = (a->y1 - a->y0 + TH - 1) / TH;
    int used = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt) reduction(+ : c0, c1, c2)
    for (int tile = 0; tile < tiles_x * tiles_y; ++tile) {
#ifdef _OPENMP
        if (tile == 0) used = omp_get_num_threads();
#endif
        Counters cnt;
        const int ty0 = a->y0 + (tile / tiles_x) * TH, tx0 = a->x0 + (tile % tiles_x) * TW;
        for (int y = ty0; y < ty0 + TH && y < a->y1; ++y)
        for (int x = tx0; x < tx0 + TW && x < a->x1; ++x) {
            float *px = rgba + 4 * ((size_t)y * a->width + x);
            for (int k = 0; k < a->n_samples; ++k) {
                uint32_t frame_id = a->first_sample + (uint32_t)k;
                const uint32_t batch = a->batch_spp > 1 ? (uint32_t)a->batch_spp : 1u;
                V4 c = main_spp(f, x, y, frame_id, a->first_sample + ((uint32_t)k / batch) * batch, cnt);
                float xs[4] = {c.x, c.y, c.z, c.w};
                if (frame_id > 0 && a->params.reprojection_mode != RPTR_REPROJECTION_MODE_DISCARD_HISTORY) // process_samples.comp:116-127
                    fold_sample(px, xs, frame_id, 1u);
                else
                    for (int j = 0; j < 4; ++j) px[j] = xs[j];
            }
        }
        c0 += cnt.closest_rays; c1 += cnt.shadow_rays; c2 += cnt.vertices;
    }
    g_last_threads = used;
    if (stats) { stats[0] = c0; stats[1] = c1; stats[2] = c2; }
    return 0;
}

// render_ray_queries (vulkan/render_vulkan.cpp:1867-1876 -> record_frame :2961-3060): the megakernel dispatched over a "virtual
// screen square" of ceil(sqrt(n)) x ceil(n / that) invocations in 32 x 16 workgroups, batch_spp layers with sample indices
// 0 .. batch_spp - 1 (accumulation_frame_offset = 0).  Invocation index (setup_pixel_assignment.glsl:21-22) = query id;
// gl_GlobalInvocationID.xy is swizzled inside the workgroup (:17-19) and is what seeds the samplers together with the REAL
// frame width.  Results are folded by accumulate_query (vulkan/accumulate.glsl:32-42), layer after layer.
// a->first_sample = view_params.frame_id of the last begin_frame, a->batch_spp = render_params.batch_spp.
// ---- intersection callbacks for the whole-path driver of oracle/_ref (ref_shim/ref_path.cpp): the reference leaves closest hit
//      and occlusion to the Vulkan driver, the composed reference path takes them from here (opaque scenes) ----
struct ref_path_hit { // mirror of the struct in ref_shim/ref_path.cpp
    float t, u, v;
    const uint64_t *qverts3;
    const uint64_t *qnuv3;
    float scale[3], offset[3];
    int32_t has_normals, has_uvs;
    float w2o[9];
    int32_t material_id;
    const uint32_t *id_4pack;
    uint32_t prim;
};
int oracle_cb_closest(void *user, const float *o, const float *d, float tmin, float tmax, ref_path_hit *out) {
    const Scene &s = static_cast<const oracle_scene *>(user)->s;
    Hit h;
    if (!closest_hit(s, v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), tmin, tmax, tmin, 0x7fffffff, h)) return 0;
    const Tri &tr = s.tris[h.tri];
    const GeomInst &g = s.ginst[tr.geom_inst];
    out->t = h.t; out->u = h.u; out->v = h.v;
    out->qverts3 = g.qverts + 3 * (size_t)tr.prim;
    out->qnuv3 = g.qnuv ? g.qnuv + 3 * (size_t)tr.prim : nullptr;
    for (int k = 0; k < 3; ++k) { out->scale[k] = g.scale[k]; out->offset[k] = g.offset[k]; }
    out->has_normals = g.has_normals; out->has_uvs = g.has_uvs;
    for (int r = 0; r < 3; ++r) { out->w2o[3 * r] = g.w2o_row[r].x; out->w2o[3 * r + 1] = g.w2o_row[r].y; out->w2o[3 * r + 2] = g.w2o_row[r].z; }
    out->material_id = g.material_id;
    out->id_4pack = reinterpret_cast<const uint32_t *>(g.tri_mat); // byte k of word i = id of triangle 4 i + k (little endian); must be 4-byte aligned
    out->prim = (uint32_t)tr.prim;
    return 1;
}
int oracle_cb_occluded(void *user, const float *o, const float *d, float tmin, float tmax) {
    const Scene &s = static_cast<const oracle_scene *>(user)->s;
    return any_hit(s, v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), tmin, tmax, [](int, float, float, float) { return true; }) ? 1 : 0;
}

// out = invocations per row / rows of the virtual square, workgroups per row / column (record_frame + dispatch_rays)
void oracle_query_dispatch(int32_t n, int32_t *out) {
    out[0] = (int)std::ceil(std::sqrt((float)n));
    out[1] = (n + out[0] - 1) / out[0];
    out[2] = (out[0] + 31) / 32;
    out[3] = (out[1] + 15) / 16;
}
// the pixel that seeds the samplers of query q (swizzled gl_GlobalInvocationID.xy of invocation index q)
void oracle_query_pixel(uint32_t q, int32_t n, uint32_t *out) {
    int32_t d[4];
    oracle_query_dispatch(n, d);
    const uint32_t groups_x = (uint32_t)d[2];
    const uint32_t group = q / 512u, local = q % 512u; // gl_WorkGroupSize = 32 x 16
    const uint32_t ix = (group % groups_x) * 32u + local % 32u, iy = (group / groups_x) * 16u + local / 32u;
    out[0] = (ix & ~0x18u) + ((iy & 0x3u) << 3);
    out[1] = (iy & ~0x3u) + ((ix & 0x18u) >> 3);
}
// accumulate_query (vulkan/accumulate.glsl:32-42) for one layer
void oracle_accumulate_query(float *r, const float *x, uint32_t sample_index) {
    for (int j = 0; j < 4; ++j) {
        float accum = sample_index > 0 ? r[j] : 0.0f;
        accum += (x[j] - accum) / (float)(sample_index + 1u);
        if (sample_index == 0) r[j] = accum;
        else r[j] += accum;
    }
}
int oracle_render_ray_queries(const oracle_scene *os, const oracle_render_args *a, const rptr_render_ray_query *queries, int32_t n, float *results) {
    Frame f = make_frame(os, a);
    const int batch = a->batch_spp > 1 ? a->batch_spp : 1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int q = 0; q < n; ++q) {
        Counters cnt;
        uint32_t pix[2];
        oracle_query_pixel((uint32_t)q, n, pix);
        for (int k = 0; k < batch; ++k) {
            const V4 c = main_spp(f, (int)pix[0], (int)pix[1], (uint32_t)k, a->first_sample, cnt, nullptr, &queries[q]);
            const float x[4] = {c.x, c.y, c.z, c.w};
            oracle_accumulate_query(results + 4 * (size_t)q, x, (uint32_t)k);
        }
    }
    return 0;
}

// The float values behind the fp16 AOV images for sample `sample_index` (the last layer of a frame is what survives):
// albedo_roughness, normal_depth and motion_jitter (optional) are W*H*4 each.  view_params (frame_id for the raster-TAA
// jitter) are those of the frame that starts at a->first_sample.
int oracle_render_aov3(const oracle_scene *os, const oracle_render_args *a, uint32_t sample_index, float *albedo_roughness, float *normal_depth,
                       float *motion_jitter) {
    Frame f = make_frame(os, a);
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = a->y0; y < a->y1; ++y) {
        Counters cnt;
        for (int x = a->x0; x < a->x1; ++x) {
            AovOut o;
            std::memset(&o, 0, sizeof(o));
            o.view_frame_id = a->first_sample;
            main_spp(f, x, y, sample_index, a->first_sample, cnt, &o);
            std::memcpy(albedo_roughness + 4 * ((size_t)y * a->width + x), o.v, 16);
            std::memcpy(normal_depth + 4 * ((size_t)y * a->width + x), o.v + 4, 16);
            if (motion_jitter) std::memcpy(motion_jitter + 4 * ((size_t)y * a->width + x), o.v + 8, 16);
        }
    }
    return 0;
}
int oracle_render_aov(const oracle_scene *os, const oracle_render_args *a, uint32_t sample_index, float *albedo_roughness, float *normal_depth) {
    oracle_render_args b = *a;
    b.first_sample = sample_index; // a frame of one sample
    return oracle_render_aov3(os, &b, sample_index, albedo_roughness, normal_depth, nullptr);
}

// One un-averaged sample layer (the vec4 main_spp returns) for every pixel of the region: sample_rgba is W*H*4.
int oracle_render_sample(const oracle_scene *os, const oracle_render_args *a, uint32_t sample_index, float *sample_rgba) {
    Frame f = make_frame(os, a);
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = a->y0; y < a->y1; ++y) {
        Counters cnt;
        for (int x = a->x0; x < a->x1; ++x) {
            V4 c = main_spp(f, x, y, sample_index, sample_index, cnt);
            float *px = sample_rgba + 4 * ((size_t)y * a->width + x);
            px[0] = c.x; px[1] = c.y; px[2] = c.z; px[3] = c.w;
        }
    }
    return 0;
}

// pieces of rt_intersect.comp:main: t_min of a query (:41) and the packing of its result (:54-66); geom_inst = instance custom
// index + geometry index, the flattened (instance, geometry) pair
static inline float ray_query_tmin(V3 origin) { return RPTR_RAY_EPSILON * length(origin); }
static inline void pack_ray_result(bool hit, float u, float v, int32_t geom_inst, int32_t prim, float *out) {
    int32_t gi = -1, pr = -1;
    float bu = -1.0f, bv = -1.0f;
    if (hit) { gi = geom_inst; pr = prim; bu = u; bv = v; }
    out[0] = bu;
    out[1] = bv;
    std::memcpy(&out[2], &gi, 4);
    std::memcpy(&out[3], &pr, 4);
}
// one 8-bit texel channel as the texture unit returns it (UNORM8, or sRGB8 through the transfer function)
float oracle_decode_texel(int32_t v, int32_t srgb) { return TextureSet::decode(v, srgb != 0); }
float oracle_ray_query_tmin(const float *origin) { return ray_query_tmin(v3(origin[0], origin[1], origin[2])); }
void oracle_pack_ray_result(int32_t hit, float u, float v, int32_t geom_inst, int32_t prim, float *out) { pack_ray_result(hit != 0, u, v, geom_inst, prim, out); }
// RQ_CLOSEST semantics (vulkan/rt_intersect.comp:28-68): result = (bary.x, bary.y, bits(instance+geometry), bits(prim)),
// miss -> (0,0,bits(-1),bits(-1)); extra_t (optional) receives t.
int oracle_trace_closest(const oracle_scene *os, const rptr_render_ray_query *q, int32_t n, float *results, float *extra_t) {
    const Scene &s = os->s;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        if (q[i].mode_or_data < 0) continue; // rt_intersect.comp:44-45
        Hit h;
        V3 o = v3(q[i].origin[0], q[i].origin[1], q[i].origin[2]), d = v3(q[i].dir[0], q[i].dir[1], q[i].dir[2]);
        const float tmin = ray_query_tmin(o);
        bool ok = closest_hit(s, o, d, tmin, q[i].t_max, tmin, 0x7fffffff, h);
        if (ok) pack_ray_result(true, h.u, h.v, s.tris[h.tri].geom_inst, s.tris[h.tri].prim, &results[4 * i]);
        else pack_ray_result(false, 0.0f, 0.0f, 0, 0, &results[4 * i]);
        if (extra_t) extra_t[i] = ok ? h.t : -1.0f;
    }
    return 0;
}

// Brute-force variant (no BVH): pins the BVH culling itself on small scenes.
int oracle_trace_closest_bruteforce(const oracle_scene *os, const rptr_render_ray_query *q, int32_t n, float *results, float *extra_t) {
    const Scene &s = os->s;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        if (q[i].mode_or_data < 0) continue;
        V3 o = v3(q[i].origin[0], q[i].origin[1], q[i].origin[2]), d = v3(q[i].dir[0], q[i].dir[1], q[i].dir[2]);
        const float tmin = RPTR_RAY_EPSILON * length(o);
        int best = -1;
        float bt = q[i].t_max, bu = 0.0f, bv = 0.0f;
        for (int id = 0; id < (int)s.tris.size(); ++id) {
            float t, u, v;
            if (!intersect_tri(s.tris[id], o, d, t, u, v)) continue;
            if (!(t > tmin && t < q[i].t_max)) continue;
            if (best < 0 || t < bt) { best = id; bt = t; bu = u; bv = v; }
        }
        int32_t gi = -1, prim = -1;
        if (best >= 0) { gi = s.tris[best].geom_inst; prim = s.tris[best].prim; }
        results[4 * i + 0] = best >= 0 ? bu : -1.0f;
        results[4 * i + 1] = best >= 0 ? bv : -1.0f;
        std::memcpy(&results[4 * i + 2], &gi, 4);
        std::memcpy(&results[4 * i + 3], &prim, 4);
        if (extra_t) extra_t[i] = best >= 0 ? bt : -1.0f;
    }
    return 0;
}

// ---- unit entry points used to pin the restatement against oracle/_ref and tests/golden -------------------------
// replay of sampler calls, same protocol as ref_pointset_replay (oracle/ref_shim/ref_pointsets.cpp)
int oracle_pointset_replay(int variant, const uint32_t *const *tables, uint32_t sample_index, uint32_t frame_id, uint32_t frame_offset, uint32_t px,
                           uint32_t py, uint32_t w, const int32_t *ops, const int32_t *args, int n_ops, float *out, uint32_t *state_out) {
    oracle_ps::QmcRng q;
    q.seed(variant, tables, sample_index, frame_id, frame_offset, px, py, w);
    if (state_out) {
        state_out[0] = variant == oracle_ps::BN ? q.pixel : q.index;
        state_out[1] = variant == oracle_ps::BN ? q.sample : q.scramble;
    }
    int n = 0;
    for (int i = 0; i < n_ops; ++i) {
        if (ops[i] == 0) out[n++] = q.draw(args[i]);
        else if (ops[i] == 1) q.set_dim(args[i]);
        else q.shift_dim(args[i]);
    }
    return n;
}
void oracle_screen_jitter(uint32_t frame_offset, uint32_t frame_id, int32_t w, int32_t h, float *out) { screen_jitter(frame_offset, frame_id, w, h, out); }
void oracle_halton_23(int32_t k, float *out) { out[0] = halton_literal(k + 1, 2); out[1] = halton_literal(k + 1, 3); }
uint32_t oracle_morton_sample_id(uint32_t sample_id, uint32_t px, uint32_t py, uint32_t tw, uint32_t th, int hash_tile, int hash_sample) {
    return oracle_ps::morton_sample_id(sample_id, px, py, tw, th, hash_tile != 0, hash_sample != 0);
}
uint32_t oracle_lcg_seed(uint32_t index, uint32_t frame, uint32_t linear) { return lcg_seed(index, frame, linear).state; }
float oracle_lcg_randomf(uint32_t *state) {
    Lcg r{*state};
    float f = lcg_randomf(r);
    *state = r.state;
    return f;
}
void oracle_sincos(float x, float *s, float *c) { sincos_pos(x, *s, *c); }
float oracle_exp(float x) { return exp_f(x); }
float oracle_acos(float x) { return acos_f(x); }
float oracle_fast_positive_atan(float y) { return fast_positive_atan(y); }

static GltfMat mat_from(const rptr_base_material *p, int tr) {
    GltfMat m;
    V3 e;
    unpack_material(m, e, *p, tr != 0, TextureSet(), V2{0.0f, 0.0f});
    return m;
}
void oracle_gltf_bsdf(const rptr_base_material *p, const float *n, const float *wo, const float *wi, int tr, float *out) {
    V3 r = gltf_bsdf(mat_from(p, tr), v3(n[0], n[1], n[2]), v3(wo[0], wo[1], wo[2]), v3(wi[0], wi[1], wi[2]), tr != 0);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float oracle_gltf_wpdf(const rptr_base_material *p, const float *n, const float *wo, const float *wi, int tr) {
    return gltf_wpdf(mat_from(p, tr), v3(n[0], n[1], n[2]), v3(wo[0], wo[1], wo[2]), v3(wi[0], wi[1], wi[2]), tr != 0);
}
// out = weight(3), w_i(3), pdf, mis_wpdf
void oracle_gltf_sample(const rptr_base_material *p, const float *n, const float *wo, const float *vx, const float *vy,
                        const float *rng_sample, const float *fresnel_sample, int tr, float *out) {
    V3 wi;
    float pdf = 0.0f, mis = 0.0f;
    V3 w = sample_gltf_brdf(mat_from(p, tr), v3(n[0], n[1], n[2]), v3(wo[0], wo[1], wo[2]), wi, pdf, mis,
                            V2{rng_sample[0], rng_sample[1]}, V2{fresnel_sample[0], fresnel_sample[1]}, v3(vx[0], vx[1], vx[2]),
                            v3(vy[0], vy[1], vy[2]), tr != 0);
    out[0] = w.x; out[1] = w.y; out[2] = w.z;
    out[3] = wi.x; out[4] = wi.y; out[5] = wi.z;
    out[6] = pdf; out[7] = mis;
}
void oracle_ortho_basis(const float *n, float *vx, float *vy) {
    V3 a, b;
    ortho_basis(a, b, v3(n[0], n[1], n[2]));
    vx[0] = a.x; vx[1] = a.y; vx[2] = a.z;
    vy[0] = b.x; vy[1] = b.y; vy[2] = b.z;
}
// out = solid angle, params(3)
void oracle_triangle_solid_angle(const float *v0, const float *v1, const float *v2, float *out) {
    V3 prm;
    out[0] = triangle_solid_angle(v3(v0[0], v0[1], v0[2]), v3(v1[0], v1[1], v1[2]), v3(v2[0], v2[1], v2[2]), prm);
    out[1] = prm.x; out[2] = prm.y; out[3] = prm.z;
}
void oracle_sample_solid_angle_polygon(const float *v0, const float *v1, const float *v2, const float *rnd, float *out) {
    V3 a = v3(v0[0], v0[1], v0[2]), b = v3(v1[0], v1[1], v1[2]), c = v3(v2[0], v2[1], v2[2]);
    V3 prm;
    float omega = triangle_solid_angle(a, b, c, prm);
    V3 d = sample_solid_angle_polygon(a, b, c, omega, prm, V2{rnd[0], rnd[1]});
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
// sample_tri_lights on an explicit light list; out = L(3), light_dir(3), light_dist, pdf, mis_wpdf
void oracle_sample_tri_lights(const rptr_tri_light_data *lights, int32_t n_lights, int32_t bin_size, const float *hit_p,
                              const float *hit_n, const float *dir_sample, const float *sel_sample, float *out) {
    oracle_scene os;
    os.s.lights.assign(lights, lights + n_lights);
    oracle_render_args a;
    std::memset(&a, 0, sizeof(a));
    a.lighting.bin_size = bin_size;
    Frame f;
    f.s = &os.s;
    f.a = &a;
    f.n_lights = n_lights;
    f.n_bins = (n_lights + bin_size - 1) / bin_size;
    V3 ld;
    float dist, pdf, mis;
    V3 L = sample_tri_lights(f, v3(hit_p[0], hit_p[1], hit_p[2]), v3(hit_n[0], hit_n[1], hit_n[2]), V2{dir_sample[0], dir_sample[1]},
                             V2{sel_sample[0], sel_sample[1]}, ld, dist, pdf, mis);
    out[0] = L.x; out[1] = L.y; out[2] = L.z;
    out[3] = ld.x; out[4] = ld.y; out[5] = ld.z;
    out[6] = dist; out[7] = pdf; out[8] = mis;
}
// shade_base_material for a constants-only material with the LCG pointset; same argument / output layout as ref_shade_base_material
void oracle_shade_base_material(const rptr_base_material *p, int bounce, int output_channel, float prev_bounce_pdf, const float *illum,
                                const float *throughput, float approx_sa, const float *wo, const float *ia, uint32_t rng_state, int max_path_depth,
                                int glossy_only_mode, const float *sun_dir, float sun_cos_angle, const float *sun_radiance,
                                const rptr_tri_light_data *lights, int n_lights, int bin_size, float *out) {
    oracle_scene os;
    if (n_lights > 0) os.s.lights.assign(lights, lights + n_lights);
    oracle_render_args a;
    std::memset(&a, 0, sizeof(a));
    a.lighting.bin_size = bin_size;
    a.params.max_path_depth = max_path_depth;
    a.params.glossy_only_mode = glossy_only_mode;
    a.params.output_channel = output_channel;
    Frame f;
    f.s = &os.s;
    f.a = &a;
    std::memset(&f.sp, 0, sizeof(f.sp));
    for (int k = 0; k < 3; ++k) f.sp.sun_dir[k] = sun_dir[k];
    f.sp.sun_cos_angle = sun_cos_angle;
    for (int k = 0; k < 4; ++k) f.sp.sun_radiance[k] = sun_radiance[k];
    f.n_lights = n_lights;
    f.n_bins = bin_size > 0 ? (n_lights + bin_size - 1) / bin_size : 0;
    f.tr = false;
    f.cam_pos = v3(0.0f);
    PathRng rng;
    rng.lcg.state = rng_state;
    rng.alpha = rng.lcg;
    rng.qmc = false;
    V3 il = v3(illum[0], illum[1], illum[2]), thr = v3(throughput[0], throughput[1], throughput[2]), w_i = v3(0.0f);
    int queries = 0;
    V3 qd = v3(0.0f);
    float qdist = 0.0f;
    const float pdf_before = prev_bounce_pdf;
    const int result = shade_base_material(f, bounce, prev_bounce_pdf, il, thr, *p, approx_sa, v3(wo[0], wo[1], wo[2]), v3(ia[0], ia[1], ia[2]),
                                           v3(ia[3], ia[4], ia[5]), v3(ia[6], ia[7], ia[8]), v3(ia[9], ia[10], ia[11]), v3(ia[12], ia[13], ia[14]), rng,
                                           w_i, nullptr, V2{0.0f, 0.0f}, V2{0.0f, 0.0f}, V2{0.0f, 0.0f}, [&](V3, V3 dir, float dist) {
                                               ++queries; qd = dir; qdist = dist;
                                               return true;
                                           });
    std::memset(out, 0, 19 * sizeof(float));
    out[0] = (float)result; out[1] = (float)bounce; out[2] = prev_bounce_pdf;
    out[3] = il.x; out[4] = il.y; out[5] = il.z; out[6] = thr.x; out[7] = thr.y; out[8] = thr.z;
    out[9] = w_i.x; out[10] = w_i.y; out[11] = w_i.z;
    out[12] = result == SHADING_RESULT_BOUNCE ? prev_bounce_pdf : 0.0f; // aux.mis_pdf; (void)pdf_before
    (void)pdf_before;
    std::memcpy(&out[13], &rng.lcg.state, 4);
    out[14] = (float)queries;
    out[15] = qd.x; out[16] = qd.y; out[17] = qd.z; out[18] = qdist;
}
// sample_direct_light for a constants-only material (no transmission); same argument / output layout as ref_sample_direct_light
void oracle_sample_direct_light(const rptr_base_material *p, const float *hp, const float *gn, const float *n, const float *vx, const float *vy,
                                const float *wo, const float *u4, const float *sun_dir, float sun_cos_angle, const float *sun_radiance,
                                const rptr_tri_light_data *lights, int n_lights, int bin_size, float *out) {
    (void)vx; (void)vy;
    oracle_scene os;
    if (n_lights > 0) os.s.lights.assign(lights, lights + n_lights);
    oracle_render_args a;
    std::memset(&a, 0, sizeof(a));
    a.lighting.bin_size = bin_size;
    Frame f;
    f.s = &os.s;
    f.a = &a;
    std::memset(&f.sp, 0, sizeof(f.sp));
    for (int k = 0; k < 3; ++k) f.sp.sun_dir[k] = sun_dir[k];
    f.sp.sun_cos_angle = sun_cos_angle;
    for (int k = 0; k < 4; ++k) f.sp.sun_radiance[k] = sun_radiance[k];
    f.n_lights = n_lights;
    f.n_bins = bin_size > 0 ? (n_lights + bin_size - 1) / bin_size : 0;
    f.tr = false;
    GltfMat m;
    V3 e;
    unpack_material(m, e, *p, false, TextureSet(), V2{0.0f, 0.0f});
    int queries = 0;
    V3 qf = v3(0.0f), qd = v3(0.0f);
    float qdist = 0.0f;
    NeeSample r = sample_direct_light(f, m, v3(hp[0], hp[1], hp[2]), v3(gn[0], gn[1], gn[2]), v3(n[0], n[1], n[2]), v3(wo[0], wo[1], wo[2]),
                                      V2{u4[0], u4[1]}, V2{u4[2], u4[3]}, [&](V3 from, V3 dir, float dist) {
                                          ++queries; qf = from; qd = dir; qdist = dist;
                                          return true;
                                      });
    std::memset(out, 0, 16 * sizeof(float));
    out[0] = r.contrib.x; out[1] = r.contrib.y; out[2] = r.contrib.z;
    if (r.mis_pdf != 0.0f) { // the reference fills aux_info only when it returns a contribution
        out[3] = r.light_dir.x; out[4] = r.light_dir.y; out[5] = r.light_dir.z; out[6] = r.light_dist;
    }
    out[7] = r.mis_pdf; out[8] = (float)queries;
    out[9] = qf.x; out[10] = qf.y; out[11] = qf.z; out[12] = qd.x; out[13] = qd.y; out[14] = qd.z; out[15] = qdist;
}
// the texture unit (shading_oracle.h TextureSet::sample) on one texture: out = rgba
void oracle_sample_texture(const rptr_texture_desc *t, float u, float v, float *out) {
    TextureSet ts;
    ts.tex = t;
    ts.n = 1;
    const TextureSet::RGBA c = ts.sample(0u, V2{u, v});
    out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = c.a;
}
// textureGrad / textureLod on one texture description (block-compressed input is decoded first, like oracle_scene_create does)
static bool decoded_copy(const rptr_texture_desc *t, rptr_texture_desc &td, std::vector<uint8_t> &store) {
    td = *t;
    if (!texture_levels_to_rgba8(td, store)) return false;
    if (td.bc_format != 0) { td.bc_format = 0; td.channels = 4; }
    td.texels = store.data();
    return true;
}
void oracle_sample_texture_grad(const rptr_texture_desc *t, float u, float v, const float *ddx, const float *ddy, float *out) {
    rptr_texture_desc td;
    std::vector<uint8_t> store;
    if (!decoded_copy(t, td, store)) return;
    TextureSet ts;
    ts.tex = &td;
    ts.n = 1;
    const TextureSet::RGBA c = ts.sample_grad(0u, V2{u, v}, V2{ddx[0], ddx[1]}, V2{ddy[0], ddy[1]});
    out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = c.a;
}
void oracle_sample_texture_lod(const rptr_texture_desc *t, float u, float v, int32_t level, float *out) {
    rptr_texture_desc td;
    std::vector<uint8_t> store;
    if (!decoded_copy(t, td, store)) return;
    TextureSet ts;
    ts.tex = &td;
    ts.n = 1;
    const TextureSet::RGBA c = ts.sample_lod(0u, V2{u, v}, level);
    out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = c.a;
}
float oracle_log2(float x) { return TextureSet::log2_positive(x); }
// all levels of a texture description as RGBA8 (missing channels: colour 0, alpha 255); returns the number of bytes
int64_t oracle_decode_texture(const rptr_texture_desc *t, uint8_t *out, int64_t capacity) {
    std::vector<uint8_t> store;
    if (!texture_levels_to_rgba8(*t, store)) return -1;
    if (t->bc_format != 0) {
        if (out && (int64_t)store.size() <= capacity) std::memcpy(out, store.data(), store.size());
        return (int64_t)store.size();
    }
    const int64_t texels = (int64_t)store.size() / t->channels;
    if (out && 4 * texels <= capacity)
        for (int64_t i = 0; i < texels; ++i)
            for (int k = 0; k < 4; ++k) out[4 * i + k] = k < t->channels ? store[(size_t)i * t->channels + k] : (k == 3 ? 255 : 0);
    return 4 * texels;
}
// rendering/rt/footprint.glsl: same operations and layouts as hostsim_footprint_op
void oracle_footprint_op(int32_t op, const float *in, float *out) {
    if (op == 0) {
        const Mat2 F = dpdxy_to_footprint(v3(in[0], in[1], in[2]), v3(in[3], in[4], in[5]), v3(in[6], in[7], in[8]));
        out[0] = F.c[0].x; out[1] = F.c[0].y; out[2] = F.c[1].x; out[3] = F.c[1].y;
    } else if (op == 1) {
        const Mat2 F = reflect_footprint(v3(in[0], in[1], in[2]), v3(in[3], in[4], in[5]), Mat2{{V2{in[6], in[7]}, V2{in[8], in[9]}}});
        out[0] = F.c[0].x; out[1] = F.c[0].y; out[2] = F.c[1].x; out[3] = F.c[1].y;
    } else {
        V3 dx, dy;
        footprint_to_dpdxy(dx, dy, v3(in[0], in[1], in[2]), Mat2{{V2{in[3], in[4]}, V2{in[5], in[6]}}});
        out[0] = dx.x; out[1] = dx.y; out[2] = dx.z; out[3] = dy.x; out[4] = dy.y; out[5] = dy.z;
    }
}
// unpack_material + get_material_alpha with 8-bit 1 x 1 textures; same output layout as ref_unpack_material
void oracle_unpack_material(const rptr_base_material *p, const rptr_texture_desc *textures, int n_textures, int transmission, float *out) {
    TextureSet ts;
    ts.tex = textures;
    ts.n = n_textures;
    GltfMat m;
    V3 e;
    std::memset(out, 0, 17 * sizeof(float));
    out[15] = unpack_material(m, e, *p, transmission != 0, ts, V2{0.0f, 0.0f});
    out[16] = material_alpha(ts, *p, V2{0.0f, 0.0f});
    out[0] = m.base_color.x; out[1] = m.base_color.y; out[2] = m.base_color.z;
    out[3] = m.metallic; out[4] = m.specular; out[5] = m.roughness; out[6] = m.ior;
    if (transmission) {
        out[7] = m.specular_transmission; out[8] = m.transmission_roughness;
        out[9] = m.transmission_color.x; out[10] = m.transmission_color.y; out[11] = m.transmission_color.z;
    }
    out[12] = e.x; out[13] = e.y; out[14] = e.z;
}
// unpack_material + get_material_alpha at a hit with image textures: uv, duvdxy = (d(uv)/dx, d(uv)/dy); transmission build; layout as above
void oracle_unpack_material_at(const rptr_base_material *p, const rptr_texture_desc *textures, int n_textures, const float *uv, const float *duvdxy,
                               float *out) {
    std::vector<rptr_texture_desc> dec((size_t)n_textures);
    std::vector<std::vector<uint8_t>> store((size_t)n_textures);
    for (int i = 0; i < n_textures; ++i)
        if (!decoded_copy(textures + i, dec[(size_t)i], store[(size_t)i])) return;
    TextureSet ts;
    ts.tex = dec.data();
    ts.n = n_textures;
    GltfMat m;
    V3 e;
    std::memset(out, 0, 17 * sizeof(float));
    const V2 at{uv[0], uv[1]}, dx{duvdxy[0], duvdxy[1]}, dy{duvdxy[2], duvdxy[3]};
    out[15] = unpack_material(m, e, *p, true, ts, at, dx, dy);
    out[16] = ts.color_param(p->base_color, at, dx, dy).a;
    out[0] = m.base_color.x; out[1] = m.base_color.y; out[2] = m.base_color.z;
    out[3] = m.metallic; out[4] = m.specular; out[5] = m.roughness; out[6] = m.ior;
    out[7] = m.specular_transmission; out[8] = m.transmission_roughness;
    out[9] = m.transmission_color.x; out[10] = m.transmission_color.y; out[11] = m.transmission_color.z;
    out[12] = e.x; out[13] = e.y; out[14] = e.z;
}
// head of main_spp for one pixel sample with the LCG pointset: out = origin(3), dir(3), bits(LCG state afterwards)
void oracle_camera_ray(const oracle_scene *os, const oracle_render_args *a, int32_t px, int32_t py, uint32_t sample_index, float *out) {
    Frame f = make_frame(os, a);
    PathRng rng;
    rng.lcg = lcg_seed(sample_index, a->frame_offset, (uint32_t)px + (uint32_t)py * (uint32_t)a->width);
    rng.alpha = rng.lcg;
    rng.qmc = false;
    V3 o, d;
    camera_ray(f, px, py, rng, a->first_sample, o, d);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = d.x; out[4] = d.y; out[5] = d.z;
    std::memcpy(out + 6, &rng.lcg, 4);
}
// out = (tmin, tmax); returns 1 when the shadow ray is cast
int32_t oracle_shadow_ray_range(const float *from, float dist, float geometry_scale, float *out) {
    return shadow_ray_range(v3(from[0], from[1], from[2]), dist, geometry_scale, out[0], out[1]) ? 1 : 0;
}
uint32_t oracle_shadow_alpha_seed(uint32_t prim, uint32_t instance, uint32_t frame_id, uint32_t frame_offset, uint32_t pixel_linear) {
    return shadow_alpha_rng(prim, instance, frame_id, frame_offset, pixel_linear).state;
}
// 1 = rejected; *lcg_state advances when a draw was needed.  material_flags: BASE_MATERIAL_NOALPHA skips the test (:204)
int32_t oracle_alpha_filter(float alpha, uint32_t material_flags, uint32_t *lcg_state) {
    if (material_flags & RPTR_BASE_MATERIAL_NOALPHA) return 0;
    Lcg r{*lcg_state};
    const bool rejected = alpha_test_rejects(alpha, r);
    *lcg_state = r.state;
    return rejected ? 1 : 0;
}
void oracle_running_mean(const float *x, float *history, uint32_t sample_base_index, uint32_t sample_batch_size) {
    fold_sample(history, x, sample_base_index, sample_batch_size);
}
// geometry_scale_to_tmin (vulkan/geometry.glsl:76-78)
float oracle_geometry_scale_to_tmin(const float *orig, float geometry_scale) { return geometry_scale_to_tmin(v3(orig[0], orig[1], orig[2]), geometry_scale); }
// bounce prologue of main_spp; same in / out layout as ref_bounce_prologue (oracle/ref_shim/ref_loop.cpp), the normal-map
// texel given as the three 8-bit values of a linear 1 x 1 texture
void oracle_bounce_prologue(const float *in, uint32_t material_flags, int32_t has_normal_map, const uint8_t *texel8, float normal_z_scale, float *out) {
    RTHit h;
    std::memset(&h, 0, sizeof(h));
    h.normal = v3(in[0], in[1], in[2]);
    h.dist = in[3];
    h.geo_normal = v3(in[4], in[5], in[6]);
    h.tangent = v3(in[7], in[8], in[9]);
    h.bitangent_l = in[10];
    rptr_base_material mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.flags = material_flags;
    mp.normal_map = has_normal_map ? 0 : -1;
    rptr_texture_desc td{1, 1, 3, RPTR_COLOR_SPACE_LINEAR, texel8};
    TextureSet ts;
    ts.tex = &td;
    ts.n = 1;
    float approx_sa;
    V3 ip, ign, in_, v_x, v_y;
    bounce_prologue(h, mp, ts, normal_z_scale, v3(in[11], in[12], in[13]), v3(in[14], in[15], in[16]), approx_sa, ip, ign, in_, v_x, v_y);
    const float o[17] = {approx_sa, ip.x, ip.y, ip.z, ign.x, ign.y, ign.z, in_.x, in_.y, in_.z, v_x.x, v_x.y, v_x.z, v_y.x, v_y.y, v_y.z, h.dist};
    std::memcpy(out, o, sizeof(o));
}
// Russian roulette of main_spp: 1 = survives (throughput divided by the survival probability), 0 = terminated
int32_t oracle_russian_roulette(int32_t bounce, int32_t rr_path_depth, float *throughput, float rr_sample) {
    if (!(bounce >= rr_path_depth)) return 1;
    V3 t = v3(throughput[0], throughput[1], throughput[2]);
    const bool alive = russian_roulette(bounce, t, rr_sample);
    throughput[0] = t.x; throughput[1] = t.y; throughput[2] = t.z;
    return alive ? 1 : 0;
}
// compute_sky_illum (vulkan/pt_megakernel.glsl:113-149); sp->sun_radiance[3] = p_sun as the shader sees it
void oracle_compute_sky_illum(const rptr_scene_params *sp, const float *dir, float prev_pdf, float *out) {
    V3 r = compute_sky_illum(*sp, v3(dir[0], dir[1], dir[2]), prev_pdf);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void oracle_skymodel_radiance(const rptr_scene_params *sp, const float *sun_dir, const float *view, float *out) {
    V3 r = skymodel_radiance(*sp, v3(sun_dir[0], sun_dir[1], sun_dir[2]), v3(view[0], view[1], view[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// the sun branch of sample_direct_light as main_spp writes it (mc/lights_sun.glsl:8-17, lights/sun.glsl:9-20): direction + pdf
void oracle_sample_sun_dir(const float *sun, float cos_radius, const float *u2, float *out) {
    V3 sun_dir = v3(sun[0], sun[1], sun[2]);
    float sn, cs;
    sincos_pos(TWO_PI_F * u2[0], sn, cs);
    float cosT = mix(1.0f, cos_radius, u2[1]);
    float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
    V3 fx, fy;
    ortho_basis(fx, fy, sun_dir);
    V3 d = mat_mul(fx, fy, sun_dir, v3(sinT * cs, sinT * sn, cosT));
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
    out[3] = 1.0f / (TWO_PI_F * (1.0f - cos_radius));
}
void oracle_sky_illum(const rptr_scene_params *sp, const float *dir, float prev_pdf, float *out) {
    V3 r = compute_sky_illum(*sp, v3(dir[0], dir[1], dir[2]), prev_pdf);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// calc_hit_attributes for triangle `prim` of flattened geometry-instance gi; out = normal(3), dist, geo_normal(3),
// material_id (as float), tangent(3), bitangent_l, uv(2)
void oracle_hit_attributes(const oracle_scene *os, int32_t gi, int32_t prim, float t, float u, float v, float *out) {
    RTHit h = calc_hit_attributes(os->s.ginst[gi], t, (uint32_t)prim, u, v);
    out[0] = h.normal.x; out[1] = h.normal.y; out[2] = h.normal.z; out[3] = h.dist;
    out[4] = h.geo_normal.x; out[5] = h.geo_normal.y; out[6] = h.geo_normal.z; out[7] = (float)h.material_id;
    out[8] = h.tangent.x; out[9] = h.tangent.y; out[10] = h.tangent.z; out[11] = h.bitangent_l;
    out[12] = h.uv.x; out[13] = h.uv.y;
}
void oracle_dequantize_position(uint64_t q, const float *scale, const float *offset, float *out) {
    V3 p = dequantize_position(q, scale, offset);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void oracle_dequantize_normal(uint32_t w, float *out) {
    V3 n = dequantize_normal(w);
    out[0] = n.x; out[1] = n.y; out[2] = n.z;
}
void oracle_dequantize_uv(uint32_t w, float *out) {
    V2 uv = dequantize_uv(w);
    out[0] = uv.x; out[1] = uv.y;
}
// librender/lights.cpp pre-pass on an explicit emitter list; returns the binned count (out sized >= 2*n + 2*bin)
int32_t oracle_bin_emitters(const rptr_tri_light_data *in, int32_t n, const rptr_light_sampling_config *ls, rptr_tri_light_data *out, int32_t max_out) {
    std::vector<Emitter> em(n);
    for (int i = 0; i < n; ++i)
        em[i] = Emitter{v3(in[i].v0[0], in[i].v0[1], in[i].v0[2]), v3(in[i].v1[0], in[i].v1[1], in[i].v1[2]),
                        v3(in[i].v2[0], in[i].v2[1], in[i].v2[2]), v3(in[i].radiance[0], in[i].radiance[1], in[i].radiance[2])};
    std::vector<float> rad = estimate_normalized_radiance(em, ls->min_perceived_receiver_dist);
    equalize_emitter_bins(em, rad, ls->bin_size);
    if ((int)em.size() > max_out) return -(int)em.size();
    for (size_t i = 0; i < em.size(); ++i) {
        out[i].v0[0] = em[i].v0.x; out[i].v0[1] = em[i].v0.y; out[i].v0[2] = em[i].v0.z;
        out[i].v1[0] = em[i].v1.x; out[i].v1[1] = em[i].v1.y; out[i].v1[2] = em[i].v1.z;
        out[i].v2[0] = em[i].v2.x; out[i].v2[1] = em[i].v2.y; out[i].v2[2] = em[i].v2.z;
        out[i].radiance[0] = em[i].radiance.x; out[i].radiance[1] = em[i].radiance.y; out[i].radiance[2] = em[i].radiance.z;
    }
    return (int32_t)em.size();
}
int32_t oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"

// ---- the temporal passes of the ENABLE_REALTIME_RESOLVE build (post_oracle.h), whole images -------------------------------------
#include "post_oracle.h"
extern "C" {
// process_samples.comp:106-113 for every pixel: accum_cur = this frame's samples (RGBA32F); stored = accumulator after the pass,
// shown = the colour handed on to the display chain
void oracle_reproject_accumulate(int32_t w, int32_t h, const float *accum_cur, const float *history, const uint16_t *nd_history, const uint16_t *nd,
                                 const uint16_t *mj, float min_sample_weight, int32_t batch, float *stored, float *shown) {
    const post::ReprojectIn in{post::Image4f{history, w, h}, post::Image4h{nd_history, w, h}, post::Image4h{nd, w, h}, post::Image4h{mj, w, h}};
#pragma omp parallel for schedule(dynamic, 4)
    for (int32_t y = 0; y < h; ++y)
        for (int32_t x = 0; x < w; ++x) {
            const size_t i = 4 * ((size_t)y * w + x);
            V4 st;
            const V4 sh = post::reproject_and_accumulate(in, V4{accum_cur[i], accum_cur[i + 1], accum_cur[i + 2], accum_cur[i + 3]}, post::IV2{x, y},
                                                         post::IV2{w, h}, min_sample_weight, batch, &st);
            stored[i] = st.x; stored[i + 1] = st.y; stored[i + 2] = st.z; stored[i + 3] = st.w;
            shown[i] = sh.x; shown[i + 1] = sh.y; shown[i + 2] = sh.z; shown[i + 3] = sh.w;
        }
}
// process_taa.comp main() for every pixel of the (w x h) LDR target; the motion image has the render size (rw x rh)
void oracle_process_taa(int32_t w, int32_t h, int32_t upscale, int32_t rw, int32_t rh, const uint8_t *current, const uint8_t *history,
                        const uint16_t *mj, uint8_t *out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int32_t y = 0; y < h; ++y)
        for (int32_t x = 0; x < w; ++x)
            post::process_taa_main(post::Image4b{current, w, h}, post::Image4b{history, w, h}, post::Image4h{mj, rw, rh}, post::IV2{x, y}, post::IV2{w, h},
                                   upscale, out + 4 * ((size_t)y * w + x));
}
} // extern "C"
