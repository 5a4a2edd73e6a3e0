// rptr_host.hpp -- host-side scene ingestion for the CUDA backend: what RenderVulkan::set_scene and the binned-lights
// extension do on the CPU before anything reaches the GPU (vulkan/render_vulkan.cpp:1554-1644,2748-2850;
// vulkan/light_sampling/render_binned_lights.cpp:68-149; librender/lights.cpp:14-349).
#pragma once
#include <string>
#include <vector>

#include "rptr_bvh.cuh"

namespace rp {

struct HostGeomInst {
    GeomInst g;       // pointers are indices-to-be-patched: see stream ids below
    int32_t geometry; // index into the scene's geometry streams
    int32_t pmesh;    // for tri_mat
    int64_t prim_offset;
    float o2w[12];
};

struct HostTexture { // one entry per texture of the scene; rgba is empty for 1 x 1 textures (folded into the materials)
    int32_t width = 0, height = 0, srgb = 0;
    int32_t levels = 1;        // mip levels in rgba, base level first, level l = max(width >> l, 1) x max(height >> l, 1)
    std::vector<uint8_t> rgba; // four channels per texel: missing colour channels 0, missing alpha 255; block-compressed input decoded
};
// decodes every level of a texture description (RGBA8 expansion, BC1 / BC3 / BC5 blocks) -- also used by the tests
void decode_texture(const rptr_texture_desc &td, HostTexture &out);

struct HostScene {
    // owned copies of the input streams (the caller's Scene is only borrowed for set_scene, app.cpp:151-175)
    std::vector<std::vector<uint64_t>> qverts, qnuv;
    std::vector<std::vector<uint8_t>> tri_mat;
    std::vector<HostGeomInst> ginst;
    std::vector<rptr_base_material> materials;   // texture handles already resolved to constants (1x1-texel mode)
    std::vector<int32_t> material_alpha8;       // per material: alpha texel (255 = opaque)
    std::vector<char> material_alpha_textured;  // per material: alpha comes from a texture larger than 1 x 1 (looked up per candidate)
    std::vector<HostTexture> textures;          // device image of Scene::textures
    bool any_textured = false;                  // some parameter kept its handle (texture larger than 1 x 1)
    float srgb_lut[256];                        // sRGB8 code value -> linear
    std::vector<float> normal_texels;           // per material: rgb (+ pad) of its 1 x 1 normal map, zeros without one
    bool any_normal_map = false;
    bool any_alpha_tested = false;              // some triangle has alpha8 != 255: traversal must run the candidate filter
    std::vector<rptr_tri_light_data> lights;
    std::vector<Tri> tris;      // flattened (instance, geometry, primitive) order; Tri::id == index
    std::vector<BvhNode> nodes; // node 0 = root
    std::vector<Tri> leaf_tris; // leaf order
    bool any_non_opaque = false;
    double bvh_build_ms = 0.0;
    float sah_cost = 0.0f;
};

// Throws std::runtime_error with a readable message on invalid input.
// with_bvh = false leaves nodes / leaf_tris empty (the device builder of rptr_bvh_build.cu takes over from `tris`)
void build_host_scene(const rptr_scene_desc &d, const rptr_light_sampling_config &ls, HostScene &out, bool with_bvh = true);
void build_bvh(HostScene &s);

// update_view_parameters (vulkan/render_vulkan.cpp:2880-2894): out = du, dv, top_left
void view_params(const rptr_camera_params &cam, int w, int h, float *du, float *dv, float *tl);
// view_params.VP (vulkan/render_vulkan.cpp:2926-2930), column-major 4 x 4
void view_projection(const rptr_camera_params &cam, int w, int h, float *vp);

// raster-TAA screen jitter of the frame (vulkan/render_vulkan.cpp:2917-2926); halton_23(k) = entry k of librender/halton.h
void screen_jitter(uint32_t frame_offset, uint32_t frame_id, int w, int h, float *out);
void halton_23(int k, float *out);

} // namespace rp
