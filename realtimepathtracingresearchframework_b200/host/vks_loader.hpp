// vks_loader.hpp -- `.vks` scene files (+ `.vkt` textures and parameter files of the scene's texture directory) for the headless driver.
//
// The C++ twin of realtimepathtracingresearchframework_b200/vks.py: the container after ext/libvkr/src/vkr.c (vkr_load_scene :771-1144,
// vkr_open_texture :211-306, vkr_load_material :521-625, vkr_dequantize_transform :1382-1408), the mapping to the backend's scene
// description after Scene::load_vkrs (librender/scene.cpp:544-1006).  File versions 3 and 4; frame 0 of the transform table.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rptr_types.h"

namespace rptr_host {

struct VksTexture {
    int32_t width = 1, height = 1, channels = 4, color_space = RPTR_COLOR_SPACE_LINEAR, bc_format = 0, mip_levels = 1;
    std::vector<uint8_t> bytes; // all levels back to back, as stored in the .vkt file (or the 1 x 1 default texel)
};

struct VksScene {
    std::vector<uint8_t> file;                      // the .vks file (per-triangle material ids point into it)
    std::vector<std::vector<uint64_t>> streams;     // per geometry: vertex stream, normal + uv stream (aligned copies)
    std::vector<rptr_geometry_desc> geometries;
    std::vector<rptr_mesh_desc> meshes;
    std::vector<std::vector<int32_t>> material_offsets;
    std::vector<rptr_pmesh_desc> pmeshes;
    std::vector<rptr_instance_desc> instances;
    std::vector<rptr_base_material> materials;
    std::vector<VksTexture> textures;
    std::vector<rptr_texture_desc> texture_descs;
    std::vector<std::string> material_names;
    int64_t total_triangles = 0;
    rptr_scene_desc desc() const;                   // valid while this object lives and is not modified
    uint64_t hash() const;                          // FNV-1a over the tables (tests: equals tests/test_cli.py's hash of vks.load_vks)
};

// throws std::runtime_error with a readable message
void load_vks(const std::string &path, VksScene &out, bool ignore_textures = false, bool load_specularity = false);

} // namespace rptr_host
