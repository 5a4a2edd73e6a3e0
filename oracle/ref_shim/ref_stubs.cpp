// oracle/ref_shim/ref_stubs.cpp -- TEST INFRASTRUCTURE.
// librender/lights.cpp also defines collect_emitters(Scene const&), which references scene/mesh members defined in
// reference translation units we do not build (scene.cpp needs libvkr, file mapping, ...).  The oracle never calls
// that overload (emitters are collected from the flat scene description); these aborting stubs only satisfy the linker.
#include <cstdio>
#include <cstdlib>
#include "scene.h"

static void unreachable(const char *what) {
    std::fprintf(stderr, "oracle/_ref: %s is not available in the reduced reference build\n", what);
    std::abort();
}
glm::mat4 AnimationData::dequantize(uint32_t, uint32_t) const { unreachable("AnimationData::dequantize"); return glm::mat4(1.0f); }
int Geometry::num_tris() const { unreachable("Geometry::num_tris"); return 0; }
len_t Mesh::num_tris() const { unreachable("Mesh::num_tris"); return 0; }
int Mesh::num_geometries() const { unreachable("Mesh::num_geometries"); return 0; }
int ParameterizedMesh::material_offset(int) const { unreachable("ParameterizedMesh::material_offset"); return 0; }
int ParameterizedMesh::triangle_material_id(index_t) const { unreachable("ParameterizedMesh::triangle_material_id"); return 0; }
bool ParameterizedMesh::per_triangle_materials() const { unreachable("ParameterizedMesh::per_triangle_materials"); return false; }

// more link-only stubs for symbols pulled in by inline code in the reference headers
FileMapping::~FileMapping() {}
const uint8_t *FileMapping::data() const { unreachable("FileMapping::data"); return nullptr; }
size_t FileMapping::nbytes() const { unreachable("FileMapping::nbytes"); return 0; }
void throw_ilen_overflow(int, intmax_t) { unreachable("throw_ilen_overflow"); }
void throw_int_overflow(intmax_t, intmax_t) { unreachable("throw_int_overflow"); }
void throw_uint_overflow(unsigned, intmax_t) { unreachable("throw_uint_overflow"); }
// glue: util/util.cpp:293-296 (that file needs OS/process helpers we do not build)
float luminance(const glm::vec3 &c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }
bool in_stack_unwind() { return false; }
void FileMapping::release_resources() {}
