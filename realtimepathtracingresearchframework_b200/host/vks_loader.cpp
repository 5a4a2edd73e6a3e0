// vks_loader.cpp -- see vks_loader.hpp.  Written from the file layout vkr.c reads and the mapping scene.cpp applies; no reference code.
#include "vks_loader.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace rptr_host {
namespace {

const int32_t VKR_MAGIC = 0xABCABC, VKT_MAGIC = 0xBC1BC1;
const size_t TRANSFORM_BYTES = 24; // 3 x f32 translation, f32 scaling, 4 x u16 quaternion

std::vector<uint8_t> read_file(const std::string &path, bool *missing = nullptr) {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        if (missing) { *missing = true; return {}; }
        throw std::runtime_error("cannot open " + path);
    }
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    f.seekg(0);
    std::vector<uint8_t> d((size_t)n);
    if (n > 0) f.read(reinterpret_cast<char *>(d.data()), n);
    return d;
}

struct Cursor {
    const std::vector<uint8_t> &d;
    const std::string &name;
    size_t pos = 0;
    template <class T> T take() {
        if (pos + sizeof(T) > d.size()) throw std::runtime_error(name + ": truncated file");
        T v;
        memcpy(&v, d.data() + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    std::string string() { // u64 length, the characters, a terminating zero (vkr_load_string)
        const uint64_t n = take<uint64_t>();
        if (pos + n + 1 > d.size()) throw std::runtime_error(name + ": truncated string");
        std::string s(reinterpret_cast<const char *>(d.data() + pos), (size_t)n);
        pos += (size_t)n + 1;
        return s;
    }
};

std::string texture_dir(const std::string &scene_file) { // buildTextureDir: name without extension + "_textures/"
    const size_t slash = scene_file.find_last_of('/');
    const size_t dot = scene_file.find_last_of('.');
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
    return (has_ext ? scene_file.substr(0, dot) : scene_file) + "_textures/";
}

// vkr_parse_material_param_file: up to max_values floats, one per line; empty when the file does not exist
std::vector<float> param_file(const std::string &path, size_t max_values) {
    std::vector<float> v;
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return v;
    float x;
    while (v.size() < max_values && fscanf(f, "%f", &x) == 1) v.push_back(x);
    fclose(f);
    return v;
}

// .vkt -> VksTexture; false when the file does not exist (textures are optional)
bool read_vkt(const std::string &path, VksTexture &t, int32_t *vk_format) {
    bool missing = false;
    const std::vector<uint8_t> d = read_file(path, &missing);
    if (missing) return false;
    Cursor c{d, path};
    if (c.take<int32_t>() != VKT_MAGIC) throw std::runtime_error(path + " is not a .vkt file");
    if (c.take<int32_t>() != 1) throw std::runtime_error(path + ": unsupported texture version");
    const int32_t n_mips = c.take<int32_t>();
    t.width = c.take<int32_t>();
    t.height = c.take<int32_t>();
    *vk_format = c.take<int32_t>();
    const uint64_t data_size = c.take<uint64_t>();
    if (n_mips < 1 || n_mips > 31 || t.width < 1 || t.height < 1) throw std::runtime_error(path + ": invalid texture header");
    c.pos += (size_t)n_mips * 24; // per level: width, height, dataSize, dataOffset
    if (c.pos + data_size > d.size()) throw std::runtime_error(path + ": truncated texel data");
    t.mip_levels = n_mips;
    t.channels = 4;
    t.bytes.assign(d.begin() + (std::ptrdiff_t)c.pos, d.begin() + (std::ptrdiff_t)(c.pos + data_size));
    return true;
}

float texture_handle(uint32_t id, uint32_t channel) { // rendering/bsdfs/texture_channel_mask.h: sign bit | channel << 29 | id
    const uint32_t bits = RPTR_TEXTURED_PARAM_MASK | (channel << 29) | id;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

// vkr_dequantize_transform + the vks axis flip of AnimationData::dequantize -> 3 x 4 row-major object-to-world
void instance_transform(const uint8_t *rec, float *m34) {
    float tr[3], scaling;
    uint16_t qc[4];
    memcpy(tr, rec, 12);
    memcpy(&scaling, rec + 12, 4);
    memcpy(qc, rec + 16, 8);
    float q[4];
    for (int i = 0; i < 4; ++i) q[i] = (float)qc[i] * (2.0f / (float)0xffff) - 1.0f;
    q[3] = -q[3];
    const float xx = q[0] * q[0], xy = q[0] * q[1], xz = q[0] * q[2], xw = q[0] * q[3];
    const float yy = q[1] * q[1], yz = q[1] * q[2], yw = q[1] * q[3], zz = q[2] * q[2], zw = q[2] * q[3];
    float m[4][3]; // float matrix[4][3] of vkr.c: column i of the glm::mat4x3
    m[0][0] = 1.0f - 2.0f * (yy + zz); m[0][1] = 2.0f * (xy - zw); m[0][2] = 2.0f * (xz + yw);
    m[1][0] = 2.0f * (xy + zw); m[1][1] = 1.0f - 2.0f * (xx + zz); m[1][2] = 2.0f * (yz - xw);
    m[2][0] = 2.0f * (xz - yw); m[2][1] = 2.0f * (yz + xw); m[2][2] = 1.0f - 2.0f * (xx + yy);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m[i][j] *= scaling;
    for (int j = 0; j < 3; ++j) m[3][j] = tr[j];
    // flip (x, y, z) -> (-x, z, y): row 0 = -x row, row 1 = z row, row 2 = y row
    for (int c = 0; c < 4; ++c) {
        m34[0 * 4 + c] = -m[c][0];
        m34[1 * 4 + c] = m[c][2];
        m34[2 * 4 + c] = m[c][1];
    }
}
void flip_float_transform(const float *m43, float *m34) { // version 3: float[4][3] stored per instance
    for (int c = 0; c < 4; ++c) {
        m34[0 * 4 + c] = -m43[3 * c + 0];
        m34[1 * 4 + c] = m43[3 * c + 2];
        m34[2 * 4 + c] = m43[3 * c + 1];
    }
}

rptr_base_material default_material() { // BaseMaterial defaults (rendering/bsdfs/base_material.h.glsl)
    rptr_base_material m;
    memset(&m, 0, sizeof(m));
    m.base_color[0] = m.base_color[1] = m.base_color[2] = 0.9f;
    m.normal_map = -1;
    m.roughness = 1.0f; m.specular = 0.5f; m.clearcoat_gloss = 0.1f; m.ior = 1.5f;
    m.transmission_color[0] = m.transmission_color[1] = m.transmission_color[2] = 1.0f;
    return m;
}

struct MeshHeader {
    std::string name;
    float scale[3], offset[3];
    uint64_t flags = 0, vertex_buffer_offset = 0, n_tris = 0;
    uint32_t material_id_base = 0, n_materials_in_range = 0;
    int64_t lod_group = 0;
    std::vector<uint64_t> segment_tris;
    std::vector<int32_t> segment_material_base;
    size_t qverts = 0, qnuv = 0, material_ids = 0; // byte offsets into the file
    int material_id_size = 1;
};

} // namespace

void load_vks(const std::string &path, VksScene &s, bool ignore_textures, bool load_specularity) {
    s = VksScene();
    s.file = read_file(path);
    Cursor c{s.file, path};
    if (c.take<int32_t>() != VKR_MAGIC) throw std::runtime_error(path + " is not a .vks file");
    const int32_t version = c.take<int32_t>();
    if (version != 3 && version != 4) throw std::runtime_error(path + ": file version " + std::to_string(version) + " is not supported (3 and 4 are)");
    c.take<uint64_t>(); // flags
    const uint64_t header_size = c.take<uint64_t>(), data_offset = c.take<uint64_t>();
    const uint64_t n_meshes = c.take<uint64_t>(), n_instances = c.take<uint64_t>(), n_materials = c.take<uint64_t>();
    c.take<uint64_t>(); // numTriangles
    const uint64_t n_groups = c.take<uint64_t>();
    uint64_t n_lod_groups = 1, n_static = n_instances, n_animated = 0, n_frames = 1;
    int64_t lod_offset = 0, anim_offset = 0;
    if (version >= 4) {
        n_lod_groups = c.take<uint64_t>();
        lod_offset = c.take<int64_t>();
        c.take<uint64_t>(); c.take<int64_t>(); // bone index tuples
        c.take<float>(); c.take<float>();      // animation start / step
        n_frames = c.take<uint64_t>();
        n_static = c.take<uint64_t>();
        n_animated = c.take<uint64_t>();
        anim_offset = c.take<int64_t>();
    }
    if (!n_meshes || !n_instances || !n_groups || !n_lod_groups) throw std::runtime_error(path + ": invalid object counts");
    if (header_size != c.pos) throw std::runtime_error(path + ": mismatching header size");
    std::vector<MeshHeader> meshes((size_t)n_meshes);
    for (size_t i = 0; i < meshes.size(); ++i) {
        MeshHeader &m = meshes[i];
        for (int k = 0; k < 3; ++k) m.scale[k] = c.take<float>();
        for (int k = 0; k < 3; ++k) m.offset[k] = c.take<float>();
        m.flags = c.take<uint64_t>();
        const uint64_t header_end = c.take<uint64_t>();
        m.vertex_buffer_offset = c.take<uint64_t>();
        const uint64_t n_segments = c.take<uint64_t>();
        m.n_tris = c.take<uint64_t>();
        m.material_id_base = c.take<uint32_t>();
        m.n_materials_in_range = c.take<uint32_t>();
        if (version >= 4) { m.lod_group = c.take<int64_t>(); c.pos += 4 * 8; }
        else c.pos += 5 * 8;
        if (n_segments > (1u << 20)) throw std::runtime_error(path + ": invalid segment count");
        for (uint64_t j = 0; j < n_segments; ++j) m.segment_tris.push_back(c.take<uint64_t>());
        for (uint64_t j = 0; j < n_segments; ++j) m.segment_material_base.push_back(c.take<int32_t>());
        m.name = c.string();
        if (header_end != c.pos) throw std::runtime_error(path + ": mismatching header offset for mesh " + std::to_string(i));
        if (m.lod_group < 0 || (uint64_t)m.lod_group >= n_lod_groups) throw std::runtime_error(path + ": invalid LoD group for mesh " + std::to_string(i));
    }
    struct Inst { int32_t mesh_id; uint32_t transform_index; };
    std::vector<Inst> insts;
    std::vector<float> inline_transforms; // version 3
    for (uint64_t g = 0; g < n_groups; ++g) {
        c.take<uint32_t>(); // flags
        const int32_t mesh_id = c.take<int32_t>();
        const uint64_t header_end = c.take<uint64_t>(), group_data = c.take<uint64_t>(), n_in_group = c.take<uint64_t>();
        c.string();
        if (group_data != c.pos) throw std::runtime_error(path + ": mismatching data offset for instance group " + std::to_string(g));
        if (mesh_id < 0 || (uint64_t)mesh_id >= n_meshes) throw std::runtime_error(path + ": instance refers to a mesh that does not exist");
        for (uint64_t j = 0; j < n_in_group; ++j) {
            if (version >= 4) insts.push_back(Inst{mesh_id, c.take<uint32_t>()});
            else {
                insts.push_back(Inst{mesh_id, (uint32_t)(inline_transforms.size() / 12)});
                for (int k = 0; k < 12; ++k) inline_transforms.push_back(c.take<float>());
            }
        }
        if (header_end != c.pos) throw std::runtime_error(path + ": mismatching header offset for instance group " + std::to_string(g));
    }
    if (insts.size() != n_instances) throw std::runtime_error(path + ": instance count does not match the groups");
    std::vector<std::vector<int64_t>> lod_groups(1);
    if (version >= 4) {
        if ((uint64_t)lod_offset != c.pos) throw std::runtime_error(path + ": invalid LoD group offset");
        lod_groups.assign((size_t)n_lod_groups, {});
        for (auto &g : lod_groups) {
            const uint64_t n = c.take<uint64_t>();
            for (uint64_t k = 0; k < n; ++k) g.push_back(c.take<int64_t>());
            c.pos += (size_t)n * 4; // detail reduction
        }
    }
    if (data_offset != c.pos) throw std::runtime_error(path + ": mismatching body data offset");
    for (uint64_t i = 0; i < n_materials; ++i) s.material_names.push_back(c.string());
    size_t offset = c.pos;
    for (MeshHeader &m : meshes) {
        if (m.vertex_buffer_offset != offset) throw std::runtime_error(path + ": mismatching data offset for mesh " + m.name);
        m.qverts = offset; offset += 24 * (size_t)m.n_tris;
        m.qnuv = offset; offset += 24 * (size_t)m.n_tris;
        m.material_id_size = (m.n_materials_in_range <= 256 || m.segment_tris.size() > 1) ? 1 : 2;
        m.material_ids = offset; offset += (size_t)m.material_id_size * (size_t)m.n_tris;
        if (m.flags & 1u) offset += 12 * (size_t)m.n_tris; // vertex-sharing indices: not needed for unrolled triangles
        if (offset > s.file.size()) throw std::runtime_error(path + ": truncated mesh data");
    }
    // ---- Scene::load_vkrs: meshes -> geometries + parameterized meshes (:596-710) ----
    s.material_offsets.resize(meshes.size());
    for (size_t i = 0; i < meshes.size(); ++i) {
        const MeshHeader &m = meshes[i];
        rptr_mesh_desc md{(int32_t)s.geometries.size(), 0};
        uint64_t base = 0;
        for (size_t j = 0; j < m.segment_tris.size(); ++j) {
            const uint64_t n = m.segment_tris[j];
            if (n > 0) {
                rptr_geometry_desc g;
                memset(&g, 0, sizeof(g));
                // the streams sit at arbitrary byte offsets of the file: copied into aligned storage
                s.streams.emplace_back(3 * (size_t)n);
                memcpy(s.streams.back().data(), s.file.data() + m.qverts + 24 * (size_t)base, 24 * (size_t)n);
                s.streams.emplace_back(3 * (size_t)n);
                memcpy(s.streams.back().data(), s.file.data() + m.qnuv + 24 * (size_t)base, 24 * (size_t)n);
                g.qverts = nullptr; g.qnormal_uv = nullptr; // set below, once the vector of streams has stopped growing
                for (int k = 0; k < 3; ++k) { g.quantized_scaling[k] = m.scale[k]; g.quantized_offset[k] = m.offset[k]; }
                g.n_tris = (int32_t)n;
                g.has_normals = g.has_uvs = 1;
                s.geometries.push_back(g);
                md.n_geometries++;
                if (!(m.segment_tris.size() == 1 && m.n_materials_in_range > 1)) s.material_offsets[i].push_back(m.segment_material_base[j]);
            }
            base += n;
        }
        s.meshes.push_back(md);
        rptr_pmesh_desc pm;
        memset(&pm, 0, sizeof(pm));
        pm.mesh_id = (int32_t)i;
        if (m.segment_tris.size() == 1 && m.n_materials_in_range > 1) {
            if (m.material_id_size != 1) throw std::runtime_error(path + ": 16-bit material ids (deprecated in the format) are not supported");
            s.material_offsets[i].assign(1, (int32_t)m.material_id_base);
            pm.tri_material_ids = s.file.data() + m.material_ids;
            pm.n_tri_material_ids = (int64_t)m.n_tris;
        }
        s.pmeshes.push_back(pm);
    }
    for (size_t g = 0; g < s.geometries.size(); ++g) {
        s.geometries[g].qverts = s.streams[2 * g].data();
        s.geometries[g].qnormal_uv = s.streams[2 * g + 1].data();
    }
    for (size_t i = 0; i < s.pmeshes.size(); ++i) { // pointers after the vectors stopped growing
        s.pmeshes[i].material_offsets = s.material_offsets[i].data();
        s.pmeshes[i].n_material_offsets = (int32_t)s.material_offsets[i].size();
    }
    // ---- instances: base LoD level only (:733-755), frame 0 of the transform table ----
    const size_t n_transforms = (size_t)(n_static + n_animated * n_frames);
    if (version >= 4 && (size_t)anim_offset + TRANSFORM_BYTES * n_transforms > s.file.size()) throw std::runtime_error(path + ": truncated transform table");
    for (const Inst &in : insts) {
        const MeshHeader &m = meshes[(size_t)in.mesh_id];
        const std::vector<int64_t> &group = lod_groups[(size_t)m.lod_group];
        if (!group.empty() && group[0] != in.mesh_id) continue;
        rptr_instance_desc id;
        id.pmesh_id = in.mesh_id;
        if (version >= 4) {
            const uint64_t idx = in.transform_index < n_static ? in.transform_index : n_static + (in.transform_index - n_static); // frame 0
            if (idx >= n_transforms) throw std::runtime_error(path + ": transform index out of range");
            instance_transform(s.file.data() + (size_t)anim_offset + TRANSFORM_BYTES * (size_t)idx, id.transform);
        } else
            flip_float_transform(inline_transforms.data() + 12 * (size_t)in.transform_index, id.transform);
        s.instances.push_back(id);
        s.total_triangles += (int64_t)m.n_tris;
    }
    // ---- materials (:818-1003): three textures each, parameter files ----
    const std::string tdir = texture_dir(path);
    for (size_t i = 0; i < s.material_names.size(); ++i) {
        const std::string &name = s.material_names[i];
        rptr_base_material m = default_material();
        std::string extended = name;
        {
            bool missing = false;
            const std::vector<uint8_t> ex = read_file(tdir + name + "_Ex.txt", &missing);
            if (!missing) extended.assign(ex.begin(), ex.end());
        }
        auto image = [&](const char *kind, const uint8_t (&texel)[4], int32_t color_space, int forced_bc, bool *has_alpha) {
            VksTexture t;
            t.color_space = color_space;
            int32_t fmt = 0;
            if (!ignore_textures && read_vkt(tdir + name + "_" + kind + ".vkt", t, &fmt)) {
                int bc = 0;
                bool alpha = false;
                switch (fmt) { // VkFormat of the file -> Image::bcFormat (scene.cpp:836-860)
                    case 131: case 132: bc = 1; break;
                    case 133: case 134: bc = -1; alpha = true; break;
                    case 137: case 138: bc = 3; alpha = true; break;
                    case 141: bc = 5; break;
                    case 37: case 43: alpha = true; break;
                    default: break;
                }
                t.bc_format = forced_bc != 99 ? forced_bc : bc;
                if (has_alpha) *has_alpha = alpha;
            } else {
                t.bytes.assign(texel, texel + 4);
                if (has_alpha) *has_alpha = false;
            }
            s.textures.push_back(std::move(t));
            return (uint32_t)(s.textures.size() - 1);
        };
        bool has_alpha = false;
        const uint8_t white[4] = {255, 255, 255, 255}, flat[4] = {127, 127, 127, 255}, spec[4] = {255, 127, 0, 255};
        const uint32_t t_color = image("BaseColor", white, RPTR_COLOR_SPACE_SRGB, 99, &has_alpha);
        if (!has_alpha) m.flags |= RPTR_BASE_MATERIAL_NOALPHA;
        m.base_color[0] = texture_handle(t_color, 0);
        m.normal_map = (int32_t)image("Normal", flat, RPTR_COLOR_SPACE_LINEAR, 5, nullptr);
        const uint32_t t_spec = image("Specular", spec, RPTR_COLOR_SPACE_LINEAR, 1, nullptr);
        m.roughness = texture_handle(t_spec, 1);
        m.metallic = texture_handle(t_spec, 2);
        if (load_specularity) m.specular = texture_handle(t_spec, 0);
        const std::vector<float> em = param_file(tdir + name + "_EmissionIntensity.txt", 4);
        float intensity = 0.0f, color[3] = {0.0f, 0.0f, 0.0f};
        if (em.size() == 1) {
            intensity = em[0];
            const std::vector<float> bc = param_file(tdir + name + "_BaseColor.txt", 3);
            if (!bc.empty() && bc.size() != 3) throw std::runtime_error("three colour components expected for the emission base colour of " + name);
            for (size_t k = 0; k < bc.size(); ++k) color[k] = bc[k];
        } else if (em.size() == 4) {
            intensity = em[0];
            for (int k = 0; k < 3; ++k) color[k] = em[1 + k];
        } else if (!em.empty())
            throw std::runtime_error("one or four components expected for the emission of " + name);
        if (intensity > 0.0f) {
            if (color[0] != 0.0f || color[1] != 0.0f || color[2] != 0.0f)
                for (int k = 0; k < 3; ++k) m.base_color[k] = color[k];
            m.emission_intensity = intensity;
        }
        float tr[4] = {0.0f, 1.5f, 0.0f, 0.0f};
        const std::vector<float> trv = param_file(tdir + name + "_SpecularTransmission.txt", 4);
        for (size_t k = 0; k < trv.size(); ++k) tr[k] = trv[k];
        m.specular_transmission = tr[0];
        const bool two_sided = extended.find("twosided") != std::string::npos || extended.find("doublesided") != std::string::npos ||
                               extended.find("TwoSided") != std::string::npos || extended.find("DoubleSided") != std::string::npos;
        if (m.specular_transmission != 0.0f && !two_sided) m.flags |= RPTR_BASE_MATERIAL_ONESIDED;
        m.ior = tr[1];
        s.materials.push_back(m);
    }
    for (const VksTexture &t : s.textures) {
        rptr_texture_desc td;
        memset(&td, 0, sizeof(td));
        td.width = t.width; td.height = t.height; td.channels = t.channels; td.color_space = t.color_space;
        td.texels = t.bytes.data();
        td.bc_format = t.bc_format; td.mip_levels = t.mip_levels;
        s.texture_descs.push_back(td);
    }
}

rptr_scene_desc VksScene::desc() const {
    rptr_scene_desc d;
    memset(&d, 0, sizeof(d));
    d.geometries = geometries.data(); d.n_geometries = (int32_t)geometries.size();
    d.meshes = meshes.data(); d.n_meshes = (int32_t)meshes.size();
    d.pmeshes = pmeshes.data(); d.n_pmeshes = (int32_t)pmeshes.size();
    d.instances = instances.data(); d.n_instances = (int32_t)instances.size();
    d.materials = materials.data(); d.n_materials = (int32_t)materials.size();
    d.textures = texture_descs.data(); d.n_textures = (int32_t)texture_descs.size();
    return d;
}

uint64_t VksScene::hash() const {
    uint64_t h = 1469598103934665603ull;
    auto eat = [&](const void *p, size_t n) {
        const unsigned char *b = (const unsigned char *)p;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    for (const rptr_geometry_desc &g : geometries) {
        eat(g.qverts, 24 * (size_t)g.n_tris);
        eat(g.qnormal_uv, 24 * (size_t)g.n_tris);
        eat(g.quantized_scaling, 12);
        eat(g.quantized_offset, 12);
    }
    for (size_t i = 0; i < pmeshes.size(); ++i) {
        eat(material_offsets[i].data(), 4 * material_offsets[i].size());
        if (pmeshes[i].tri_material_ids) eat(pmeshes[i].tri_material_ids, (size_t)pmeshes[i].n_tri_material_ids);
    }
    for (const rptr_instance_desc &in : instances) {
        float tr[12];
        for (int k = 0; k < 12; ++k) tr[k] = in.transform[k] + 0.0f; // the sign of a zero is not part of the identity of a transform
        eat(&in.pmesh_id, 4);
        eat(tr, 48);
    }
    eat(materials.data(), materials.size() * sizeof(rptr_base_material));
    for (const VksTexture &t : textures) {
        const int32_t head[6] = {t.width, t.height, t.channels, t.color_space, t.bc_format, t.mip_levels};
        eat(head, sizeof(head));
        eat(t.bytes.data(), t.bytes.size());
    }
    return h;
}

} // namespace rptr_host
