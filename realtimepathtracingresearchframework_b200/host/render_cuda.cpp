// render_cuda.cpp -- see render_cuda.h.  Reference-side glue only; every arithmetic step of the hot path is behind the C ABI.
#include "render_cuda.h"
#include "lights.h"            // librender/lights.h: collect_emitters, update_light_sampling

#include <cstdlib>
#include <cstring>

#include "error_io.h"      // util/error_io.h: throw_error
#include "scene.h"         // librender/scene.h
#include "quantize.h"      // librender/quantize.h
#include "../rendering/lights/sky_model_arhosek/sky_model.h"
#include "../rendering/pointsets/sobol_tables.h" // SobolMatrix, SobolInversion_1_0 (vulkan/pointsets/render_sobol.cpp:14)
#include "../rendering/pointsets/bn_tables.h"    // sobol_256spp_256d, scramblingTile_yx_d_1spp (vulkan/pointsets/render_bn.cpp)

namespace glsl {
using namespace glm;
#include "../rendering/language.hpp"
#include "../rendering/color/color_matching.h"
#include "../rendering/color/color_matching.glsl"
}

static_assert(sizeof(rptr_base_material) == sizeof(BaseMaterial), "BaseMaterial layout");
static_assert(sizeof(rptr_render_params) == sizeof(RenderParams), "RenderParams layout");
static_assert(sizeof(rptr_light_sampling_config) == sizeof(LightSamplingConfig), "LightSamplingConfig layout");
static_assert(sizeof(rptr_render_ray_query) == sizeof(RenderRayQuery), "RenderRayQuery layout");
static_assert(sizeof(rptr_camera_params) == sizeof(RenderCameraParams), "RenderCameraParams layout");
static_assert(sizeof(rptr_backend_options) == sizeof(RenderBackendOptions), "RenderBackendOptions layout");
static_assert(offsetof(rptr_backend_options, render_upscale_factor) == offsetof(RenderBackendOptions, render_upscale_factor), "RenderBackendOptions layout");
static_assert(offsetof(rptr_backend_options, enable_raytraced_dof) == offsetof(RenderBackendOptions, enable_raytraced_dof), "RenderBackendOptions layout");
static_assert(offsetof(rptr_base_material, emission_intensity) == offsetof(BaseMaterial, emission_intensity), "BaseMaterial layout");
static_assert(offsetof(rptr_render_params, output_channel) == offsetof(RenderParams, output_channel), "RenderParams layout");

RenderCuda::RenderCuda(int device_ordinal) {
    if (rptr_cuda_create(device_ordinal, &ctx) != 0)
        throw_error("cuda backend: %s", rptr_cuda_last_error(nullptr));
}
RenderCuda::~RenderCuda() { rptr_cuda_destroy(ctx); }

void RenderCuda::check(int rc) const {
    if (rc != 0) throw_error("cuda backend: %s", rptr_cuda_last_error(ctx));
}

std::string RenderCuda::name() const { return rptr_cuda_name(); }

std::vector<std::string> const &RenderCuda::variant_names() const {
    static const std::vector<std::string> names = {"PT_WAVEFRONT"};
    return names;
}
int RenderCuda::variant_index(char const *name) { return std::strcmp(name, "PT_WAVEFRONT") == 0 ? 0 : -1; }

void RenderCuda::initialize(const int w, const int h) {
    fb_width = w;
    fb_height = h;
    // RenderVulkan::initialize sizes its LDR render targets by options.render_upscale_factor (vulkan/render_vulkan.cpp:255-263)
    check(rptr_cuda_set_option(ctx, "render_upscale_factor", options.render_upscale_factor < 1 ? 1 : options.render_upscale_factor));
#ifdef ENABLE_REALTIME_RESOLVE
    check(rptr_cuda_set_option(ctx, "realtime_resolve", 1)); // the temporal build: reproject_and_accumulate + the TAA step
#endif
    check(rptr_cuda_initialize(ctx, w, h));
}

// RenderBackend::create_processing_step (librender/render_backend.h:84; vulkan/render_vulkan_extensions.cpp:36-39): the application
// asks for the TAA step under ENABLE_REALTIME_RESOLVE (app.cpp:97-99) and runs it after end_frame (app.cpp:517-520).
namespace {
struct ProcessTAACuda : RenderExtension {
    RenderCuda *backend;
    explicit ProcessTAACuda(RenderCuda *b) : backend(b) {}
    std::string name() const override { return "CUDA TAA Processing Extension"; }
    void initialize(const int, const int) override {}
    void update_scene_from_backend(const Scene &) override {}
    void process(CommandStream *, int) override { backend->process_taa(); }
};
} // namespace
std::unique_ptr<RenderExtension> RenderCuda::create_processing_step(RenderProcessingStep step) {
    if (step == RenderProcessingStep::TAA) return std::unique_ptr<RenderExtension>(new ProcessTAACuda(this));
    return RenderBackend::create_processing_step(step);
}
void RenderCuda::process_taa() { check(rptr_cuda_process_taa(ctx)); }

// Scene -> rptr_scene_desc.  Geometry must be unindexed with quantised positions, which is what the reference's own
// backend requires as well (REQUIRE_UNROLLED_VERTICES / QUANTIZED_POSITIONS: vulkan/render_vulkan.cpp:575-596).
void RenderCuda::set_scene(const Scene &scene) {
    std::vector<rptr_geometry_desc> geoms;
    std::vector<rptr_mesh_desc> meshes;
    for (const Mesh &mesh : scene.meshes) {
        rptr_mesh_desc md{(int32_t)geoms.size(), (int32_t)mesh.geometries.size()};
        for (const Geometry &g : mesh.geometries) {
            if (!(g.format_flags & Geometry::QuantizedPositions) || !(g.format_flags & Geometry::ImplicitIndices))
                throw_error("cuda backend: expecting unindexed mesh data with quantized positions");
            rptr_geometry_desc gd{};
            gd.qverts = (const uint64_t *)g.vertices.data();
            gd.n_tris = g.num_tris();
            const bool qnuv = (g.format_flags & Geometry::QuantizedNormalsAndUV) != 0;
            if (!g.normals.empty() && !qnuv) throw_error("cuda backend: expecting quantized normals and uvs");
            gd.qnormal_uv = g.normals.empty() ? nullptr : (const uint64_t *)g.normals.data();
            gd.has_normals = !g.normals.empty();
            gd.has_uvs = !g.uvs.empty() || (qnuv && !g.normals.empty());
            for (int k = 0; k < 3; ++k) {
                gd.quantized_scaling[k] = g.quantized_scaling[k];
                gd.quantized_offset[k] = g.quantized_offset[k];
            }
            geoms.push_back(gd);
        }
        meshes.push_back(md);
    }
    std::vector<rptr_pmesh_desc> pmeshes;
    std::vector<std::vector<uint8_t>> ids8(scene.parameterized_meshes.size());
    for (size_t p = 0; p < scene.parameterized_meshes.size(); ++p) {
        const ParameterizedMesh &pm = scene.parameterized_meshes[p];
        rptr_pmesh_desc pd{};
        pd.mesh_id = pm.mesh_id;
        pd.n_material_offsets = (int32_t)pm.material_offsets.size();
        pd.material_offsets = pm.material_offsets.data();
        if (pm.per_triangle_materials()) { // the device keeps 8-bit ids like the reference (vulkan/render_vulkan.cpp:1060-1217)
            const len_t n = pm.num_triangle_material_ids();
            ids8[p].resize((size_t)n);
            for (len_t t = 0; t < n; ++t) ids8[p][(size_t)t] = (uint8_t)pm.triangle_material_id((index_t)t);
            pd.tri_material_ids = ids8[p].data();
            pd.n_tri_material_ids = (int64_t)n;
        }
        pmeshes.push_back(pd);
    }
    std::vector<rptr_instance_desc> instances;
    for (const Instance &inst : scene.instances) {
        rptr_instance_desc id{};
        id.pmesh_id = inst.parameterized_mesh_id;
        const glm::mat4 m = scene.animation_data.at(inst.animation_data_index).dequantize(inst.transform_index, 0); // frame 0
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) id.transform[4 * r + c] = m[c][r]; // 3x4 row-major, as handed to the TLAS (:1262-1268)
        instances.push_back(id);
    }
    // Scene::textures (util/image.h:10-27): every image as it is -- all mip levels back to back, raw RGBA8 or the block-compressed
    // payload of a .vks scene (bcFormat 1 / -1 / 3 / 5, librender/scene.cpp:836-930); the backend decodes the blocks on upload and
    // selects levels from the ray footprints like the megakernel's textureGrad reads (USE_MIPMAPPING)
    std::vector<rptr_texture_desc> textures;
    for (const Image &img : scene.textures) {
        rptr_texture_desc td{};
        td.width = img.width; td.height = img.height; td.channels = img.channels;
        td.color_space = img.color_space == SRGB ? RPTR_COLOR_SPACE_SRGB : RPTR_COLOR_SPACE_LINEAR;
        td.texels = img.img.data();
        td.bc_format = img.bcFormat;
        td.mip_levels = img.mip_levels();
        textures.push_back(td);
    }
    // emitters: the reference's own collection and binning (RenderBinnedLightsVulkan::update_scene_from_backend + update_lights,
    // vulkan/light_sampling/render_binned_lights.cpp:68-127, on librender/lights.cpp) -- the adapter links librender anyway, so the
    // backend receives the binned buffer instead of re-deriving it (its own restatement serves hosts without librender)
    BinnedLightSampling binned;
    {
        const std::vector<TriLight> emitters = collect_emitters(scene);
        if (!emitters.empty()) update_light_sampling(binned, emitters, lighting_params);
    }
    static_assert(sizeof(TriLight) == sizeof(rptr_tri_light_data), "TriLightData layout");
    rptr_scene_desc d{};
    d.binned_lights = binned.emitters.empty() ? nullptr : reinterpret_cast<const rptr_tri_light_data *>(binned.emitters.data());
    d.n_binned_lights = (int32_t)binned.emitters.size();
    d.textures = textures.data(); d.n_textures = (int32_t)textures.size();
    d.geometries = geoms.data(); d.n_geometries = (int32_t)geoms.size();
    d.meshes = meshes.data(); d.n_meshes = (int32_t)meshes.size();
    d.pmeshes = pmeshes.data(); d.n_pmeshes = (int32_t)pmeshes.size();
    d.instances = instances.data(); d.n_instances = (int32_t)instances.size();
    d.materials = (const rptr_base_material *)scene.materials.data(); d.n_materials = (int32_t)scene.materials.size();
    check(rptr_cuda_set_scene(ctx, &d, (const rptr_light_sampling_config *)&lighting_params));
    has_lights = rptr_cuda_get_lights(ctx, nullptr, 0) > 0;
}

// RenderVulkan::update_config + update_sky_light (vulkan/render_vulkan.cpp:2954-2959, vulkan/render_sky.cpp:25-72):
// the fit runs here, on the reference's side of the boundary, with librender's own sky_model.cpp and colour tables.
void RenderCuda::update_config(SceneConfig const &config) {
    rptr_scene_params sp{};
    glm::vec3 sun_dir = glm::normalize(config.sun_dir);
    ArHosekSkyModelState state;
    arhosek_rgb_skymodelstate_alloc_init(config.turbidity, dot(config.albedo, glm::vec3(0.3333f)), sun_dir.y, &state);
    for (int k = 0; k < 3; ++k) sp.sun_dir[k] = sun_dir[k];
    sp.sun_cos_angle = std::cos(glm::radians(0.53f) / 2.0f);
    for (int i = 0; i < 9; ++i)
        for (int k = 0; k < 3; ++k) sp.sky_configs[i][k] = (float)state.configs[k][i];
    for (int k = 0; k < 3; ++k) sp.sky_radiances[k] = (float)state.radiances[k];
    ArHosekSkyModelState sunState;
    arhosekskymodelstate_alloc_init(state.elevation, state.turbidity, state.albedo, &sunState);
    glm::vec3 xyz(0.0f);
    int numSamples = 0;
    float last_wavelength = CM_CIE_MIN;
    for (int i = 0; i < CM_CIE_SAMPLES; ++i) {
        float wavelength = float(i) * float(CM_CIE_MAX - CM_CIE_MIN) / float(CM_CIE_SAMPLES - 1) + float(CM_CIE_MIN);
        if (wavelength > 720.0f) break;
        float radiance = (float)arhosekskymodel_solar_radiance(&sunState, sun_dir.y, 0.0, wavelength);
        radiance -= (float)arhosekskymodel_radiance(&sunState, sun_dir.y, 0.0, wavelength);
        {
            using namespace glsl;
            xyz += glm::vec3(CM_TABLE_X[i], CM_TABLE_Y[i], CM_TABLE_Z[i]) * radiance;
        }
        ++numSamples;
        last_wavelength = wavelength;
    }
    xyz *= float(last_wavelength - CM_CIE_MIN) / float(numSamples);
    if (sun_dir.y > 0.0f && glm::all(glm::greaterThanEqual(xyz, glm::vec3(0.0f)))) {
        glm::vec3 rgb = 0.01f * glsl::xyz_to_srgb(xyz);
        for (int k = 0; k < 3; ++k) sp.sun_radiance[k] = rgb[k];
        sp.sun_radiance[3] = 1.0f; // the light-count rule (:67-70) is applied by the backend, which knows light_count
    }
    sp.normal_z_scale = 1.0f / config.bump_scale;
    check(rptr_cuda_set_scene_params(ctx, &sp));
}

// normalize_options / configure_for (librender/render_backend.h:84-85; vulkan/render_vulkan.cpp:1878-1917): options the backend
// cannot honour make configure_for return false with the recovery mask filled in, which sends the app through its fallback
// (app.cpp:400-431) instead of rendering something else than was asked for.
void RenderCuda::normalize_options(RenderBackendOptions &rbo, int variant_idx) const {
    if (rptr_cuda_normalize_options(ctx, reinterpret_cast<rptr_backend_options *>(&rbo), variant_idx) != 0)
        throw_error("cuda backend: %s", rptr_cuda_last_error(ctx));
}
bool RenderCuda::configure_for(RenderBackendOptions const &rbo, int variant_idx, AvailableRenderBackendOptions *available_recovery_options) {
    if (rbo.rng_variant != RNG_VARIANT_UNIFORM) { // the tables have to be on the device before the variant can be selected
        RenderBackendOptions saved = options;
        options.rng_variant = rbo.rng_variant;
        applied_rng_variant = -1;
        apply_rng_variant();
        options = saved;
    }
    rptr_backend_options closest;
    const bool ok = rptr_cuda_configure_for(ctx, reinterpret_cast<const rptr_backend_options *>(&rbo), variant_idx, &closest) == 0;
    if (available_recovery_options) { // an option is "available" when its requested value is one the backend renders with
        const RenderBackendOptions &c = reinterpret_cast<const RenderBackendOptions &>(closest);
#define RPTR_AVAILABLE(type, name, default_, flags) available_recovery_options->name = c.name == rbo.name;
        RENDER_BACKEND_OPTIONS(RPTR_AVAILABLE)
#undef RPTR_AVAILABLE
    }
    if (!ok) println(CLL::WARNING, "cuda backend: %s", rptr_cuda_last_error(ctx));
    else applied_rng_variant = rbo.rng_variant;
    return ok;
}

void RenderCuda::enable_ray_queries(const int max_queries, const int max_queries_per_pixel) {
    check(rptr_cuda_enable_ray_queries(ctx, max_queries, max_queries_per_pixel));
}
bool RenderCuda::render_ray_queries(int num_queries, const RenderParams &params_, int variant_idx, CommandStream *) {
    check(rptr_cuda_render_ray_queries(ctx, num_queries, (const rptr_render_params *)&params_, variant_idx));
    return true;
}
void RenderCuda::write_ray_queries(const RenderRayQuery *queries, int first, int count) {
    check(rptr_cuda_write_ray_queries(ctx, (const rptr_render_ray_query *)queries, first, count));
}
void RenderCuda::read_ray_results(glm::vec4 *results, int first, int count) { check(rptr_cuda_read_ray_results(ctx, (float *)results, first, count)); }
void RenderCuda::ray_query_buffers(void **device_queries, void **device_results, size_t *capacity) {
    check(rptr_cuda_ray_query_buffers(ctx, device_queries, device_results, capacity));
}

// options.rng_variant (librender/render_params.glsl.h:76) + the tables the reference's pointset extensions upload
// (RenderSobolVulkan::update_random_buf, RenderBNPointsVulkan::update_random_buf)
void RenderCuda::apply_rng_variant() {
    const int v = options.rng_variant;
    if (v == applied_rng_variant) return;
    static_assert(sizeof(SobolMatrix[0]) == 4 && sizeof(SobolInversion_1_0[0]) == 4 && sizeof(sobol_256spp_256d[0]) == 4, "table element size");
    if (v == RNG_VARIANT_SOBOL || v == RNG_VARIANT_Z_SBL) {
        check(rptr_cuda_set_pointset_table(ctx, RPTR_POINTSET_SOBOL_MATRIX, (const uint32_t *)SobolMatrix, sizeof(SobolMatrix) / 4));
        check(rptr_cuda_set_pointset_table(ctx, RPTR_POINTSET_SOBOL_TILE_INVERT, (const uint32_t *)SobolInversion_1_0, sizeof(SobolInversion_1_0) / 4));
    } else if (v == RNG_VARIANT_BN) {
        check(rptr_cuda_set_pointset_table(ctx, RPTR_POINTSET_BN_SOBOL, (const uint32_t *)sobol_256spp_256d, sizeof(sobol_256spp_256d) / 4));
        check(rptr_cuda_set_pointset_table(ctx, RPTR_POINTSET_BN_SCRAMBLING_1SPP, (const uint32_t *)scramblingTile_yx_d_1spp,
                                           sizeof(scramblingTile_yx_d_1spp) / 4));
    }
    check(rptr_cuda_set_option(ctx, "rng_variant", v));
    applied_rng_variant = v;
}

void RenderCuda::begin_frame(CommandStream *, const RenderConfiguration &config) {
    apply_rng_variant();
    this->camera = config.camera;
    this->time = config.time;
    this->reset_accumulation = config.reset_accumulation;
    this->freeze_frame = config.freeze_frame;
    check(rptr_cuda_begin_frame(ctx, (const rptr_camera_params *)&config.camera, (const rptr_render_params *)&params,
                                (const rptr_light_sampling_config *)&lighting_params, config.reset_accumulation, config.freeze_frame, config.time));
}
void RenderCuda::draw_frame(CommandStream *, int variant_idx) { check(rptr_cuda_draw_frame(ctx, variant_idx)); }
void RenderCuda::end_frame(CommandStream *, int variant_idx) { check(rptr_cuda_end_frame(ctx, variant_idx)); }

RenderStats RenderCuda::render(const RenderConfiguration &config) {
    begin_frame(nullptr, config);
    draw_frame(nullptr, config.active_variant);
    end_frame(nullptr, config.active_variant);
    return stats();
}

RenderStats RenderCuda::stats() {
    rptr_render_stats s{};
    check(rptr_cuda_stats(ctx, &s));
    RenderStats r;
    r.render_time = s.render_time;
    r.rays_per_second = s.rays_per_second;
    r.spp = s.spp;
    r.frame_stats_delay = s.frame_stats_delay;
    r.has_valid_frame_stats = s.has_valid_frame_stats != 0;
    r.total_device_bytes_allocated = (size_t)s.total_device_bytes_allocated;
    r.max_device_bytes_allocated = (size_t)s.max_device_bytes_allocated;
    r.device_bytes_currently_allocated = (size_t)s.device_bytes_currently_allocated;
    return r;
}
void RenderCuda::flush_pipeline() { check(rptr_cuda_flush(ctx)); }

glm::uvec3 RenderCuda::get_framebuffer_size() const { // the (upscaled) LDR target, as RenderVulkan::get_framebuffer_size
    uint32_t w = 0, h = 0, c = 0;
    check(rptr_cuda_framebuffer_size(ctx, &w, &h, &c));
    return glm::uvec3(w, h, c);
}
size_t RenderCuda::readback_framebuffer(size_t bufferSize, unsigned char *buffer, bool) { return rptr_cuda_readback_u8(ctx, bufferSize, buffer); }
size_t RenderCuda::readback_framebuffer(size_t bufferSize, float *buffer, bool) { return rptr_cuda_readback_f32(ctx, bufferSize, buffer); }
size_t RenderCuda::readback_aov(AOVBufferIndex aovIndex, size_t bufferSize, uint16_t *buffer, bool) {
    return rptr_cuda_readback_aov(ctx, (int32_t)aovIndex, bufferSize, buffer);
}

int RenderCuda::trace_ray(const RenderRayQuery *queries, int num_queries, glm::vec4 *results) {
    check(rptr_cuda_trace_rays(ctx, (const rptr_render_ray_query *)queries, num_queries, (float *)results, nullptr));
    return num_queries;
}

// the device is chosen like CUDA applications usually are: RPTR_CUDA_DEVICE (ordinal among the visible devices), default 0
RenderBackend *create_cuda_backend(Display &) {
    const char *dev = std::getenv("RPTR_CUDA_DEVICE");
    return new RenderCuda(dev ? std::atoi(dev) : 0);
}
