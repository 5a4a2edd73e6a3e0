// render_cuda.h -- the reference-side adapter: `RenderCuda : RenderBackend, RaytraceBackend` over librptr_cuda.so.
//
// This file is meant to be dropped into the reference tree as cuda/render_cuda.h (the directory the reference's build
// already expects: CMakeLists.txt:148-152 `add_subdirectory(cuda)` / `render_backends -> render_cuda`) and compiles
// against the reference's own headers.  It contains no rendering code: it flattens `Scene` into rptr_scene_desc, runs
// the reference's own sky fit and forwards the frame protocol to the C ABI (include/rptr_cuda.h).
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "render_backend.h"   // librender/render_backend.h:68-116
#include "rptr_cuda.h"        // this repository: include/rptr_cuda.h

struct Scene;
struct Display;

struct RenderCuda : RenderBackend {
    explicit RenderCuda(int device_ordinal = 0);
    ~RenderCuda() override;

    std::string name() const override;
    void initialize(const int fb_width, const int fb_height) override;
    std::unique_ptr<RenderExtension> create_processing_step(RenderProcessingStep step) override;
    void process_taa(); // ProcessTAAVulkan::process
    std::vector<std::string> const &variant_names() const override;
    int variant_index(char const *name) override;

    void set_scene(const Scene &scene) override;
    void update_config(SceneConfig const &config) override;
    void normalize_options(RenderBackendOptions &rbo, int variant_idx) const override;
    bool configure_for(RenderBackendOptions const &rbo, int variant_idx, AvailableRenderBackendOptions *available_recovery_options = nullptr) override;

    void begin_frame(CommandStream *cmd_stream, const RenderConfiguration &config) override;
    void draw_frame(CommandStream *cmd_stream, int variant_idx = 0) override;
    void end_frame(CommandStream *cmd_stream, int variant_idx = 0) override;
    RenderStats stats() override;
    void flush_pipeline() override;

    // path tracing on caller-supplied rays (librender/render_backend.h:101-102; app.cpp:77-79 enables it under ENABLE_CUDA).
    // The reference keeps queries / results in device buffers its data-capture module fills; here they are reachable as
    // device addresses (ray_query_buffers) or through the two host-side copies.
    void enable_ray_queries(const int max_queries = DEFAULT_RAY_QUERY_BUDGET, const int max_queries_per_pixel = 0) override;
    bool render_ray_queries(int num_queries, const RenderParams &params, int variant_idx = 0, CommandStream *cmd_stream = nullptr) override;
    void write_ray_queries(const RenderRayQuery *queries, int first, int count);
    void read_ray_results(glm::vec4 *results, int first, int count);
    void ray_query_buffers(void **device_queries, void **device_results, size_t *capacity);

    glm::uvec3 get_framebuffer_size() const override;
    size_t readback_framebuffer(size_t bufferSize, unsigned char *buffer, bool force_refresh = false) override;
    size_t readback_framebuffer(size_t bufferSize, float *buffer, bool force_refresh = false) override;
    size_t readback_aov(AOVBufferIndex aovIndex, size_t bufferSize, uint16_t *buffer, bool force_refresh = false) override;

    // RaytraceBackend::trace_ray with the RenderRayQuery wire format (librender/render_params.glsl.h:165-170,
    // vulkan/rt_intersect.comp:53-67); the rt_datacapture types of librender/raytrace_backend.h are not in the release
    int trace_ray(const RenderRayQuery *queries, int num_queries, glm::vec4 *results);

protected:
    RenderStats render(const RenderConfiguration &config) override;

private:
    void check(int rc) const;
    void apply_rng_variant();
    int applied_rng_variant = RNG_VARIANT_UNIFORM;
    rptr_ctx *ctx = nullptr;
    int fb_width = 0, fb_height = 0;
    bool has_lights = false;
};

// typedef RenderBackend* (*create_backend_function)(Display&)  (librender/render_backend.h:118-119)
RenderBackend *create_cuda_backend(Display &display);
