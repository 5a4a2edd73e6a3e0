/* rptr_cuda.h -- C ABI of librptr_cuda.so, the B200 (sm_100a) wavefront path tracer that sits behind the
 * reference's RenderBackend / RaytraceBackend plugin surface (`rptr --backend cuda`).
 *
 * Every entry point is what the reference-side adapter `RenderCuda : RenderBackend, RaytraceBackend`
 * binds; the comment on each names the reference interface it replaces (paths relative to the reference tree).
 * Conventions: plain pointers and sizes only; int return, 0 = ok, non-zero = error with a message available from
 * rptr_cuda_last_error(); nothing throws across the boundary (the adapter turns errors into throw_error(),
 * util/error_io.h:27-30).  All host pointers are borrowed for the duration of the call only (the reference destroys
 * its Scene right after set_scene, app.cpp:151-175).  One context = one GPU; calls on a context must come from one
 * thread at a time (the reference's render thread, SURVEY 8b "Threading").
 */
#ifndef RPTR_CUDA_H
#define RPTR_CUDA_H

#include "rptr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rptr_ctx rptr_ctx;

/* Device counters accumulated since the last rptr_cuda_reset_counters(): the inputs of the roofline formula in
 * DESIGN.md section 7 (the reference's REPORT_RAY_STATS counters are commented out, vulkan/render_vulkan.cpp:2191-2225). */
typedef struct rptr_counters {
    uint64_t samples;          /* pixel samples completed */
    uint64_t closest_rays;     /* closest-hit rays traced */
    uint64_t shadow_rays;      /* any-hit rays traced */
    uint64_t shaded_vertices;  /* path vertices shaded (hits) */
    uint64_t closest_nodes;    /* BVH nodes fetched by closest-hit rays */
    uint64_t closest_tris;     /* triangles tested by closest-hit rays */
    uint64_t shadow_nodes;
    uint64_t shadow_tris;
    uint64_t launches;         /* kernels launched by this library */
    double ms_trace;           /* device time (CUDA events) in the closest-hit stage; needs option "stage_timing" */
    double ms_shadow;
    double ms_shade;
    double ms_other;           /* raygen + resolve */
    uint64_t trace_launches;   /* number of closest-hit kernel launches timed in ms_trace */
    uint64_t node_bytes;       /* size of one BVH node record fetched per node visit */
    uint64_t tri_bytes;        /* size of one traversal triangle record fetched per triangle test */
    uint64_t bvh_nodes;        /* number of (4-wide) BVH nodes of the current scene */
    double bvh_build_ms;       /* wall time of the last BVH build (host SAH or device builder) */
    uint64_t trace_overlap;    /* 1: option "overlap_shadow" is on -- ms_trace / trace_launches then cover closest-hit AND shadow
                                  launches (each shadow launch shares the GPU with the next closest-hit launch) and ms_shadow is 0 */
    uint64_t num_sms;          /* streaming multiprocessors of the device (persistent grids are sized in multiples of it) */
} rptr_counters;

/* create_cuda_backend(Display&) / ~RenderBackend  (librender/render_backend.h:118-119, main.cpp:273-285).
 * Fails (non-zero, *out = NULL) when no CUDA device is usable: there is no CPU fallback. */
int rptr_cuda_create(int device_ordinal, rptr_ctx **out);
void rptr_cuda_destroy(rptr_ctx *ctx);
/* last error message of ctx (or of the failed rptr_cuda_create when ctx == NULL) */
const char *rptr_cuda_last_error(const rptr_ctx *ctx);
/* RenderBackend::name() (librender/render_backend.h:79) */
const char *rptr_cuda_name(void);

/* RenderBackend::initialize(fb_width, fb_height) (librender/render_backend.h:86; vulkan/render_vulkan.cpp:245-249):
 * (re)allocates the RGBA32F accumulator, zeroes frame_id and frame_offset. */
int rptr_cuda_initialize(rptr_ctx *ctx, int32_t width, int32_t height);

/* RenderBackend::set_scene(const Scene&) + RenderExtension::update_scene_from_backend of the binned-lights extension
 * (librender/render_backend.h:92; vulkan/render_vulkan.cpp:1554-1644; vulkan/light_sampling/render_binned_lights.cpp:68-149).
 * Copies everything to the device, builds the BVH, collects+bins emitters unless desc->binned_lights is given. */
int rptr_cuda_set_scene(rptr_ctx *ctx, const rptr_scene_desc *desc, const rptr_light_sampling_config *lighting);
/* binned TriLightData[] in NEE order (the LIGHTS_BIND_POINT buffer); returns the count, copies up to max_lights */
int32_t rptr_cuda_get_lights(rptr_ctx *ctx, rptr_tri_light_data *out, int32_t max_lights);

/* RenderBackend::update_config(SceneConfig) after the host-side sky fit (vulkan/render_vulkan.cpp:2954-2959,
 * vulkan/render_sky.cpp:25-72).  sun_radiance[3] is 1 (sun up) or 0 as the fit leaves it; the light-count rule of
 * render_sky.cpp:67-70 is applied by the backend. */
int rptr_cuda_set_scene_params(rptr_ctx *ctx, const rptr_scene_params *params);

/* Backend options that are compile-time switches or host options in the reference:
 *   "transmission"  0/1  GLTF_SUPPORT_TRANSMISSION[_ROUGHNESS] (off in the megakernel build, rendering/bsdfs/gltf_bsdf.glsl:10-13)
 *   "rng_variant"   RenderBackendOptions::rng_variant (librender/render_params.glsl.h:34-37,76): 0 UNIFORM (LCG), 1 BN, 2 SOBOL,
 *                   3 Z_SBL; 1-3 need their tables (rptr_cuda_set_pointset_table) before the next draw_frame
 *   "wave_paths"    max paths in flight per wavefront pass (memory/occupancy knob)
 *   "aov_buffers"   0/1  write the fp16 AOV images (default 1: ENABLE_AOV_BUFFERS, vulkan/gpu_params.glsl:19)
 *   "overlap_shadow" 0/1 (default 1) launch the shadow rays of bounce d on a second stream beside the closest-hit rays of bounce
 *                   d + 1 (independent work; the persistent grids interleave SM by SM, hiding each other's tails)
 *   "tail_kernel"   0/1 (default 1) a drained warp of the persistent trace kernel hands its last rays to k_trace_tail, which walks each
 *                   with eight lanes breadth-wise (csrc/rptr_trace_tail.cuh); images are identical either way
 *   "reorder_bounce", "reorder_shadow"  0 (default) = trace the queues in screen order; 1..4 = bin the bounce / shadow queue by a
 *                   12-bit key of ray origin and direction first (csrc/rptr_reorder.cuh), -1 = key chosen per scene.  Images are
 *                   identical either way; measured without gain on the target scenes (profiles/r02_sweeps.md)
 *   "concurrent_waves" 1..4 (default 1) sub-waves of whole sample layers side by side on their own streams
 *   "stage_timing"  0/1  time each stage with CUDA events into rptr_counters.ms_*
 *   "bvh_builder"   1 (default) = binned-SAH build on the device, 0 = the same algorithm on the host inside set_scene, kept as
 *                   the A/B (both replace the driver's BLAS/TLAS build, vulkan/vulkanrt_utils.cpp:82-167; images are identical
 *                   either way; a device build that fails -- memory, depth -- falls back to the host builder)
 *   "trace_kernel"  0 = persistent speculative traversal kernel, 1 = one-ray-per-thread kernel (A/B reference)
 *   "tile_rank", "tile_world", "tile_rows": screen-space sharding across GPUs (interleaved bands of tile_rows rows)
 *   "render_upscale_factor" RenderBackendOptions::render_upscale_factor (also through rptr_cuda_configure_for): the LDR target is
 *                   that many times the render size (vulkan/render_vulkan.cpp:255-263); takes effect at the next initialize
 *   "realtime_resolve" 0/1 (default 0) the reference's ENABLE_REALTIME_RESOLVE build (CMakeLists.txt:98, OFF by default): with it
 *                   reprojection_mode = ACCUMULATE runs reproject_and_accumulate (rendering/postprocess/reprojection.glsl) in place of
 *                   the running mean -- the accumulator's alpha then holds 1 - sample weight -- and rptr_cuda_process_taa exists.
 *                   Needs the AOV images and the whole frame on one GPU.  Without it ACCUMULATE is the running mean, as in the
 *                   reference's default build (vulkan/render_vulkan.cpp:1911-1915).
 */
int rptr_cuda_set_option(rptr_ctx *ctx, const char *name, int64_t value);

/* RenderSobolVulkan::update_random_buf / RenderBNPointsVulkan::update_random_buf (vulkan/pointsets/render_sobol.cpp:77-104,
 * render_bn.cpp:77-126): hands the reference's sampler tables to the backend.  table = RPTR_POINTSET_*; count must be the
 * table's element count; the data is copied to the device. */
int rptr_cuda_set_pointset_table(rptr_ctx *ctx, int32_t table, const uint32_t *data, size_t count);

/* RenderBackend::begin_frame / draw_frame / end_frame (librender/render_backend.h:97-99;
 * vulkan/render_vulkan.cpp:1919-2002, 2157-2178, 2017-2155).  begin_frame applies the counter protocol
 * (reset: frame_offset += frame_id unless frozen, frame_id = 0) and update_view_parameters (:2880-2941);
 * draw_frame renders params.batch_spp sample layers; end_frame resolves and advances frame_id. Asynchronous. */
int rptr_cuda_begin_frame(rptr_ctx *ctx, const rptr_camera_params *camera, const rptr_render_params *params,
                          const rptr_light_sampling_config *lighting, int32_t reset_accumulation, int32_t freeze_frame, double time);
int rptr_cuda_draw_frame(rptr_ctx *ctx, int32_t variant);
int rptr_cuda_end_frame(rptr_ctx *ctx, int32_t variant);

/* RenderBackend::stats() / flush_pipeline() (librender/render_backend.h:109-110; vulkan/render_vulkan.cpp:2229-2248). Synchronise. */
int rptr_cuda_stats(rptr_ctx *ctx, rptr_render_stats *out);
int rptr_cuda_flush(rptr_ctx *ctx);
int rptr_cuda_get_counters(rptr_ctx *ctx, rptr_counters *out);
int rptr_cuda_reset_counters(rptr_ctx *ctx);
/* frame_id / frame_offset / accumulated_spp as the reference keeps them (vulkan/render_vulkan.h:166-168) */
int rptr_cuda_frame_state(rptr_ctx *ctx, uint32_t *frame_id, uint32_t *frame_offset, uint32_t *accumulated_spp);

/* RenderGraphic::get_framebuffer_size / readback_framebuffer(float*) / (unsigned char*)
 * (util/display/render_graphic.h:27-37; vulkan/render_vulkan.cpp:2250-2287): RGBA, row-major, top row first.
 * Return the number of elements written or 0 when the buffer is too small.  As in the reference the float image has the
 * render size (width*height*4 elements) while framebuffer_size and the 8-bit image are those of the LDR render target,
 * render_upscale_factor times larger in x and y: factor 2 replicates every pixel 2 x 2, any other factor > 1 stores the pixels
 * at their own coordinates and leaves the rest of the target untouched (vulkan/process_samples.comp:192-199, as written). */
int rptr_cuda_framebuffer_size(rptr_ctx *ctx, uint32_t *width, uint32_t *height, uint32_t *channels);
size_t rptr_cuda_readback_f32(rptr_ctx *ctx, size_t n_elems, float *dst);
size_t rptr_cuda_readback_u8(rptr_ctx *ctx, size_t n_elems, uint8_t *dst);
/* RenderGraphic::readback_aov (util/display/render_graphic.h:39-43; vulkan/render_vulkan.cpp:2290-2294): the RGBA16F AOV
 * images the megakernel stores for the first path vertex (vulkan/accumulate.glsl:77-103): aov_index 0 = albedo.rgb +
 * roughness (AOVAlbedoRoughnessIndex), 1 = normal.xyz + depth (AOVNormalDepthIndex), 2 = screen-space motion.xy against the
 * previous begin_frame's view + screen_jitter.xy (AOVMotionJitterIndex; geometry is static, so the motion is the camera's;
 * NaN before a second begin_frame exists, like the reference's zero VP_reference), as half-float bit patterns, RGBA,
 * row-major, top row first.  Every sample layer overwrites them; they hold the frame's last layer.  Returns the number of
 * elements written (width*height*4) or 0: buffer too small, option "aov_buffers" off, or aov_index outside 0..2. */
size_t rptr_cuda_readback_aov(rptr_ctx *ctx, int32_t aov_index, size_t n_elems, uint16_t *dst);
/* device address of the RGBA32F accumulator (for the multi-GPU reduce over NCCL); valid until initialize/destroy */
int rptr_cuda_framebuffer_device_ptr(rptr_ctx *ctx, void **ptr);
/* the cudaStream_t all work of this context is enqueued on (CommandStream of util/device_backend.h:14-22): lets a
 * caller bracket frames with its own events or order a collective after end_frame without a host sync */
int rptr_cuda_stream_handle(rptr_ctx *ctx, void **stream);

/* RaytraceBackend::trace_ray / RQ_CLOSEST (librender/raytrace_backend.h:18; vulkan/rt_intersect.comp:30-68): opaque closest
 * hit over (RAY_EPSILON * |origin|, t_max).  results[i] = (bary.x, bary.y, bits(instance+geometry index), bits(primitive
 * index)); miss = (-1, -1, bits(-1), bits(-1)); a query with mode_or_data < 0 is skipped and its slot left as the caller
 * passed it.  hit_t (optional) receives the hit distance or -1.  Host buffers. */
int rptr_cuda_trace_rays(rptr_ctx *ctx, const rptr_render_ray_query *queries, int32_t n, float *results, float *hit_t);

/* RenderBackend::enable_ray_queries(max_queries, max_queries_per_pixel) (librender/render_backend.h:101;
 * vulkan/render_vulkan.cpp:430-455; app.cpp:77-79 calls it with (DEFAULT_RAY_QUERY_BUDGET, 2)): sizes the device-side
 * ray_query_buffer / ray_result_buffer for max(width * height * max_queries_per_pixel, max_queries) queries; initialize()
 * re-sizes them for the new frame (:366-369).  The result buffer starts zeroed. */
int rptr_cuda_enable_ray_queries(rptr_ctx *ctx, int32_t max_queries, int32_t max_queries_per_pixel);
/* The reference keeps queries and results in device buffers that the producer of the queries fills (its data-capture module,
 * which is not part of the tree): these are the device addresses (RenderRayQuery[capacity], vec4[capacity]) for a CUDA-side
 * producer, and the host-side copies for everybody else.  Valid until enable_ray_queries / initialize / destroy. */
int rptr_cuda_ray_query_buffers(rptr_ctx *ctx, void **queries, void **results, size_t *capacity);
int rptr_cuda_write_ray_queries(rptr_ctx *ctx, const rptr_render_ray_query *queries, int32_t first, int32_t n);
int rptr_cuda_read_ray_results(rptr_ctx *ctx, float *results /* 4 per query */, int32_t first, int32_t n);
/* RenderBackend::render_ray_queries(num_queries, params, variant_idx, cmd_stream) (librender/render_backend.h:102;
 * vulkan/render_vulkan.cpp:1867-1876, 2961-3060; vulkan/pt_megakernel.glsl:276-283, 297-301, 327-334): the path tracer on the
 * first num_queries rays of the query buffer instead of camera rays -- origin, direction and t_max from the query, t_min = 0 --
 * with the view (frame_dims, frame_id, frame_offset) and RenderParams of the last begin_frame; `params` is accepted and
 * ignored exactly like the reference ignores it.  Query q is invocation gl_GlobalInvocationIndex = q of a 2-D dispatch of
 * ceil(sqrt(n)) columns in 32 x 16 workgroups and is seeded like the (swizzled) pixel of that invocation
 * (vulkan/setup_pixel_assignment.glsl:17-22).  batch_spp layers with sample indices 0 .. batch_spp - 1
 * (accumulation_frame_offset = 0) are folded into ray_results[q] by accumulate_query (vulkan/accumulate.glsl:32-42) in sample
 * order: layer 0 stores its sample, layer k > 0 ADDS the updated mean to the stored value (kept as the reference writes it;
 * with the usual batch_spp = 1 the result is simply the sample).  Does not touch the frame counters.  Asynchronous.
 * Not done: the reference's megakernel also writes the AOV images of the virtual pixels that fall inside the frame. */
int rptr_cuda_render_ray_queries(rptr_ctx *ctx, int32_t num_queries, const rptr_render_params *params, int32_t variant);

/* RenderBackend::normalize_options / configure_for (librender/render_backend.h:84-85; vulkan/render_vulkan.cpp:1878-1917;
 * librender/render_backend.cpp:59-96).  normalize: options that do not apply to this backend's only integrator are reset to
 * their defaults (for the reference's megakernel program every RenderBackendOptions member applies, so this is the identity
 * but for range clamping of the enums).  configure_for: 0 = the backend now renders with these options (rng_variant is
 * switched; its tables must have been handed over); non-zero = unsupported, with the reason in last_error, and
 * *available (optional) receives the closest supported set -- the reference's "fallback exists" recovery path (app.cpp:400-431).
 * Unsupported here: light_sampling_variant NONE (the reference's megakernel does not build with it either: wpdf_direct_light
 * is undefined without lights_linear.glsl, rendering/mc/shade_base_material.glsl:36), enable_taa without option "realtime_resolve".
 * render_upscale_factor is taken over for the next initialize. */
int rptr_cuda_normalize_options(rptr_ctx *ctx, rptr_backend_options *options, int32_t variant);
/* ProcessTAAVulkan::process (vulkan/processing/process_taa.cpp:93-136, process_taa.comp): the LDR post-process the application
 * runs after end_frame when options.enable_taa && params.reprojection_mode != NONE (app.cpp:517-520): blends the frame's LDR
 * target with the previous frame's (Lanczos-5 resampled along the motion image, weight 0.15) and clamps to the 3 x 3
 * neighbourhood's spread; skipped while frame_id <= 1.  The next readback_u8 returns the processed target.  Needs option
 * "realtime_resolve".  The shader updates the target in place while neighbouring invocations still read it; here every read
 * sees the target as process_samples left it. */
int rptr_cuda_process_taa(rptr_ctx *ctx);
int rptr_cuda_configure_for(rptr_ctx *ctx, const rptr_backend_options *options, int32_t variant, rptr_backend_options *available);

/* Multi-GPU (the reference has no multi-GPU path to mirror): one context per GPU, scene replicated, the frame
 * sharded into interleaved bands of "tile_rows" rows (band b belongs to rank b % world; every rank seeds its pixels with the
 * GLOBAL pixel id, so the image does not depend on the number of GPUs), and ONE collective per readback: a sum of the RGBA32F
 * accumulators over NCCL -- ranks hold exact zeros outside their bands, so the sum is a gather, exact in fp32.
 * NCCL is loaded at run time (dlopen "libnccl.so.2": the copy the process already has, e.g. PyTorch's, else the system's;
 * RPTR_NCCL_LIB overrides), so single-GPU users need none.
 *   one process per GPU:  rank 0 calls rptr_cuda_comm_unique_id and ships the 128 bytes to the others (MPI, torch.distributed,
 *                         a file ...); every rank calls rptr_cuda_comm_init_rank; per readback every rank calls
 *                         rptr_cuda_reduce_framebuffer(root) and the root (every rank for root < 0 = all-reduce) reads the whole
 *                         image with rptr_cuda_readback_f32.
 *   one process, n GPUs:  (the reference's single render thread) rptr_cuda_comm_init_all over the n contexts, frames issued
 *                         context by context (all calls are asynchronous), rptr_cuda_reduce_framebuffer_all per readback.
 * comm_init sets the options tile_world / tile_rank; the reduce is enqueued on the context's stream (no host sync) and stays
 * valid until the next draw_frame. */
int rptr_cuda_comm_unique_id(void *id, size_t bytes);
int rptr_cuda_comm_init_rank(rptr_ctx *ctx, int32_t world, int32_t rank, const void *id, size_t bytes);
int rptr_cuda_comm_init_all(rptr_ctx **ctxs, int32_t n);
int rptr_cuda_comm_destroy(rptr_ctx *ctx);
int rptr_cuda_reduce_framebuffer(rptr_ctx *ctx, int32_t root);
int rptr_cuda_reduce_framebuffer_all(rptr_ctx **ctxs, int32_t n, int32_t root);

/* util/write_image.cpp:34-66 (WriteImage::write_pfm): "<prefix>.pfm", RGB, bottom row first, little-endian.
 * Pure host helper so validation mode (libapp/app_state.cpp:362-388) can be replayed without the app. */
int rptr_write_pfm(const char *prefix, uint32_t width, uint32_t height, uint32_t channels, const float *pixels);

#ifdef __cplusplus
}
#endif
#endif /* RPTR_CUDA_H */
