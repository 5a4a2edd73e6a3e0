/* rptr_types.h -- plain-C mirror of the POD types that cross the reference's backend boundary.
 *
 * Every struct here restates (layout-compatible, field for field) a type the reference shares between its
 * C++ host and its shaders; the adapter in INTEGRATION.md static_asserts sizeof/offsetof against the originals.
 * Reference anchors (relative to the reference tree):
 *   rptr_base_material        <- rendering/bsdfs/base_material.h.glsl:13-34   (80 B)
 *   rptr_render_params        <- librender/render_params.glsl.h:130-155       (80 B)
 *   rptr_light_sampling_config<- librender/render_params.glsl.h:123-128       (16 B)
 *   rptr_scene_config         <- librender/render_params.glsl.h:157-162       (32 B)
 *   rptr_render_ray_query     <- librender/render_params.glsl.h:165-170       (32 B)
 *   rptr_tri_light_data       <- rendering/lights/tri.h.glsl:13-26            (48 B)
 *   rptr_camera_params        <- librender/render_backend.h:26-31             (40 B)
 *   rptr_render_stats         <- librender/render_backend.h:15-24
 *   rptr_backend_options      <- librender/render_params.glsl.h:73-119        (32 B)
 *   rptr_scene_params         <- vulkan/gpu_params.glsl:113-131 (SceneParams, the fitted sky/sun block)
 *   rptr_geometry/mesh/pmesh/instance <- librender/mesh.h:10-41,78-121, librender/scene.h:48-72
 */
#ifndef RPTR_TYPES_H
#define RPTR_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rendering/bsdfs/base_material.h.glsl:7-11 */
#define RPTR_BASE_MATERIAL_NOALPHA 0x01u
#define RPTR_BASE_MATERIAL_ONESIDED 0x02u
#define RPTR_BASE_MATERIAL_VOLUME 0x04u
#define RPTR_BASE_MATERIAL_EXTENDED 0x08u
#define RPTR_BASE_MATERIAL_NEURAL 0x10u

/* rendering/rt/geometry.h.glsl:66-70 */
#define RPTR_GEOMETRY_FLAGS_NOALPHA 0x01u
#define RPTR_GEOMETRY_FLAGS_IMPLICIT_INDICES 0x02u
#define RPTR_GEOMETRY_FLAGS_EXTENDED_SHADER 0x04u
#define RPTR_GEOMETRY_FLAGS_THIN 0x08u
#define RPTR_GEOMETRY_FLAGS_DYNAMIC 0x10u

/* rendering/postprocess/reprojection.h:11-13 (RenderParams::reprojection_mode); without ENABLE_REALTIME_RESOLVE only
 * DISCARD_HISTORY changes the resolve: the frame's samples replace the history instead of being folded into it */
#define RPTR_REPROJECTION_MODE_NONE 0
#define RPTR_REPROJECTION_MODE_DISCARD_HISTORY 1
#define RPTR_REPROJECTION_MODE_ACCUMULATE 2

/* vulkan/gpu_params.glsl:27-29 */
#define RPTR_RAY_EPSILON 0.000005f

/* librender/render_params.glsl.h:16-19 */
#define RPTR_MAX_PATH_DEPTH 9
#define RPTR_DEFAULT_RR_PATH_DEPTH 2
#define RPTR_BINNED_LIGHTS_BIN_MAX_SIZE 16
#define RPTR_GLOSSY_MODE_ROUGHNESS_THRESHOLD 0.1f

/* rendering/mc/light_sampling.h:11-12 */
#define RPTR_LIGHT_SAMPLING_VARIANT_NONE 0
#define RPTR_LIGHT_SAMPLING_VARIANT_RIS 1

/* librender/render_params.glsl.h:34-37 */
#define RPTR_RNG_VARIANT_UNIFORM 0
#define RPTR_RNG_VARIANT_BN 1
#define RPTR_RNG_VARIANT_SOBOL 2
#define RPTR_RNG_VARIANT_Z_SBL 3

/* tables of the low-discrepancy samplers, as the reference holds them (rendering/pointsets/sobol_data.h:13-17,
 * bn_data.h:12-27; sources rendering/pointsets/sobol_tables.h, bn_tables.h) */
#define RPTR_POINTSET_SOBOL_MATRIX 0       /* SobolMatrix            uint32[1024 * 32] */
#define RPTR_POINTSET_SOBOL_TILE_INVERT 1  /* SobolInversion_1_0     uint32[256 * 256] */
#define RPTR_POINTSET_BN_SOBOL 2           /* sobol_256spp_256d      uint32[256 * 256] */
#define RPTR_POINTSET_BN_SCRAMBLING_1SPP 3 /* scramblingTile_yx_d_1spp uint32[128 * 128 * 8] */

typedef struct rptr_base_material {
    float base_color[3];
    int32_t normal_map; /* -1 = none */
    uint32_t flags;
    float roughness;
    float specular;
    float metallic;
    float sheen;
    float sheen_tint;
    float clearcoat;
    float clearcoat_gloss;
    float ior;
    float specular_transmission;
    float anisotropy;
    float specular_tint;
    float transmission_color[3];
    float emission_intensity;
} rptr_base_material;

typedef struct rptr_render_params {
    int32_t batch_spp;
    int32_t max_path_depth;
    int32_t rr_path_depth;
    int32_t glossy_only_mode;
    float aperture_radius;
    float focus_distance;
    float pixel_radius;
    float variance_radius;
    int32_t output_channel;
    int32_t output_moment;
    float exposure;
    int32_t early_tone_mapping_mode;
    int32_t reprojection_mode;
    int32_t spp_accumulation_window;
    int32_t enable_raster_taa;
    int32_t render_upscale_factor;
    float focal_length;
    int32_t _pad3, _pad4, _pad5;
} rptr_render_params;

typedef struct rptr_light_sampling_config {
    float light_mis_angle;
    int32_t bin_size;
    float min_perceived_receiver_dist;
    float min_radiance;
} rptr_light_sampling_config;

typedef struct rptr_scene_config {
    float bump_scale;
    float sun_dir[3];
    float turbidity;
    float albedo[3];
} rptr_scene_config;

typedef struct rptr_render_ray_query {
    float origin[3];
    int32_t mode_or_data;
    float dir[3];
    float t_max;
} rptr_render_ray_query;

typedef struct rptr_tri_light_data {
    float v0[3];
    float v1[3];
    float v2[3];
    float radiance[3];
} rptr_tri_light_data;

typedef struct rptr_camera_params {
    float pos[3];
    float dir[3];
    float up[3];
    float fovy; /* degrees */
} rptr_camera_params;

typedef struct rptr_render_stats {
    float render_time; /* ms of device time for the last draw_frame..end_frame */
    float rays_per_second;
    int32_t spp;
    int16_t frame_stats_delay;
    uint8_t has_valid_frame_stats;
    uint8_t _pad;
    uint64_t total_device_bytes_allocated;
    uint64_t max_device_bytes_allocated;
    uint64_t device_bytes_currently_allocated;
} rptr_render_stats;

/* RenderBackendOptions (librender/render_params.glsl.h:73-119), member for member; bool members are one byte like the C++ type */
typedef struct rptr_backend_options {
    int32_t rng_variant;                 /* RPTR_RNG_VARIANT_* */
    int32_t light_sampling_variant;      /* RPTR_LIGHT_SAMPLING_VARIANT_* */
    int32_t light_sampling_bucket_count; /* 16 */
    uint8_t unroll_bounces;
    uint8_t _pad0[3];
    int32_t render_upscale_factor;       /* 1 */
    uint8_t enable_rayqueries;
    uint8_t force_bvh_rebuild;
    uint8_t _pad1[2];
    int32_t rebuild_triangle_budget;     /* 500000 */
    uint8_t enable_taa;
    uint8_t enable_raytraced_dof;        /* true */
    uint8_t _pad2[2];
} rptr_backend_options;

/* The fitted sky/sun block the kernels read; produced on the host by update_sky_light
 * (vulkan/render_sky.cpp:25-72) from an rptr_scene_config. */
typedef struct rptr_scene_params {
    float sky_configs[9][4]; /* SkyModelParams.configs (rgb + pad) */
    float sky_radiances[4];
    float sun_dir[3];
    float sun_cos_angle;
    float sun_radiance[4]; /* rgb + w = probability of picking the sun in NEE */
    float normal_z_scale;
    int32_t _pad[3];
} rptr_scene_params;

/* One triangle soup of a Mesh (librender/mesh.h:10-41). Vertices are unrolled (3 per triangle,
 * REQUIRE_UNROLLED_VERTICES, vulkan/gpu_params.glsl:9) and quantised to 21 bits per axis
 * (librender/dequantize.glsl:8-21). qnormal_uv packs an oct-encoded normal (low 32 bits) and a
 * quantised uv (high 32 bits) per vertex (librender/dequantize.glsl:23-48); NULL = neither. */
typedef struct rptr_geometry_desc {
    const uint64_t *qverts;
    const uint64_t *qnormal_uv;
    float quantized_scaling[3];
    float quantized_offset[3];
    int32_t n_tris;
    int32_t has_normals;
    int32_t has_uvs;
    int32_t _pad;
} rptr_geometry_desc;

typedef struct rptr_mesh_desc {
    int32_t first_geometry; /* index into rptr_scene_desc.geometries */
    int32_t n_geometries;
} rptr_mesh_desc;

/* librender/mesh.h:78-121 */
typedef struct rptr_pmesh_desc {
    int32_t mesh_id;
    int32_t n_material_offsets;        /* 0 -> offset 0 for every geometry */
    const int32_t *material_offsets;   /* per geometry of the mesh */
    const uint8_t *tri_material_ids;   /* per triangle over the whole mesh, or NULL */
    int64_t n_tri_material_ids;
} rptr_pmesh_desc;

/* librender/mesh.h:125-129 after AnimationData::dequantize (librender/scene.cpp:22-41):
 * object-to-world, 3 rows x 4 columns, row-major (the layout handed to the TLAS,
 * vulkan/render_vulkan.cpp:1262-1268). */
typedef struct rptr_instance_desc {
    int32_t pmesh_id;
    float transform[12];
} rptr_instance_desc;

/* A texture as the reference's Image holds it (util/image.h:8-27): 8-bit texels, `channels` per pixel, row-major.
 * This round the backend accepts 1 x 1 textures only ("1x1-texel mode"): a material parameter that carries a
 * texture handle (rendering/bsdfs/texture_channel_mask.h) then resolves to one texel, UNORM8 -> value / 255, colour
 * channels of an sRGB texture through the sRGB transfer function.  Larger textures are rejected with an error. */
#define RPTR_COLOR_SPACE_LINEAR 0
#define RPTR_COLOR_SPACE_SRGB 1
/* One image of Scene::textures (util/image.h:10-27).  `texels` holds mip_levels levels back to back, the base level first,
 * level k + 1 being (max(w / 2, 1), max(h / 2, 1)) of level k (Image::mip_levels, util/image.cpp:23-40).  bc_format is
 * Image::bcFormat: 0 = `channels` bytes per texel; 1 = BC1 RGB, -1 = BC1 RGBA (1-bit alpha), 3 = BC3, 5 = BC5 UNORM (two
 * channels: the normal maps of .vks scenes) -- the formats librender/scene.cpp:836-930 produces; 4 x 4 blocks, every level
 * padded to whole blocks (vulkan/resource_utils.cpp:85-99).  Block-compressed levels are decoded to RGBA8 inside set_scene. */
typedef struct rptr_texture_desc {
    int32_t width, height;
    int32_t channels;    /* 1..4; missing colour channels read 0, missing alpha reads 255 (bc_format 0 only) */
    int32_t color_space; /* RPTR_COLOR_SPACE_* */
    const uint8_t *texels;
    int32_t bc_format;   /* 0, 1, -1, 3, 5 */
    int32_t mip_levels;  /* levels stored in texels; 0 reads as 1 */
} rptr_texture_desc;

/* rendering/bsdfs/texture_channel_mask.h:20-27: a float material parameter with the sign bit set is a texture handle */
#define RPTR_TEXTURED_PARAM_MASK 0x80000000u
#define RPTR_GET_TEXTURE_CHANNEL(x) (((x) >> 29) & 0x3u)
#define RPTR_GET_TEXTURE_ID(x) ((x) & 0x1fffffffu)

typedef struct rptr_scene_desc {
    const rptr_geometry_desc *geometries;
    int32_t n_geometries;
    const rptr_mesh_desc *meshes;
    int32_t n_meshes;
    const rptr_pmesh_desc *pmeshes;
    int32_t n_pmeshes;
    const rptr_instance_desc *instances;
    int32_t n_instances;
    const rptr_base_material *materials;
    int32_t n_materials;
    /* Optional: emitters already collected+binned by the caller (TriLightData order is part of the
     * NEE contract). NULL -> the backend runs collect_emitters + update_light_sampling itself
     * (librender/lights.cpp:14-90). */
    const rptr_tri_light_data *binned_lights;
    int32_t n_binned_lights;
    /* Scene::textures (librender/scene.h); material parameters refer to them through texture handles. May be NULL. */
    const rptr_texture_desc *textures;
    int32_t n_textures;
} rptr_scene_desc;

#ifdef __cplusplus
}
#endif

#endif /* RPTR_TYPES_H */
