"""ctypes mirrors of include/rptr_types.h (the POD types crossing the reference's backend boundary).

Field order, names and defaults follow the reference headers cited in rptr_types.h
(rendering/bsdfs/base_material.h.glsl:13-34, librender/render_params.glsl.h:123-170, librender/render_backend.h:15-31).
"""
import ctypes as C

BASE_MATERIAL_NOALPHA = 0x01
BASE_MATERIAL_ONESIDED = 0x02
BASE_MATERIAL_VOLUME = 0x04
BASE_MATERIAL_EXTENDED = 0x08
GEOMETRY_FLAGS_NOALPHA = 0x01
RNG_VARIANT_UNIFORM, RNG_VARIANT_BN, RNG_VARIANT_SOBOL, RNG_VARIANT_Z_SBL = 0, 1, 2, 3  # librender/render_params.glsl.h:34-37
REPROJECTION_MODE_NONE, REPROJECTION_MODE_DISCARD_HISTORY, REPROJECTION_MODE_ACCUMULATE = 0, 1, 2  # rendering/postprocess/reprojection.h:11-13
MAX_PATH_DEPTH = 9
DEFAULT_RR_PATH_DEPTH = 2
DEFAULT_RAY_QUERY_BUDGET = 512 * 512  # librender/render_params.glsl.h:172
LIGHT_SAMPLING_VARIANT_NONE, LIGHT_SAMPLING_VARIANT_RIS = 0, 1  # rendering/mc/light_sampling.h:11-12

f32, i32, u32 = C.c_float, C.c_int32, C.c_uint32


class _Pod(C.Structure):
    _defaults_ = {}

    def __init__(self, **kw):
        super().__init__()
        for k, v in {**self._defaults_, **kw}.items():
            cur = getattr(self, k)
            if isinstance(cur, C.Array):
                for i, x in enumerate(v):
                    cur[i] = x
            else:
                setattr(self, k, v)


class BaseMaterial(_Pod):
    _fields_ = [("base_color", f32 * 3), ("normal_map", i32), ("flags", u32), ("roughness", f32), ("specular", f32),
                ("metallic", f32), ("sheen", f32), ("sheen_tint", f32), ("clearcoat", f32), ("clearcoat_gloss", f32),
                ("ior", f32), ("specular_transmission", f32), ("anisotropy", f32), ("specular_tint", f32),
                ("transmission_color", f32 * 3), ("emission_intensity", f32)]
    _defaults_ = dict(base_color=(0.9, 0.9, 0.9), normal_map=-1, roughness=1.0, specular=0.5, clearcoat_gloss=0.1,
                      ior=1.5, transmission_color=(1.0, 1.0, 1.0))


class RenderParams(_Pod):
    _fields_ = [("batch_spp", i32), ("max_path_depth", i32), ("rr_path_depth", i32), ("glossy_only_mode", i32),
                ("aperture_radius", f32), ("focus_distance", f32), ("pixel_radius", f32), ("variance_radius", f32),
                ("output_channel", i32), ("output_moment", i32), ("exposure", f32), ("early_tone_mapping_mode", i32),
                ("reprojection_mode", i32), ("spp_accumulation_window", i32), ("enable_raster_taa", i32),
                ("render_upscale_factor", i32), ("focal_length", f32), ("_pad3", i32), ("_pad4", i32), ("_pad5", i32)]
    _defaults_ = dict(batch_spp=1, max_path_depth=MAX_PATH_DEPTH, rr_path_depth=DEFAULT_RR_PATH_DEPTH, focus_distance=2.5,
                      pixel_radius=1.0, variance_radius=4.0, early_tone_mapping_mode=-1, spp_accumulation_window=8,
                      render_upscale_factor=1, focal_length=35.0)


class RenderBackendOptions(_Pod):
    """librender/render_params.glsl.h:73-119"""
    _fields_ = [("rng_variant", i32), ("light_sampling_variant", i32), ("light_sampling_bucket_count", i32), ("unroll_bounces", C.c_uint8),
                ("_pad0", C.c_uint8 * 3), ("render_upscale_factor", i32), ("enable_rayqueries", C.c_uint8), ("force_bvh_rebuild", C.c_uint8),
                ("_pad1", C.c_uint8 * 2), ("rebuild_triangle_budget", i32), ("enable_taa", C.c_uint8), ("enable_raytraced_dof", C.c_uint8),
                ("_pad2", C.c_uint8 * 2)]
    _defaults_ = dict(rng_variant=0, light_sampling_variant=1, light_sampling_bucket_count=16, render_upscale_factor=1,
                      rebuild_triangle_budget=500000, enable_raytraced_dof=1)


class LightSamplingConfig(_Pod):
    _fields_ = [("light_mis_angle", f32), ("bin_size", i32), ("min_perceived_receiver_dist", f32), ("min_radiance", f32)]
    _defaults_ = dict(bin_size=16, min_perceived_receiver_dist=15.0)


class SceneConfig(_Pod):
    _fields_ = [("bump_scale", f32), ("sun_dir", f32 * 3), ("turbidity", f32), ("albedo", f32 * 3)]
    _defaults_ = dict(bump_scale=1.0, sun_dir=(0.0, 1.0, 0.0), turbidity=3.0, albedo=(0.2, 0.2, 0.2))


class RenderRayQuery(_Pod):
    _fields_ = [("origin", f32 * 3), ("mode_or_data", i32), ("dir", f32 * 3), ("t_max", f32)]


class TriLightData(_Pod):
    _fields_ = [("v0", f32 * 3), ("v1", f32 * 3), ("v2", f32 * 3), ("radiance", f32 * 3)]


class RenderCameraParams(_Pod):
    _fields_ = [("pos", f32 * 3), ("dir", f32 * 3), ("up", f32 * 3), ("fovy", f32)]
    _defaults_ = dict(pos=(0.0, 2.0, 5.0), dir=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0), fovy=65.0)


class RenderStats(_Pod):
    _fields_ = [("render_time", f32), ("rays_per_second", f32), ("spp", i32), ("frame_stats_delay", C.c_int16),
                ("has_valid_frame_stats", C.c_uint8), ("_pad", C.c_uint8), ("total_device_bytes_allocated", C.c_uint64),
                ("max_device_bytes_allocated", C.c_uint64), ("device_bytes_currently_allocated", C.c_uint64)]


class SceneParams(_Pod):
    _fields_ = [("sky_configs", (f32 * 4) * 9), ("sky_radiances", f32 * 4), ("sun_dir", f32 * 3), ("sun_cos_angle", f32),
                ("sun_radiance", f32 * 4), ("normal_z_scale", f32), ("_pad", i32 * 3)]


class GeometryDesc(_Pod):
    _fields_ = [("qverts", C.c_void_p), ("qnormal_uv", C.c_void_p), ("quantized_scaling", f32 * 3),
                ("quantized_offset", f32 * 3), ("n_tris", i32), ("has_normals", i32), ("has_uvs", i32), ("_pad", i32)]


class MeshDesc(_Pod):
    _fields_ = [("first_geometry", i32), ("n_geometries", i32)]


class PMeshDesc(_Pod):
    _fields_ = [("mesh_id", i32), ("n_material_offsets", i32), ("material_offsets", C.c_void_p),
                ("tri_material_ids", C.c_void_p), ("n_tri_material_ids", C.c_int64)]


class InstanceDesc(_Pod):
    _fields_ = [("pmesh_id", i32), ("transform", f32 * 12)]
    _defaults_ = dict(transform=(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0))


class TextureDesc(_Pod):
    # texels: mip_levels levels back to back (0 reads as 1); bc_format = Image::bcFormat (0 raw, 1 / -1 BC1 RGB / RGBA, 3 BC3, 5 BC5)
    _fields_ = [("width", i32), ("height", i32), ("channels", i32), ("color_space", i32), ("texels", C.POINTER(C.c_uint8)),
                ("bc_format", i32), ("mip_levels", i32)]


COLOR_SPACE_LINEAR, COLOR_SPACE_SRGB = 0, 1


def texture_handle(texture_id, channel=0):
    """A float material parameter that refers to a texture (rendering/bsdfs/texture_channel_mask.h:20-27): sign bit set,
    bits 29-30 = channel, bits 0-28 = texture id."""
    import struct
    return struct.unpack("<f", struct.pack("<I", 0x80000000 | ((channel & 3) << 29) | (texture_id & 0x1fffffff)))[0]


class SceneDesc(_Pod):
    _fields_ = [("geometries", C.POINTER(GeometryDesc)), ("n_geometries", i32), ("meshes", C.POINTER(MeshDesc)),
                ("n_meshes", i32), ("pmeshes", C.POINTER(PMeshDesc)), ("n_pmeshes", i32),
                ("instances", C.POINTER(InstanceDesc)), ("n_instances", i32), ("materials", C.POINTER(BaseMaterial)),
                ("n_materials", i32), ("binned_lights", C.POINTER(TriLightData)), ("n_binned_lights", i32),
                ("textures", C.POINTER(TextureDesc)), ("n_textures", i32)]


assert C.sizeof(BaseMaterial) == 80 and C.sizeof(RenderParams) == 80 and C.sizeof(LightSamplingConfig) == 16
assert C.sizeof(SceneConfig) == 32 and C.sizeof(RenderRayQuery) == 32 and C.sizeof(TriLightData) == 48
assert C.sizeof(RenderCameraParams) == 40 and C.sizeof(SceneParams) == 208 and C.sizeof(RenderBackendOptions) == 32
