"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and for oracle/_ref/libref.so (the reference's own
sources compiled from /root/reference).  TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from realtimepathtracingresearchframework_b200 import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
f32p = C.POINTER(C.c_float)


class OracleRenderArgs(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("camera", T.RenderCameraParams), ("params", T.RenderParams),
                ("lighting", T.LightSamplingConfig), ("scene_params", T.SceneParams), ("frame_offset", C.c_uint32),
                ("first_sample", C.c_uint32), ("n_samples", C.c_int32), ("x0", C.c_int32), ("y0", C.c_int32),
                ("x1", C.c_int32), ("y1", C.c_int32), ("transmission", C.c_int32), ("n_threads", C.c_int32),
                ("rng_variant", C.c_int32), ("batch_spp", C.c_int32), ("pointset_tables", C.c_void_p * 4),
                ("vp_reference", C.c_float * 16)]


def build(force=False):
    """make -C oracle (liboracle.so, and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.exists("/root/reference/rendering"):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)
    return so


def _fp(a):
    return a.ctypes.data_as(f32p)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
        L.oracle_scene_destroy.argtypes = [C.c_void_p]
        L.oracle_scene_num_tris.restype = C.c_int64
        L.oracle_scene_num_tris.argtypes = [C.c_void_p]
        L.oracle_scene_num_lights.argtypes = [C.c_void_p]
        L.oracle_scene_get_lights.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_view_params.argtypes = [C.POINTER(T.RenderCameraParams), C.c_int32, C.c_int32, f32p]
        L.oracle_render.argtypes = [C.c_void_p, C.POINTER(OracleRenderArgs), f32p, C.POINTER(C.c_uint64)]
        L.oracle_render_sample.argtypes = [C.c_void_p, C.POINTER(OracleRenderArgs), C.c_uint32, f32p]
        L.oracle_render_aov.argtypes = [C.c_void_p, C.POINTER(OracleRenderArgs), C.c_uint32, f32p, f32p]
        L.oracle_render_aov3.argtypes = [C.c_void_p, C.POINTER(OracleRenderArgs), C.c_uint32, f32p, f32p, f32p]
        L.oracle_view_projection.argtypes = [C.POINTER(T.RenderCameraParams), C.c_int32, C.c_int32, f32p]
        L.oracle_render_ray_queries.argtypes = [C.c_void_p, C.POINTER(OracleRenderArgs), C.c_void_p, C.c_int32, f32p]
        for n in ("oracle_trace_closest", "oracle_trace_closest_bruteforce"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_int32, f32p, f32p]
        L.oracle_pointset_replay.argtypes = [C.c_int, C.POINTER(C.c_void_p)] + [C.c_uint32] * 6 + [C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                                                                       C.c_int, f32p, C.POINTER(C.c_uint32)]
        L.oracle_screen_jitter.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, f32p]
        L.oracle_halton_23.argtypes = [C.c_int32, f32p]
        L.oracle_morton_sample_id.restype = C.c_uint32
        L.oracle_morton_sample_id.argtypes = [C.c_uint32] * 5 + [C.c_int, C.c_int]
        L.oracle_lcg_seed.restype = C.c_uint32
        L.oracle_lcg_seed.argtypes = [C.c_uint32] * 3
        L.oracle_lcg_randomf.restype = C.c_float
        L.oracle_lcg_randomf.argtypes = [C.POINTER(C.c_uint32)]
        L.oracle_sincos.argtypes = [C.c_float, f32p, f32p]
        for n in ("oracle_exp", "oracle_acos", "oracle_fast_positive_atan"):
            getattr(L, n).restype = C.c_float
            getattr(L, n).argtypes = [C.c_float]
        L.oracle_gltf_wpdf.restype = C.c_float
        L.oracle_hit_attributes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, f32p]
        L.oracle_sky_illum.argtypes = [C.POINTER(T.SceneParams), f32p, C.c_float, f32p]
        L.oracle_shade_base_material.argtypes = [C.POINTER(T.BaseMaterial), C.c_int, C.c_int, C.c_float, f32p, f32p, C.c_float, f32p, f32p, C.c_uint32, C.c_int, C.c_int, f32p, C.c_float, f32p, C.c_void_p, C.c_int, C.c_int, f32p]
        L.oracle_sample_direct_light.argtypes = [C.POINTER(T.BaseMaterial)] + [f32p] * 8 + [C.c_float, f32p, C.c_void_p, C.c_int, C.c_int, f32p]
        L.oracle_unpack_material.argtypes = [C.POINTER(T.BaseMaterial), C.POINTER(T.TextureDesc), C.c_int, C.c_int, f32p]
        L.oracle_skymodel_radiance.argtypes = [C.POINTER(T.SceneParams), f32p, f32p, f32p]
        L.oracle_compute_sky_illum.argtypes = [C.POINTER(T.SceneParams), f32p, C.c_float, f32p]
        L.oracle_camera_ray.argtypes = [C.c_void_p, C.POINTER(OracleRenderArgs), C.c_int32, C.c_int32, C.c_uint32, f32p]
        L.oracle_geometry_scale_to_tmin.restype = C.c_float
        L.oracle_geometry_scale_to_tmin.argtypes = [f32p, C.c_float]
        L.oracle_running_mean.argtypes = [f32p, f32p, C.c_uint32, C.c_uint32]
        L.oracle_shadow_ray_range.restype = C.c_int32
        L.oracle_shadow_ray_range.argtypes = [f32p, C.c_float, C.c_float, f32p]
        L.oracle_shadow_alpha_seed.restype = C.c_uint32
        L.oracle_shadow_alpha_seed.argtypes = [C.c_uint32] * 5
        L.oracle_alpha_filter.restype = C.c_int32
        L.oracle_alpha_filter.argtypes = [C.c_float, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oracle_ray_query_tmin.restype = C.c_float
        L.oracle_ray_query_tmin.argtypes = [f32p]
        L.oracle_pack_ray_result.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_int32, f32p]
        L.oracle_decode_texel.restype = C.c_float
        L.oracle_decode_texel.argtypes = [C.c_int32, C.c_int32]
        L.oracle_bounce_prologue.argtypes = [f32p, C.c_uint32, C.c_int32, C.c_void_p, C.c_float, f32p]
        L.oracle_russian_roulette.restype = C.c_int32
        L.oracle_russian_roulette.argtypes = [C.c_int32, C.c_int32, f32p, C.c_float]
        L.oracle_sample_sun_dir.argtypes = [f32p, C.c_float, f32p, f32p]
        L.oracle_dequantize_position.argtypes = [C.c_uint64, f32p, f32p, f32p]
        L.oracle_dequantize_normal.argtypes = [C.c_uint32, f32p]
        L.oracle_dequantize_uv.argtypes = [C.c_uint32, f32p]
        L.oracle_reproject_accumulate.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
        L.oracle_process_taa.argtypes = [C.c_int32] * 5 + [C.c_void_p] * 4
        _lib = L
    return _lib


def reproject_accumulate(accum_cur, history, nd_history, nd, mj, min_sample_weight, batch, fn=None):
    """process_samples.comp:106-113 over a frame (reprojection.glsl).  accum_cur / history: (h, w, 4) float32; nd_history / nd / mj:
    (h, w, 4) float16 or uint16 bit patterns.  Returns (stored accumulator, colour shown).  fn: the same entry point of another
    library (tests/hostsim) instead of the oracle's."""
    h, w = accum_cur.shape[:2]
    arrs = [np.ascontiguousarray(accum_cur, np.float32), np.ascontiguousarray(history, np.float32)] + \
           [np.ascontiguousarray(a).view(np.uint16) for a in (nd_history, nd, mj)]
    stored, shown = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    (fn or lib().oracle_reproject_accumulate)(w, h, *[a.ctypes.data for a in arrs], float(min_sample_weight), int(batch), stored.ctypes.data,
                                              shown.ctypes.data)
    return stored, shown


def process_taa(current, history, mj, upscale=1, fn=None):
    """process_taa.comp over the LDR target: current / history (h, w, 4) uint8, mj the (h / upscale, w / upscale, 4) motion image."""
    h, w = current.shape[:2]
    rh, rw = mj.shape[:2]
    cur, his, m = np.ascontiguousarray(current, np.uint8), np.ascontiguousarray(history, np.uint8), np.ascontiguousarray(mj).view(np.uint16)
    out = np.zeros((h, w, 4), np.uint8)
    (fn or lib().oracle_process_taa)(w, h, int(upscale), rw, rh, cur.ctypes.data, his.ctypes.data, m.ctypes.data, out.ctypes.data)
    return out


def ref():
    """oracle/_ref/libref.so or None when it has not been built (it needs /root/reference at build time)."""
    global _ref
    if _ref is None:
        so = os.path.join(_HERE, "_ref", "libref.so")
        if not os.path.exists(so):
            return None
        R = C.CDLL(so)
        R.ref_lcg_seed.restype = C.c_uint32
        R.ref_lcg_seed.argtypes = [C.c_uint32] * 3
        R.ref_lcg_randomf.restype = C.c_float
        R.ref_lcg_randomf.argtypes = [C.POINTER(C.c_uint32)]
        R.ref_fast_positive_atan.restype = C.c_float
        R.ref_fast_positive_atan.argtypes = [C.c_float]
        R.ref_gltf_wpdf.restype = C.c_float
        R.ref_quantize_position.restype = C.c_uint64
        R.ref_quantize_normal.restype = C.c_uint32
        R.ref_quantize_uv.restype = C.c_uint32
        R.ref_dequantize_position.argtypes = [C.c_uint64, f32p, f32p, f32p]
        R.ref_dequantize_normal.argtypes = [C.c_uint32, f32p]
        R.ref_dequantize_uv.argtypes = [C.c_uint32, f32p]
        R.ref_sky_fit.argtypes = [C.POINTER(T.SceneConfig), C.POINTER(T.SceneParams)]
        R.ref_shade_base_material.argtypes = [C.POINTER(T.BaseMaterial), C.c_int, C.c_int, C.c_float, f32p, f32p, C.c_float, f32p, f32p, C.c_uint32, C.c_int, C.c_int, f32p, C.c_float, f32p, C.c_void_p, C.c_int, C.c_int, f32p]
        R.ref_sample_direct_light.argtypes = [C.POINTER(T.BaseMaterial)] + [f32p] * 8 + [C.c_float, f32p, C.c_void_p, C.c_int, C.c_int, f32p]
        R.ref_unpack_material.argtypes = [C.POINTER(T.BaseMaterial), f32p, C.c_int, C.c_int, f32p]
        R.ref_skymodel_radiance.argtypes = [C.POINTER(T.SceneParams), f32p, f32p, f32p]
        R.ref_compute_sky_illum.argtypes = [C.POINTER(T.SceneParams), f32p, f32p, C.c_float, f32p]
        R.ref_camera_ray.argtypes = [f32p] + [C.c_uint32] * 6 + [C.c_int32, f32p, f32p]
        R.ref_geometry_scale_to_tmin.restype = C.c_float
        R.ref_geometry_scale_to_tmin.argtypes = [f32p, C.c_float]
        R.ref_running_mean.argtypes = [f32p, f32p, C.c_uint32, C.c_uint32]
        R.ref_test_visibility.argtypes = [f32p, f32p, C.c_float, C.c_float] + [C.c_uint32] * 6 + [f32p, C.c_int32, C.c_int32, f32p]
        R.ref_alpha_filter.restype = C.c_int32
        R.ref_alpha_filter.argtypes = [C.c_float, C.c_uint32, C.POINTER(C.c_uint32)]
        R.ref_screen_jitter.argtypes = [C.c_uint32] * 4 + [f32p]
        R.ref_tonemap_srgb.argtypes = [C.c_int32, f32p, f32p]
        R.ref_ray_query.argtypes = [f32p, C.c_int32, f32p, C.c_int32, C.c_int32, C.c_int32, f32p, f32p]
        R.ref_srgb_to_linear.restype = C.c_float
        R.ref_srgb_to_linear.argtypes = [C.c_float]
        R.ref_view_params.argtypes = [C.POINTER(T.RenderCameraParams), C.c_uint32, C.c_uint32, f32p]
        R.ref_bounce_prologue.argtypes = [f32p, C.c_uint32, C.c_int32, C.c_float, f32p]
        R.ref_russian_roulette.restype = C.c_int32
        R.ref_russian_roulette.argtypes = [C.c_int32, C.c_int32, f32p, C.c_float]
        R.ref_sample_sun_dir.argtypes = [f32p, C.c_float, f32p, f32p]
        R.ref_pointset_table.argtypes = [C.c_int, C.POINTER(C.c_uint32)]
        R.ref_pointset_replay.argtypes = [C.c_int] + [C.c_uint32] * 7 + [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, f32p,
                                                                        C.POINTER(C.c_uint32)]
        R.ref_halton_23.argtypes = [f32p]
        R.ref_morton_sample_id.restype = C.c_uint32
        R.ref_morton_sample_id.argtypes = [C.c_uint32] * 5 + [C.c_int, C.c_int]
        _ref = R
    return _ref


class OracleScene:
    def __init__(self, scene, lighting=None):
        self.scene = scene
        self.lighting = lighting or T.LightSamplingConfig()
        d = scene.desc()
        self.h = lib().oracle_scene_create(C.byref(d), C.byref(self.lighting))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_scene_destroy(self.h)
            self.h = None

    @property
    def num_tris(self):
        return lib().oracle_scene_num_tris(self.h)

    def lights(self):
        n = lib().oracle_scene_num_lights(self.h)
        arr = (T.TriLightData * max(n, 1))()
        lib().oracle_scene_get_lights(self.h, arr)
        return np.frombuffer(arr, dtype=np.float32).reshape(-1, 12)[:n].copy()

    def _args(self, width, height, camera, scene_params, params=None, frame_offset=0, first_sample=0, n_samples=1,
              region=None, transmission=0, n_threads=0, rng_variant=0, batch_spp=1, pointset_tables=None, vp_reference=None):
        a = OracleRenderArgs()
        a.width, a.height = width, height
        a.camera = camera
        a.params = params or T.RenderParams()
        a.lighting = self.lighting
        a.scene_params = scene_params
        a.frame_offset, a.first_sample, a.n_samples = frame_offset, first_sample, n_samples
        x0, y0, x1, y1 = region or (0, 0, width, height)
        a.x0, a.y0, a.x1, a.y1 = x0, y0, x1, y1
        a.transmission, a.n_threads = transmission, n_threads
        a.rng_variant, a.batch_spp = rng_variant, batch_spp
        if vp_reference is not None:  # view_params.VP_reference (16 floats, column-major); zero = before the first frame
            a.vp_reference[:] = [float(v) for v in np.asarray(vp_reference, np.float32).reshape(16)]
        if rng_variant != 0:
            if pointset_tables is None:
                raise ValueError("rng_variant != 0 needs pointset_tables (four uint32 arrays)")
            self._tables = [np.ascontiguousarray(t, np.uint32) for t in pointset_tables]  # keep alive during the call
            for i, t in enumerate(self._tables):
                a.pointset_tables[i] = t.ctypes.data
        return a

    def render(self, width, height, camera, scene_params, spp, out=None, **kw):
        """Running mean over `spp` frames of batch_spp=1 -> (H, W, 4) float32, plus (closest, shadow, vertices) counters."""
        a = self._args(width, height, camera, scene_params, n_samples=spp, **kw)
        img = np.zeros((height, width, 4), np.float32) if out is None else out
        stats = (C.c_uint64 * 3)()
        lib().oracle_render(self.h, C.byref(a), _fp(img), stats)
        return img, tuple(int(x) for x in stats)

    def render_sample(self, width, height, camera, scene_params, sample_index, **kw):
        a = self._args(width, height, camera, scene_params, **kw)
        img = np.zeros((height, width, 4), np.float32)
        lib().oracle_render_sample(self.h, C.byref(a), sample_index, _fp(img))
        return img

    def render_aov(self, width, height, camera, scene_params, sample_index, **kw):
        """Float values behind the fp16 AOV images (albedo+roughness, normal+depth) for one sample layer."""
        a = self._args(width, height, camera, scene_params, **kw)
        ar = np.zeros((height, width, 4), np.float32)
        nd = np.zeros((height, width, 4), np.float32)
        lib().oracle_render_aov(self.h, C.byref(a), sample_index, _fp(ar), _fp(nd))
        return ar, nd

    def render_aov3(self, width, height, camera, scene_params, sample_index, **kw):
        """As render_aov plus the motion / jitter image; first_sample (kw) = frame_id of the frame the layer belongs to,
        vp_reference (kw) = VP of the previous frame (view_projection of its camera)."""
        a = self._args(width, height, camera, scene_params, **kw)
        ar = np.zeros((height, width, 4), np.float32)
        nd = np.zeros((height, width, 4), np.float32)
        mj = np.zeros((height, width, 4), np.float32)
        lib().oracle_render_aov3(self.h, C.byref(a), sample_index, _fp(ar), _fp(nd), _fp(mj))
        return ar, nd, mj

    def render_ray_queries(self, width, height, camera, scene_params, queries, view_frame_id=0, **kw):
        """RenderBackend::render_ray_queries: (n, 8) RenderRayQuery rows -> (n, 4) results; kw as for render (batch_spp, params,
        frame_offset, rng_variant, ...).  view_frame_id = view_params.frame_id of the last begin_frame."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 8)
        a = self._args(width, height, camera, scene_params, first_sample=view_frame_id, **kw)
        res = np.zeros((q.shape[0], 4), np.float32)
        lib().oracle_render_ray_queries(self.h, C.byref(a), q.ctypes.data, q.shape[0], _fp(res))
        return res

    def trace_closest(self, queries, bruteforce=False):
        """queries: structured (n, 8) float32 view of RenderRayQuery -> (results (n,4) float32 bits, t (n,))."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 8)
        n = q.shape[0]
        res = np.zeros((n, 4), np.float32)
        t = np.zeros(n, np.float32)
        fn = lib().oracle_trace_closest_bruteforce if bruteforce else lib().oracle_trace_closest
        fn(self.h, q.ctypes.data, n, _fp(res), _fp(t))
        return res, t


def table_ptrs(tables):
    """(C array of 4 pointers, keep-alive list) for oracle_pointset_replay."""
    keep = [np.ascontiguousarray(t, np.uint32) for t in tables]
    arr = (C.c_void_p * 4)(*[t.ctypes.data for t in keep])
    return arr, keep


def view_params(camera, w, h):
    out = np.zeros(9, np.float32)
    lib().oracle_view_params(C.byref(camera), w, h, _fp(out))
    return out


def view_projection(camera, w, h):
    """view_params.VP (16 floats, column-major) for a camera and frame size."""
    out = np.zeros(16, np.float32)
    lib().oracle_view_projection(C.byref(camera), w, h, _fp(out))
    return out


def sky_fit(config=None):
    """update_sky_light through the reference's own sky_model.cpp (needs oracle/_ref)."""
    R = ref()
    if R is None:
        raise RuntimeError("oracle/_ref/libref.so is not built (needs /root/reference at build time)")
    cfg = config or T.SceneConfig()
    sp = T.SceneParams()
    R.ref_sky_fit(C.byref(cfg), C.byref(sp))
    return sp
