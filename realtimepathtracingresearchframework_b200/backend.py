"""Host-side mirror of the reference's plugin surface for the CUDA backend, over the C ABI of librptr_cuda.so.

`RenderCuda` follows `struct RenderBackend : RenderGraphic` (librender/render_backend.h:68-116,
util/display/render_graphic.h:11-44): same method names, argument meaning, counter protocol and error behaviour
(errors raise, like the reference's throw_error -> logged_exception, util/error_io.h:27-30; readback returns 0 elements
when the buffer is too small, vulkan/render_vulkan.cpp:2262-2263).  The C++ adapter a maintainer would add to the
reference tree is shown in INTEGRATION.md; this module is the same thing for Python callers (tests, bench.py).

There is no CPU fallback: constructing RenderCuda without the CUDA library or without a GPU raises.
"""
import ctypes as C
import json
import os

import numpy as np

from . import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librptr_cuda.so")

ABI_SYMBOLS = [
    "rptr_cuda_create", "rptr_cuda_destroy", "rptr_cuda_last_error", "rptr_cuda_name", "rptr_cuda_initialize",
    "rptr_cuda_set_scene", "rptr_cuda_get_lights", "rptr_cuda_set_scene_params", "rptr_cuda_set_option",
    "rptr_cuda_begin_frame", "rptr_cuda_draw_frame", "rptr_cuda_end_frame", "rptr_cuda_stats", "rptr_cuda_flush",
    "rptr_cuda_get_counters", "rptr_cuda_reset_counters", "rptr_cuda_frame_state", "rptr_cuda_framebuffer_size",
    "rptr_cuda_readback_f32", "rptr_cuda_readback_u8", "rptr_cuda_framebuffer_device_ptr", "rptr_cuda_stream_handle",
    "rptr_cuda_trace_rays", "rptr_cuda_set_pointset_table", "rptr_cuda_readback_aov",
    "rptr_cuda_enable_ray_queries", "rptr_cuda_ray_query_buffers", "rptr_cuda_write_ray_queries", "rptr_cuda_read_ray_results",
    "rptr_cuda_render_ray_queries", "rptr_cuda_normalize_options", "rptr_cuda_configure_for", "rptr_cuda_process_taa",
    "rptr_cuda_comm_unique_id", "rptr_cuda_comm_init_rank", "rptr_cuda_comm_init_all", "rptr_cuda_comm_destroy",
    "rptr_cuda_reduce_framebuffer", "rptr_cuda_reduce_framebuffer_all",
    "rptr_write_pfm",
]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "closest_rays", "shadow_rays", "shaded_vertices", "closest_nodes",
                                          "closest_tris", "shadow_nodes", "shadow_tris", "launches")] + \
               [(n, C.c_double) for n in ("ms_trace", "ms_shadow", "ms_shade", "ms_other")] + \
               [(n, C.c_uint64) for n in ("trace_launches", "node_bytes", "tri_bytes", "bvh_nodes")] + [("bvh_build_ms", C.c_double),
                                                                                                  ("trace_overlap", C.c_uint64), ("num_sms", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class RptrError(RuntimeError):
    """The reference's logged_exception (util/error_io.h:27-30)."""


_lib = None


def load_library(path=None):
    """dlopen librptr_cuda.so and declare the prototypes of include/rptr_cuda.h.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("RPTR_CUDA_LIB") or LIB_PATH  # RPTR_CUDA_LIB: tuning variants built by tools/sweep.py
    if not os.path.exists(p):
        raise RptrError("librptr_cuda.so is not built (%s); run `python -m realtimepathtracingresearchframework_b200.build`. "
                        "There is no CPU fallback." % p)
    L = C.CDLL(p)
    vp, i32, i64, u32p = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_uint32)
    L.rptr_cuda_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.rptr_cuda_destroy.argtypes = [vp]
    L.rptr_cuda_destroy.restype = None
    L.rptr_cuda_last_error.argtypes = [vp]
    L.rptr_cuda_last_error.restype = C.c_char_p
    L.rptr_cuda_name.restype = C.c_char_p
    L.rptr_cuda_initialize.argtypes = [vp, i32, i32]
    L.rptr_cuda_set_scene.argtypes = [vp, C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    L.rptr_cuda_get_lights.argtypes = [vp, vp, i32]
    L.rptr_cuda_get_lights.restype = i32
    L.rptr_cuda_set_scene_params.argtypes = [vp, C.POINTER(T.SceneParams)]
    L.rptr_cuda_set_option.argtypes = [vp, C.c_char_p, i64]
    L.rptr_cuda_set_pointset_table.argtypes = [vp, i32, u32p, C.c_size_t]
    L.rptr_cuda_begin_frame.argtypes = [vp, C.POINTER(T.RenderCameraParams), C.POINTER(T.RenderParams),
                                        C.POINTER(T.LightSamplingConfig), i32, i32, C.c_double]
    L.rptr_cuda_draw_frame.argtypes = [vp, i32]
    L.rptr_cuda_end_frame.argtypes = [vp, i32]
    L.rptr_cuda_stats.argtypes = [vp, C.POINTER(T.RenderStats)]
    L.rptr_cuda_flush.argtypes = [vp]
    L.rptr_cuda_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.rptr_cuda_reset_counters.argtypes = [vp]
    L.rptr_cuda_frame_state.argtypes = [vp, u32p, u32p, u32p]
    L.rptr_cuda_framebuffer_size.argtypes = [vp, u32p, u32p, u32p]
    L.rptr_cuda_readback_f32.argtypes = [vp, C.c_size_t, vp]
    L.rptr_cuda_readback_f32.restype = C.c_size_t
    L.rptr_cuda_readback_u8.argtypes = [vp, C.c_size_t, vp]
    L.rptr_cuda_readback_u8.restype = C.c_size_t
    L.rptr_cuda_readback_aov.argtypes = [vp, i32, C.c_size_t, vp]
    L.rptr_cuda_readback_aov.restype = C.c_size_t
    L.rptr_cuda_framebuffer_device_ptr.argtypes = [vp, C.POINTER(vp)]
    L.rptr_cuda_stream_handle.argtypes = [vp, C.POINTER(vp)]
    L.rptr_cuda_trace_rays.argtypes = [vp, vp, i32, vp, vp]
    L.rptr_cuda_enable_ray_queries.argtypes = [vp, i32, i32]
    L.rptr_cuda_ray_query_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.rptr_cuda_write_ray_queries.argtypes = [vp, vp, i32, i32]
    L.rptr_cuda_read_ray_results.argtypes = [vp, vp, i32, i32]
    L.rptr_cuda_render_ray_queries.argtypes = [vp, i32, C.POINTER(T.RenderParams), i32]
    L.rptr_cuda_normalize_options.argtypes = [vp, C.POINTER(T.RenderBackendOptions), i32]
    L.rptr_cuda_configure_for.argtypes = [vp, C.POINTER(T.RenderBackendOptions), i32, C.POINTER(T.RenderBackendOptions)]
    L.rptr_cuda_process_taa.argtypes = [vp]
    L.rptr_cuda_comm_unique_id.argtypes = [vp, C.c_size_t]
    L.rptr_cuda_comm_init_rank.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.rptr_cuda_comm_init_all.argtypes = [C.POINTER(vp), i32]
    L.rptr_cuda_comm_destroy.argtypes = [vp]
    L.rptr_cuda_reduce_framebuffer.argtypes = [vp, i32]
    L.rptr_cuda_reduce_framebuffer_all.argtypes = [C.POINTER(vp), i32, i32]
    L.rptr_write_pfm.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    if path is None:
        _lib = L
    return L


def load_sky_fit(config=None):
    """SceneParams for a SceneConfig from the table of fits shipped in data/sky_fits.json.

    The fit itself (Hosek-Wilkie, 41k lines of coefficient tables + the CIE tables) stays on the reference's side of
    the boundary: in the rptr integration the adapter calls librender's own sky_model.cpp and hands the
    resulting SceneParams to rptr_cuda_set_scene_params.  Standalone callers get the fits generated by
    oracle/gen_golden.py from the reference sources for the configurations the tests and bench use.
    """
    cfg = config or T.SceneConfig()
    key = "%.6g|%.6g,%.6g,%.6g|%.6g|%.6g,%.6g,%.6g" % (cfg.bump_scale, *cfg.sun_dir, cfg.turbidity, *cfg.albedo)
    with open(os.path.join(_HERE, "data", "sky_fits.json")) as f:
        table = json.load(f)
    if key not in table:
        raise RptrError("no pre-fitted sky for SceneConfig %s; pass a SceneParams produced by the reference's "
                        "update_sky_light (vulkan/render_sky.cpp:25-72) to update_config(scene_params=...)" % key)
    e = table[key]
    sp = T.SceneParams()
    for i in range(9):
        for j in range(4):
            sp.sky_configs[i][j] = e["sky_configs"][i][j]
    for j in range(4):
        sp.sky_radiances[j] = e["sky_radiances"][j]
        sp.sun_radiance[j] = e["sun_radiance"][j]
    for j in range(3):
        sp.sun_dir[j] = e["sun_dir"][j]
    sp.sun_cos_angle = e["sun_cos_angle"]
    sp.normal_z_scale = e["normal_z_scale"]
    return sp


class RenderConfiguration:
    """librender/render_backend.h:33-40"""

    def __init__(self, camera, time=0.0, active_variant=0, reset_accumulation=False, freeze_frame=False):
        self.camera, self.time, self.active_variant = camera, time, active_variant
        self.reset_accumulation, self.freeze_frame = reset_accumulation, freeze_frame


POINTSET_TABLE_NAMES = ("sobol_matrix", "sobol_tile_invert", "bn_sobol", "bn_scrambling_1spp")  # RPTR_POINTSET_* order


def load_pointset_tables():
    """The reference's sampler tables (SobolMatrix, SobolInversion_1_0, sobol_256spp_256d, scramblingTile_yx_d_1spp) as four
    uint32 arrays in RPTR_POINTSET_* order.

    Like the sky fit, the tables stay on the reference's side of the boundary: in the rptr integration the adapter hands
    rendering/pointsets/{sobol,bn}_tables.h to rptr_cuda_set_pointset_table the way render_sobol.cpp / render_bn.cpp
    upload them.  Standalone callers get data/pointset_tables.npz, extracted from those headers by oracle/gen_golden.py.
    """
    import numpy as np
    z = np.load(os.path.join(_HERE, "data", "pointset_tables.npz"))
    return [np.ascontiguousarray(z[k], dtype=np.uint32) for k in POINTSET_TABLE_NAMES]


class RenderCuda:
    """`RenderBackend` for `--backend cuda` (one instance = one B200)."""

    VARIANTS = ["PT_WAVEFRONT"]

    def __init__(self, device=0, library=None):
        self._L = load_library(library)
        self._h = C.c_void_p()
        if self._L.rptr_cuda_create(int(device), C.byref(self._h)) != 0:
            raise RptrError(self._L.rptr_cuda_last_error(None).decode())
        # public members mutated by callers, as in the reference (librender/render_backend.h:69-76)
        self.options = T.RenderBackendOptions()
        self.params = T.RenderParams()
        self.lighting_params = T.LightSamplingConfig()
        self.camera = T.RenderCameraParams()
        self.device = int(device)

    # -- lifetime ----------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.rptr_cuda_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise RptrError(self._L.rptr_cuda_last_error(self._h).decode())

    # -- RenderBackend ------------------------------------------------------------------------------------------------
    def name(self):
        return self._L.rptr_cuda_name().decode()

    def variant_names(self):
        return list(self.VARIANTS)

    def variant_index(self, name):
        return self.VARIANTS.index(name) if name in self.VARIANTS else -1

    def initialize(self, fb_width, fb_height):
        self._check(self._L.rptr_cuda_initialize(self._h, fb_width, fb_height))
        self.render_width, self.render_height = int(fb_width), int(fb_height)

    def set_scene(self, scene):
        d = scene.desc()
        self._check(self._L.rptr_cuda_set_scene(self._h, C.byref(d), C.byref(self.lighting_params)))

    def update_config(self, scene_config=None, scene_params=None):
        sp = scene_params if scene_params is not None else load_sky_fit(scene_config)
        self._check(self._L.rptr_cuda_set_scene_params(self._h, C.byref(sp)))

    def set_option(self, name, value):
        self._check(self._L.rptr_cuda_set_option(self._h, name.encode(), int(value)))

    def set_pointset_table(self, table, data):
        """RenderSobolVulkan / RenderBNPointsVulkan::update_random_buf: table = RPTR_POINTSET_* index, data = uint32 array."""
        import numpy as np
        a = np.ascontiguousarray(data, dtype=np.uint32)
        self._check(self._L.rptr_cuda_set_pointset_table(self._h, int(table), a.ctypes.data_as(C.POINTER(C.c_uint32)), a.size))

    def set_rng_variant(self, variant, tables=None):
        """options.rng_variant (librender/render_params.glsl.h:34-37,76) + the tables that variant reads."""
        variant = int(variant)
        if variant != 0:
            tabs = tables if tables is not None else load_pointset_tables()
            for i in {1: (2, 3), 2: (0,), 3: (0, 1)}[variant]:
                self.set_pointset_table(i, tabs[i])
        self.set_option("rng_variant", variant)
        self.options.rng_variant = variant

    def begin_frame(self, cmd_stream, config):
        self.camera = config.camera
        self._check(self._L.rptr_cuda_begin_frame(self._h, C.byref(config.camera), C.byref(self.params), C.byref(self.lighting_params),
                                                  int(config.reset_accumulation), int(config.freeze_frame), float(config.time)))

    def draw_frame(self, cmd_stream=None, variant_idx=0):
        self._check(self._L.rptr_cuda_draw_frame(self._h, variant_idx))

    def end_frame(self, cmd_stream=None, variant_idx=0):
        self._check(self._L.rptr_cuda_end_frame(self._h, variant_idx))

    def render(self, cmd_stream, config):
        """RenderBackend::render (librender/render_backend.cpp): begin + draw + end, returns stats()."""
        self.begin_frame(cmd_stream, config)
        self.draw_frame(cmd_stream, config.active_variant)
        self.end_frame(cmd_stream, config.active_variant)
        return self.stats()

    def stats(self):
        s = T.RenderStats()
        self._check(self._L.rptr_cuda_stats(self._h, C.byref(s)))
        return s

    def flush_pipeline(self):
        self._check(self._L.rptr_cuda_flush(self._h))

    # -- RenderGraphic ------------------------------------------------------------------------------------------------
    def get_framebuffer_size(self):
        w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._L.rptr_cuda_framebuffer_size(self._h, C.byref(w), C.byref(h), C.byref(c)))
        return w.value, h.value, c.value

    def readback_framebuffer(self, buffer):
        """buffer: float32 (linear HDR running mean) or uint8 (sRGB) numpy array; returns elements written or 0."""
        if buffer.dtype == np.float32:
            return self._L.rptr_cuda_readback_f32(self._h, buffer.size, buffer.ctypes.data)
        if buffer.dtype == np.uint8:
            return self._L.rptr_cuda_readback_u8(self._h, buffer.size, buffer.ctypes.data)
        raise TypeError("readback_framebuffer takes float32 or uint8 buffers")

    def framebuffer(self):
        # the float image has the render size; get_framebuffer_size() is the (upscaled) LDR target's (render_vulkan.cpp:2250-2287)
        out = np.empty((self.render_height, self.render_width, 4), np.float32)
        if self.readback_framebuffer(out) != out.size:
            raise RptrError("readback failed: " + self._L.rptr_cuda_last_error(self._h).decode())
        return out

    def framebuffer_ldr(self):
        """The sRGB8 render target (RGBA, upscaled by render_upscale_factor; after process_taa() the processed frame)."""
        w, h, c = self.get_framebuffer_size()
        out = np.empty((h, w, c), np.uint8)
        if self.readback_framebuffer(out) != out.size:
            raise RptrError("readback failed: " + self._L.rptr_cuda_last_error(self._h).decode())
        return out

    def process_taa(self):
        """ProcessTAAVulkan::process (vulkan/processing/process_taa.cpp:93-136): after end_frame, when options.enable_taa and
        params.reprojection_mode != NONE (app.cpp:517-520).  Needs option realtime_resolve."""
        self._check(self._L.rptr_cuda_process_taa(self._h))

    def readback_aov(self, aov_index, buffer):
        """RenderGraphic::readback_aov: half-float RGBA into a uint16 / float16 array; returns the element count (0 = unavailable)."""
        return self._L.rptr_cuda_readback_aov(self._h, int(aov_index), buffer.size, buffer.ctypes.data)

    def aov(self, aov_index):
        a = np.zeros((self.render_height, self.render_width, 4), np.float16)
        if self.readback_aov(aov_index, a) != a.size:
            raise RptrError("AOV %d is not available" % aov_index)
        return a

    def framebuffer_device_ptr(self):
        p = C.c_void_p()
        self._check(self._L.rptr_cuda_framebuffer_device_ptr(self._h, C.byref(p)))
        return p.value

    def stream_handle(self):
        p = C.c_void_p()
        self._check(self._L.rptr_cuda_stream_handle(self._h, C.byref(p)))
        return p.value or 0

    # -- multi-GPU: screen-space sharding + one NCCL reduce per readback (include/rptr_cuda.h) -----------------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes from rank 0, to be shipped to every rank (torch.distributed, MPI, a file ...)."""
        buf = C.create_string_buffer(128)
        L = load_library()
        if L.rptr_cuda_comm_unique_id(buf, 128) != 0:
            raise RptrError(L.rptr_cuda_last_error(None).decode())
        return buf.raw

    def comm_init_rank(self, world, rank, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._L.rptr_cuda_comm_init_rank(self._h, int(world), int(rank), buf, 128))

    @staticmethod
    def comm_init_all(backends):
        """One process driving several GPUs: a communicator over the given backends (rank = position in the list)."""
        arr = (C.c_void_p * len(backends))(*[b._h for b in backends])
        if backends[0]._L.rptr_cuda_comm_init_all(arr, len(backends)) != 0:
            raise RptrError(backends[0].last_error())

    def reduce_framebuffer(self, root=0):
        """Collective: every rank calls it; afterwards readback on `root` (every rank for root < 0) returns the whole image."""
        self._check(self._L.rptr_cuda_reduce_framebuffer(self._h, int(root)))

    @staticmethod
    def reduce_framebuffer_all(backends, root=0):
        arr = (C.c_void_p * len(backends))(*[b._h for b in backends])
        if backends[0]._L.rptr_cuda_reduce_framebuffer_all(arr, len(backends), int(root)) != 0:
            raise RptrError(backends[0].last_error())

    # -- options (librender/render_backend.h:84-85) ---------------------------------------------------------------------
    def normalize_options(self, rbo, variant_idx=0):
        self._check(self._L.rptr_cuda_normalize_options(self._h, C.byref(rbo), variant_idx))

    def configure_for(self, rbo, variant_idx=0, available_recovery_options=None):
        """True: the backend renders with `rbo` from now on.  False: unsupported (reason in last_error(); the closest supported
        set is written to available_recovery_options when given) -- the caller falls back like app.cpp:400-431."""
        avail = available_recovery_options if available_recovery_options is not None else T.RenderBackendOptions()
        if self._L.rptr_cuda_configure_for(self._h, C.byref(rbo), variant_idx, C.byref(avail)) != 0:
            return False
        self.options = T.RenderBackendOptions.from_buffer_copy(rbo)
        return True

    def last_error(self):
        return self._L.rptr_cuda_last_error(self._h).decode()

    # -- ray queries through the integrator (librender/render_backend.h:101-102) -------------------------------------------
    def enable_ray_queries(self, max_queries=T.DEFAULT_RAY_QUERY_BUDGET, max_queries_per_pixel=0):
        self._check(self._L.rptr_cuda_enable_ray_queries(self._h, int(max_queries), int(max_queries_per_pixel)))

    def ray_query_capacity(self):
        cap = C.c_size_t()
        self._check(self._L.rptr_cuda_ray_query_buffers(self._h, None, None, C.byref(cap)))
        return cap.value

    def write_ray_queries(self, queries, first=0):
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 8)
        self._check(self._L.rptr_cuda_write_ray_queries(self._h, q.ctypes.data, int(first), q.shape[0]))
        return q.shape[0]

    def render_ray_queries(self, num_queries, params=None, variant_idx=0, cmd_stream=None):
        p = params if params is not None else self.params
        self._check(self._L.rptr_cuda_render_ray_queries(self._h, int(num_queries), C.byref(p), variant_idx))
        return True

    def read_ray_results(self, n, first=0):
        res = np.zeros((n, 4), np.float32)
        self._check(self._L.rptr_cuda_read_ray_results(self._h, res.ctypes.data, int(first), int(n)))
        return res

    # -- RaytraceBackend ----------------------------------------------------------------------------------------------
    def trace_ray(self, queries):
        """queries: (n, 8) float32 rows laid out as RenderRayQuery -> ((n, 4) result words, (n,) hit distance)."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 8)
        res = np.zeros((q.shape[0], 4), np.float32)
        t = np.zeros(q.shape[0], np.float32)
        self._check(self._L.rptr_cuda_trace_rays(self._h, q.ctypes.data, q.shape[0], res.ctypes.data, t.ctypes.data))
        return res, t

    # -- extras ---------------------------------------------------------------------------------------------------------
    def lights(self):
        n = self._L.rptr_cuda_get_lights(self._h, None, 0)
        arr = (T.TriLightData * max(n, 1))()
        self._L.rptr_cuda_get_lights(self._h, arr, n)
        return np.frombuffer(arr, dtype=np.float32).reshape(-1, 12)[:n].copy()

    def counters(self):
        c = Counters()
        self._check(self._L.rptr_cuda_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        self._check(self._L.rptr_cuda_reset_counters(self._h))

    def frame_state(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._L.rptr_cuda_frame_state(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def render_spp(self, camera, spp, batch_spp=None, reset=True):
        """Validation-mode loop (libapp/app_state.h:90-99): frames of batch_spp until `spp` samples are accumulated."""
        done = 0
        batch = batch_spp or self.params.batch_spp
        while done < spp:
            self.params.batch_spp = min(batch, spp - done)  # next_frame_spp clamps the last frame
            cfg = RenderConfiguration(camera, reset_accumulation=(reset and done == 0))
            self.begin_frame(None, cfg)
            self.draw_frame(None, 0)
            self.end_frame(None, 0)
            done += self.params.batch_spp
        self.params.batch_spp = batch
        return self.stats()


def write_pfm(prefix, pixels):
    """WriteImage::write_pfm (util/write_image.cpp:34-66) through the library's host helper."""
    px = np.ascontiguousarray(pixels, dtype=np.float32)
    h, w, c = px.shape
    if load_library().rptr_write_pfm(str(prefix).encode(), w, h, c, px.ctypes.data) != 0:
        raise RptrError("write_pfm failed for %s" % prefix)


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(x) for x in f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, 3)
    return data[::-1].copy()  # top row first


def create_cuda_backend(display=None, device=0):
    """The factory the reference calls: typedef RenderBackend* (*create_backend_function)(Display&)."""
    return RenderCuda(device=device)
