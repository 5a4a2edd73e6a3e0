"""Drives oracle/_ref's whole-path driver (oracle/ref_shim/ref_path.cpp): one path sample composed from the reference-executed
pieces, with closest hit / occlusion answered by the oracle's intersection code through C function pointers."""
import ctypes as C

import numpy as np

from realtimepathtracingresearchframework_b200 import types as T


def _cases():
    from realtimepathtracingresearchframework_b200 import scenes
    from test_hostsim_parity import emissive_soup
    return {
        # BASELINE configs[0] (reduced frame), Lambert + emissive quad: tri-light NEE, emitter MIS
        "cornell": (scenes.cornell_box, dict(), (192, 108), 3),
        # configs[1]-style scene: diffuse + GGX, sun + sky NEE, slanted sun
        "random20k": (lambda: scenes.random_triangles(20000), dict(sun_dir=(0.35, 0.8, 0.45)), (160, 90), 3),
        # per-triangle material ids, a transformed instance, several light bins, p_sun = 0.5
        "emissive_instanced": (emissive_soup, dict(sun_dir=(0.35, 0.8, 0.45)), (128, 72), 2),
    }


class _Cases(dict):
    def __missing__(self, k):
        self.update(_cases())
        return dict.__getitem__(self, k)

    def items(self):
        if not len(self):
            self.update(_cases())
        return dict.items(self)


CASES = _Cases()


class RefPathArgs(C.Structure):
    _fields_ = [("cam", C.c_float * 12), ("width", C.c_uint32), ("height", C.c_uint32), ("frame_offset", C.c_uint32), ("frame_id", C.c_uint32),
                ("max_path_depth", C.c_int32), ("rr_path_depth", C.c_int32), ("glossy_only_mode", C.c_int32), ("output_channel", C.c_int32),
                ("sp", T.SceneParams), ("materials", C.c_void_p), ("lights", C.c_void_p), ("n_lights", C.c_int32), ("bin_size", C.c_int32),
                ("closest", C.c_void_p), ("occluded", C.c_void_p), ("user", C.c_void_p)]


def ref_path_render(oracle_scene, scene, width, height, camera, scene_params, spp, params=None, frame_offset=0, region=None):
    """Running mean over `spp` frames of one sample rendered by the composed reference path -> (H, W, 4) float32."""
    from oracle import pyoracle as po
    R, L = po.ref(), po.lib()
    if R is None or not hasattr(R, "ref_path_render"):
        raise RuntimeError("oracle/_ref/libref.so with the whole-path driver is not built")
    p = params or T.RenderParams()
    a = RefPathArgs()
    vp = po.view_params(camera, width, height)  # du, dv, top_left (pinned to the reference's update_view_parameters)
    a.cam[0:3] = list(camera.pos)
    a.cam[3:12] = [float(x) for x in vp]
    a.width, a.height, a.frame_offset, a.frame_id = width, height, frame_offset, 0
    a.max_path_depth, a.rr_path_depth, a.glossy_only_mode, a.output_channel = p.max_path_depth, p.rr_path_depth, p.glossy_only_mode, p.output_channel
    a.sp = T.SceneParams.from_buffer_copy(scene_params)
    lights = oracle_scene.lights()
    n_lights = len(lights)
    a.sp.sun_radiance[3] = a.sp.sun_radiance[3] * 0.5 if n_lights > 0 else 1.0  # vulkan/render_sky.cpp:67-70
    mats = (T.BaseMaterial * len(scene.materials))(*scene.materials)
    larr = np.ascontiguousarray(lights, np.float32)
    a.materials = C.cast(mats, C.c_void_p)
    a.lights = larr.ctypes.data if n_lights else None
    a.n_lights, a.bin_size = n_lights, oracle_scene.lighting.bin_size
    a.closest = C.cast(L.oracle_cb_closest, C.c_void_p)
    a.occluded = C.cast(L.oracle_cb_occluded, C.c_void_p)
    a.user = oracle_scene.h
    img = np.zeros((height, width, 4), np.float32)
    x0, y0, x1, y1 = region or (0, 0, width, height)
    R.ref_path_render.argtypes = [C.POINTER(RefPathArgs), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_int32, po.f32p]
    R.ref_path_render(C.byref(a), x0, y0, x1, y1, 0, spp, img.ctypes.data_as(po.f32p))
    return img
