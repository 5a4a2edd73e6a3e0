"""The fp16 AOV images (RenderGraphic::readback_aov; vulkan/accumulate.glsl:89-103, pt_megakernel.glsl:482-486, 670-672,
shade_base_material.glsl:28-31): albedo + roughness and normal + depth of the first path vertex.  CPU half: the product's
shading code (tests/hostsim) against the oracle on the float values behind the images."""
import ctypes as C

import numpy as np
import pytest

from realtimepathtracingresearchframework_b200 import load_sky_fit, scenes, types as T


@pytest.mark.parametrize("make", [scenes.cornell_box, lambda: scenes.random_triangles(8000), scenes.alpha_tested_soup])
def test_first_vertex_attributes_match_oracle(hostsim, oracle, make):
    lib = C.CDLL(hostsim)
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_render_sample_aov.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_uint32, oracle.f32p, oracle.f32p]
    s = make()
    sp = load_sky_fit()
    o = oracle.OracleScene(s)
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = lib.hostsim_scene_create(C.byref(d), C.byref(ls))
    W, H = 160, 90
    ar, nd = o.render_aov(W, H, s.camera, sp, 1)
    a = o._args(W, H, s.camera, sp, first_sample=1)
    img = np.zeros((H, W, 4), np.float32)
    aov = np.zeros((H, W, 8), np.float32)
    lib.hostsim_render_sample_aov(hs, C.byref(a), 1, oracle._fp(img), oracle._fp(aov))
    lib.hostsim_scene_destroy(hs)
    assert np.array_equal(aov[..., :4].view(np.uint32), ar.view(np.uint32))
    assert np.array_equal(aov[..., 4:].view(np.uint32), nd.view(np.uint32))
    hit = np.isfinite(nd[..., 3])
    assert hit.any()
    # hits: unit normal, positive finite depth, roughness in (0, 1]; misses: zero normal, infinite depth, roughness 1
    assert np.allclose(np.linalg.norm(nd[hit][:, :3], axis=1), 1.0, atol=1e-5) and (nd[hit][:, 3] > 0).all()
    assert (ar[hit][:, 3] > 0).all() and (ar[hit][:, 3] <= 1).all()
    if (~hit).any():
        assert (nd[~hit][:, :3] == 0).all() and np.isinf(nd[~hit][:, 3]).all() and (ar[~hit] == [0, 0, 0, 1]).all()
