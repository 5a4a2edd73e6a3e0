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


# ---- motion / jitter image (AOVMotionJitterIndex; vulkan/accumulate.glsl:77-87, render_vulkan.cpp:1986-1998, 2926-2930) ------
def moved(camera, dx):
    c = T.RenderCameraParams.from_buffer_copy(camera)
    c.pos[0] += dx
    return c


def test_view_projection_is_a_projection_of_the_camera(hostsim, oracle):
    """VP restated from glm's published algorithms (glm is not in the reference tree): product and oracle agree bit for bit,
    and the matrix does what render_vulkan.cpp:2926-2930 says -- the view centre lands on NDC (0, 0), pixel centres on their
    NDC positions (y down, GLToVulkan), depth w = distance along dir."""
    lib = C.CDLL(hostsim)
    lib.hostsim_view_projection.argtypes = [C.POINTER(T.RenderCameraParams), C.c_int32, C.c_int32, oracle.f32p]
    s = scenes.cornell_box()
    W, H = 160, 90
    vp = oracle.view_projection(s.camera, W, H)
    mine = np.zeros(16, np.float32)
    lib.hostsim_view_projection(C.byref(s.camera), W, H, oracle._fp(mine))
    assert np.array_equal(vp.view(np.uint32), mine.view(np.uint32))
    M = vp.reshape(4, 4).T.astype(np.float64)  # column-major -> rows
    du_dv_tl = oracle.view_params(s.camera, W, H).astype(np.float64)
    du, dv, tl = du_dv_tl[:3], du_dv_tl[3:6], du_dv_tl[6:]
    pos = np.array(list(s.camera.pos), np.float64)
    for (px, py) in [(0.5, 0.5), (0.25, 0.75), (0.9, 0.1)]:
        p = pos + 3.0 * (tl + px * du + py * dv)
        clip = M @ np.append(p, 1.0)
        ndc = clip[:2] / clip[3]
        assert np.allclose(ndc, [2 * px - 1, 2 * py - 1], atol=2e-5), (px, py, ndc)
        assert clip[3] > 0


def test_motion_jitter_image_matches_oracle(hostsim, oracle):
    lib = C.CDLL(hostsim)
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_render_sample_aov3.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_uint32, oracle.f32p, oracle.f32p]
    s = scenes.random_triangles(8000)
    sp = load_sky_fit()
    o = oracle.OracleScene(s)
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = lib.hostsim_scene_create(C.byref(d), C.byref(ls))
    W, H = 160, 90
    taa = T.RenderParams()
    taa.enable_raster_taa = 1
    prev = moved(s.camera, -0.05)
    cases = {
        "first frame": dict(first_sample=0),                                                        # VP_reference = 0
        "static camera": dict(first_sample=2, vp_reference=oracle.view_projection(s.camera, W, H)),
        "moved camera": dict(first_sample=2, vp_reference=oracle.view_projection(prev, W, H)),
        "raster taa": dict(first_sample=3, frame_offset=6, params=taa, vp_reference=oracle.view_projection(prev, W, H)),
    }
    out = {}
    try:
        for name, kw in cases.items():
            sample = kw["first_sample"]
            ar, nd, mj = o.render_aov3(W, H, s.camera, sp, sample, **kw)
            a = o._args(W, H, s.camera, sp, **kw)
            img = np.zeros((H, W, 4), np.float32)
            aov = np.zeros((H, W, 12), np.float32)
            lib.hostsim_render_sample_aov3(hs, C.byref(a), sample, oracle._fp(img), oracle._fp(aov))
            assert np.array_equal(aov[..., :4].view(np.uint32), ar.view(np.uint32)), name
            assert np.array_equal(aov[..., 4:8].view(np.uint32), nd.view(np.uint32)), name
            assert np.array_equal(aov[..., 8:].view(np.uint32), mj.view(np.uint32)), name
            out[name] = (nd, mj)
    finally:
        lib.hostsim_scene_destroy(hs)
    nd, mj = out["first frame"]
    assert np.isnan(mj[..., :2]).all() and (mj[..., 2:] == 0).all()  # 0 / max(0, 0): no reference view yet
    nd, mj = out["static camera"]
    hit = np.isfinite(nd[..., 3])
    assert (mj[hit] == 0).all() and (mj[..., 2:] == 0).all()
    # misses project vec3(2e32): behind this camera, so w clamps to 0 and inf - inf = NaN (in front it would be 0)
    assert (np.isnan(mj[~hit][:, :2]) | (mj[~hit][:, :2] == 0)).all()
    nd, mj = out["moved camera"]
    hit = np.isfinite(nd[..., 3])
    assert hit.any() and (mj[..., 2:] == 0).all()
    # camera moved +x by 0.05: every visible point was further right on the reference frame; parallax ~ 1 / depth
    assert (mj[hit][:, 0] > 0).all() and np.abs(mj[hit][:, 1]).max() < 1e-3
    near, far = nd[..., 3] < np.median(nd[hit][:, 3]), hit & (nd[..., 3] > np.median(nd[hit][:, 3]))
    assert mj[near & hit][:, 0].mean() > mj[far][:, 0].mean()
    nd, mj = out["raster taa"]
    sj = np.zeros(2, np.float32)
    oracle.lib().oracle_screen_jitter(6, 3, W, H, oracle._fp(sj))
    assert (mj[..., 2] == sj[0]).all() and (mj[..., 3] == sj[1]).all() and (sj != 0).any()
