"""Regression guard for the checker itself: the oracle's images for a few small frames are frozen as SHA-256 digests.  These
are the oracle outputs the CUDA path was bit-identical to on a B200 at the end of round 1 (same scenes as tests/test_gpu_parity.py
at reduced size); a later change that moves any of them has changed the oracle's arithmetic, not just its structure, and
needs the GPU parity run again -- then re-freeze with `python tests/test_oracle_frozen.py --print`."""
import hashlib
import sys

import numpy as np
import pytest

FROZEN = {
    "cornell 96x54 3spp": "6c90e17101c64348f84b28cae2709de6",
    "random3000 sun 96x54 3spp": "b699c8f2792e83ab2537cf0e6fe2cf11",
    "alpha soup 96x54 2spp offset 5": "8d68cf9c6c7b973f68f4625692d28ebe",
    "emissive soup 96x54 2spp": "85fc2300281021b9cb21f9c1e0d0ff08",
    "transmission 96x54 2spp": "383544632e3eda5d85f82f002737b50a",
    "sobol 96x54 2spp": "7fbca7d61902d2ccca04e2da6d593535",
    "blue noise batch 96x54 4spp": "195e6eefa66b97a8b581274f07fc7819",
    "z-sobol alpha 96x54 2spp": "f16c9c3c32a1955ce1c1b4231733ae84",
}


def cases():
    from realtimepathtracingresearchframework_b200 import load_pointset_tables, load_sky_fit, scenes, types as T
    from test_hostsim_parity import emissive_soup
    sun = dict(sun_dir=(0.35, 0.8, 0.45))
    tables = load_pointset_tables()
    tr = scenes.random_triangles(3000)
    for j, m in enumerate(tr.materials):
        if j % 2 == 1:
            m.specular_transmission, m.metallic = 0.8, 0.0
            m.flags = T.BASE_MATERIAL_NOALPHA | T.BASE_MATERIAL_EXTENDED
    return {
        "cornell 96x54 3spp": (scenes.cornell_box(), {}, dict(spp=3)),
        "random3000 sun 96x54 3spp": (scenes.random_triangles(3000), sun, dict(spp=3)),
        "alpha soup 96x54 2spp offset 5": (scenes.alpha_tested_soup(3000), sun, dict(spp=2, frame_offset=5)),
        "emissive soup 96x54 2spp": (emissive_soup(), sun, dict(spp=2)),
        "transmission 96x54 2spp": (tr, {}, dict(spp=2, transmission=1)),
        "sobol 96x54 2spp": (scenes.random_triangles(3000), sun, dict(spp=2, rng_variant=2, pointset_tables=tables)),
        "blue noise batch 96x54 4spp": (scenes.random_triangles(3000), sun, dict(spp=4, batch_spp=4, rng_variant=1, pointset_tables=tables)),
        "z-sobol alpha 96x54 2spp": (scenes.alpha_tested_soup(3000), sun, dict(spp=2, rng_variant=3, pointset_tables=tables)),
    }


def digest(oracle, scene, sky, kw):
    from realtimepathtracingresearchframework_b200 import load_sky_fit, types as T
    img, _ = oracle.OracleScene(scene).render(96, 54, scene.camera, load_sky_fit(T.SceneConfig(**sky)), **kw)
    assert np.isfinite(img).all() and img[..., :3].max() > 0
    return hashlib.sha256(np.ascontiguousarray(img, np.float32).tobytes()).hexdigest()[:32]


def product_code_agrees(hostsim, oracle, scene, sky, kw):
    """first sample layer: the product's host build (independent code) against the oracle, bit for bit"""
    import ctypes as C
    from realtimepathtracingresearchframework_b200 import load_sky_fit, types as T
    H = C.CDLL(hostsim)
    H.hostsim_scene_create.restype = C.c_void_p
    H.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    H.hostsim_scene_destroy.argtypes = [C.c_void_p]
    H.hostsim_render_sample.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_uint32, oracle.f32p]
    kw = {k: v for k, v in kw.items() if k not in ("spp", "batch_spp")}
    sp = load_sky_fit(T.SceneConfig(**sky))
    o = oracle.OracleScene(scene)
    ref = o.render_sample(96, 54, scene.camera, sp, 0, **kw)
    a = o._args(96, 54, scene.camera, sp, first_sample=0, **kw)
    ls = T.LightSamplingConfig()
    d = scene.desc()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    img = np.zeros((54, 96, 4), np.float32)
    H.hostsim_render_sample(hs, C.byref(a), 0, oracle._fp(img))
    H.hostsim_scene_destroy(hs)
    return np.array_equal(ref.view(np.uint32), img.view(np.uint32))


@pytest.mark.parametrize("name", sorted(FROZEN))
def test_oracle_image_is_unchanged(hostsim, oracle, name):
    scene, sky, kw = cases()[name]
    if digest(oracle, scene, sky, kw) == FROZEN[name]:
        return
    # the digests were taken with this image's libm (tanf of the camera basis, pow of the sRGB decode).  If the independent
    # product build still agrees with the oracle bit for bit, the platform moved, not the oracle.
    if product_code_agrees(hostsim, oracle, scene, sky, kw):
        pytest.skip("digest differs but product code and oracle agree bit for bit: different libm / platform; re-freeze here")
    pytest.fail("the oracle's arithmetic changed for: " + name)


def test_every_case_is_frozen():
    assert sorted(cases()) == sorted(FROZEN)


if __name__ == "__main__" and "--print" in sys.argv:
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle
    for name, (scene, sky, kw) in cases().items():
        print('    "%s": "%s",' % (name, digest(pyoracle, scene, sky, kw)))
