"""A small tour of every kernel variant for compute-sanitizer (memcheck / racecheck / initcheck), sized to finish under the
tool's 10-100x slow-down:

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py

Covers: persistent trace kernels (closest / any-hit, with and without the alpha filter), the one-ray-per-thread kernels,
k_shade in its three feature variants (LCG, triangle lights, ALL = QMC + AOV + transmission + normal maps), raygen,
resolve (progressive + discard-history), the tile sort of multi-material scenes, LDR / AOV read-backs, ray queries, the
device builder builder and screen-space sharding (tile_rank / tile_world)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtimepathtracingresearchframework_b200 import RenderConfiguration, RenderCuda, load_pointset_tables, scenes, types as T  # noqa: E402

W, H = 96, 54


def backend(scene, **options):
    r = RenderCuda(device=0)
    r.initialize(W, H)
    for k, v in options.items():
        r.set_option(k, v)
    r.set_scene(scene)
    r.update_config(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    return r


def tour():
    tables = load_pointset_tables()
    n = 0
    # LCG path, host SAH and device builder, persistent and one-ray-per-thread kernels, several waves per frame
    for opts in (dict(), dict(bvh_builder=0), dict(trace_kernel=1), dict(wave_paths=W * H), dict(overlap_shadow=0, stage_timing=1)):
        r = backend(scenes.random_triangles(3000), **opts)
        r.render_spp(scenes.random_triangles(3000).camera, 3, batch_spp=3)
        assert np.isfinite(r.framebuffer()).all()
        r.close()
        n += 1
    # alpha-tested, textured, normal-mapped materials with every pointset; AOV + LDR read-backs
    s = scenes.alpha_tested_soup(4000)
    for variant in (0, 1, 2, 3):
        r = backend(s, transmission=1)
        if variant:
            r.set_rng_variant(variant, tables)
        r.params.enable_raster_taa = variant & 1
        r.render_spp(s.camera, 2, batch_spp=2)
        for i in range(3):
            r.aov(i)
        for channel in (0, 1, 2, 3):
            r.params.output_channel = channel
            r.params.early_tone_mapping_mode = channel - 1
            r.render_spp(s.camera, 1, reset=False)
            ldr = np.zeros((H, W, 4), np.uint8)
            assert r.readback_framebuffer(ldr) == ldr.size
        r.close()
        n += 1
    # triangle-light NEE (binned RIS), instancing, discard-history resolve, ray queries
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_hostsim_parity import emissive_soup
    s = emissive_soup()
    r = backend(s)
    r.params.reprojection_mode = 1
    r.render_spp(s.camera, 2)
    rng = np.random.default_rng(1)
    q = np.zeros((500, 8), np.float32)
    q[:, 0:3] = rng.uniform(-2, 2, (500, 3))
    d = rng.normal(size=(500, 3))
    q[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    q[:, 7] = 1e20
    q[::7, 3] = -1.0  # skipped queries
    r.trace_ray(q)
    r.close()
    n += 1
    # round 2: image textures with mip chains and block compression (footprint level of detail, anisotropic taps), the temporal passes
    # (reprojection accumulate + TAA on an upscaled LDR target) over a moving camera, a .vks scene file
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_texture_lod import textured_mip_scene
    s = textured_mip_scene(True)
    r = backend(s, transmission=1, realtime_resolve=1, render_upscale_factor=2)
    r.initialize(W, H)   # the upscale factor sizes the LDR target at initialize
    r.set_scene(s)
    r.update_config(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    r.params.reprojection_mode = T.REPROJECTION_MODE_ACCUMULATE
    for k in range(4):
        cam = T.RenderCameraParams.from_buffer_copy(s.camera)
        cam.pos[0] += 0.05 * k
        r.params.batch_spp = 1
        r.render(None, RenderConfiguration(cam, reset_accumulation=(k == 0)))
        r.process_taa()
        assert r.framebuffer_ldr().shape == (2 * H, 2 * W, 4)
    r.close()
    n += 1
    import tempfile
    import vks_util
    from realtimepathtracingresearchframework_b200 import vks
    with tempfile.TemporaryDirectory() as d:
        path, _ = vks_util.write_test_scene(d)
        s = vks.load_vks(path)
    r = backend(s, transmission=1)
    r.render_spp(scenes.look_at_camera((0, 2, 14), (0, 0, 0), fovy=50.0), 2)
    assert np.isfinite(r.framebuffer()).all()
    r.close()
    n += 1
    # screen-space sharding: two ranks of the same frame
    s = scenes.cornell_box()
    for rank in (0, 1):
        r = backend(s, tile_world=2, tile_rank=rank, tile_rows=8)
        r.render_spp(s.camera, 2)
        r.framebuffer()
        r.close()
        n += 1
    return n


def tiny_tour():
    """racecheck-sized: 32 x 18 frames, a few hundred triangles -- the persistent closest-hit / any-hit kernels on their two streams
    (the in-place illum update of the shadow kernel beside the next closest-hit launch), the alpha variants, sub-waves side by
    side, ray queries through the integrator and the RaytraceBackend service.  Run against a build with small trace CTAs:
        RPTR_CUDA_LIB=variants/librptr_cuda_t128.so compute-sanitizer --tool racecheck python tools/sanitize.py --tiny"""
    global W, H
    W, H = 32, 18
    n = 0
    for s, opts in ((scenes.random_triangles(400, box=3.0, edge=0.8), dict()), (scenes.alpha_tested_soup(600), dict()),
                    (scenes.random_triangles(400, box=3.0, edge=0.8), dict(concurrent_waves=2)),
                    (scenes.alpha_tested_soup(600), dict(reorder_bounce=2, reorder_shadow=3)),  # binned queues (rptr_reorder.cuh)
                    (scenes.random_triangles(400, box=3.0, edge=0.8), dict(tail_kernel=0))):   # every other context hands its tails to k_trace_tail
        s.camera = scenes.look_at_camera((0, 0, 10), (0, 0, 0), fovy=50.0)
        r = backend(s, **opts)
        r.render_spp(s.camera, 4, batch_spp=4)
        assert np.isfinite(r.framebuffer()).all()
        r.enable_ray_queries(256, 0)
        rng = np.random.default_rng(2)
        q = np.zeros((200, 8), np.float32)
        q[:, 0:3] = rng.uniform(-3, 3, (200, 3))
        d = rng.normal(size=(200, 3))
        q[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
        q[:, 7] = 1e20
        r.write_ray_queries(q)
        r.render_ray_queries(200)
        r.read_ray_results(200)
        r.trace_ray(q)
        r.close()
        n += 1
    return n


if __name__ == "__main__":
    if "--tiny" in sys.argv:
        print("sanitize tiny tour: %d contexts ok" % tiny_tour())
    else:
        print("sanitize tour: %d contexts ok" % tour())
