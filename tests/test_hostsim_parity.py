"""The code the CUDA kernels execute (csrc/rptr_shading.cuh, rptr_bvh.cuh are __host__ __device__; csrc/rptr_host.cpp is
the product's scene ingestion) compiled for the CPU by tests/hostsim and compared BIT FOR BIT with the independent
oracle: different source, different BVH, same arithmetic contract.  This is the CPU-runnable half of the parity gate;
tests/test_gpu_parity.py repeats it on the device through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from realtimepathtracingresearchframework_b200 import load_sky_fit, scenes, types as T


@pytest.fixture(scope="module")
def H(hostsim, oracle):
    lib = C.CDLL(hostsim)
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_render_sample.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_uint32, oracle.f32p]
    lib.hostsim_get_lights.argtypes = [C.c_void_p, C.c_void_p]
    lib.hostsim_num_lights.argtypes = [C.c_void_p]
    return lib


def emissive_soup(n=3000):
    """random triangles with per-triangle material ids, a few emissive -> several light bins + p_sun = 0.5"""
    s = scenes.Scene()
    g = scenes.random_triangle_grid(n, seed=77, box=4.0, edge=0.6)
    scale, base = 2.0 ** -16, -16.0
    geo = scenes.Geometry(scenes.pack_qverts(g.reshape(-1, 3)), (scale,) * 3, (base + 2.0 ** -17,) * 3)
    mesh = s.add_mesh([geo])
    s.materials = [T.BaseMaterial(base_color=(0.7, 0.6, 0.5), roughness=0.4, ior=1.5, flags=T.BASE_MATERIAL_NOALPHA),
                   T.BaseMaterial(base_color=(0.2, 0.5, 0.8), roughness=0.15, metallic=1.0, flags=T.BASE_MATERIAL_NOALPHA),
                   T.BaseMaterial(base_color=(1.0, 0.9, 0.7), emission_intensity=25.0, flags=T.BASE_MATERIAL_NOALPHA),
                   T.BaseMaterial(base_color=(0.4, 0.9, 0.4), emission_intensity=3.0, ior=1.0, flags=T.BASE_MATERIAL_NOALPHA)]
    ids = (np.arange(n) % 2).astype(np.uint8)
    ids[::37] = 2
    ids[5::91] = 3
    pm = s.add_pmesh(mesh, [0], tri_material_ids=ids)
    s.add_instance(pm)
    # a second, transformed instance of the same mesh (non-uniform scale + rotation + translation)
    t = np.array([[0.8, -0.3, 0.1, 3.0], [0.2, 0.9, 0.0, -1.0], [0.0, 0.1, 1.1, 0.5]], np.float32)
    s.add_instance(pm, t)
    s.camera = scenes.look_at_camera((0, 1, 14), (0.5, 0, 0), fovy=50.0)
    return s


CASES = {
    "cornell": (scenes.cornell_box, dict()),
    "random20k": (lambda: scenes.random_triangles(20000), dict()),
    "random20k_slanted_sun": (lambda: scenes.random_triangles(20000), dict(sun_dir=(0.35, 0.8, 0.45))),
    "emissive_instanced": (emissive_soup, dict(sun_dir=(0.35, 0.8, 0.45))),
    # vertex normals + uvs (rendering/rt/hit.glsl:58-128): smooth shading, uv-derivative tangents under one-texel normal
    # maps, normals on the far side of the geometric normal, instances with non-uniform scale and a mirrored instance
    "smooth_shaded": (scenes.smooth_shaded_scene, dict(sun_dir=(0.35, 0.8, 0.45))),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_product_code_matches_oracle_bit_for_bit(H, oracle, case):
    make, sky = CASES[case]
    s = make()
    sp = load_sky_fit(T.SceneConfig(**sky))
    o = oracle.OracleScene(s)
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    assert hs
    try:
        n = H.hostsim_num_lights(hs)
        arr = (T.TriLightData * max(n, 1))()
        H.hostsim_get_lights(hs, arr)
        assert np.array_equal(np.frombuffer(arr, np.float32).reshape(-1, 12)[:n], o.lights()), "binned light buffers differ"
        W, Hh = 160, 90
        for sample in (0, 3):
            ref = o.render_sample(W, Hh, s.camera, sp, sample)
            a = o._args(W, Hh, s.camera, sp)
            img = np.zeros((Hh, W, 4), np.float32)
            H.hostsim_render_sample(hs, C.byref(a), sample, oracle._fp(img))
            assert np.isfinite(ref).all()
            assert ref[..., :3].max() > 0
            assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), "%d pixels differ" % (ref != img).any(-1).sum()
    finally:
        H.hostsim_scene_destroy(hs)


def test_transmission_build_matches_oracle(H, oracle):
    """GLTF_SUPPORT_TRANSMISSION[_ROUGHNESS] variant (pipeline_pt hit groups): thick (ONESIDED) + thin glass."""
    s = scenes.random_triangles(4000)
    for j, m in enumerate(s.materials):
        if j % 4 == 1:
            m.specular_transmission, m.metallic, m.flags = 0.9, 0.0, T.BASE_MATERIAL_NOALPHA | T.BASE_MATERIAL_ONESIDED | T.BASE_MATERIAL_EXTENDED
        if j % 4 == 3:
            m.specular_transmission, m.metallic, m.flags = 0.7, 0.0, T.BASE_MATERIAL_NOALPHA | T.BASE_MATERIAL_EXTENDED
    sp = load_sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    o = oracle.OracleScene(s)
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    W, Hh = 128, 72
    ref = o.render_sample(W, Hh, s.camera, sp, 1, transmission=1)
    a = o._args(W, Hh, s.camera, sp, transmission=1)
    img = np.zeros((Hh, W, 4), np.float32)
    H.hostsim_render_sample(hs, C.byref(a), 1, oracle._fp(img))
    H.hostsim_scene_destroy(hs)
    assert np.array_equal(ref.view(np.uint32), img.view(np.uint32))
    # and transmission really changes the image
    assert not np.array_equal(ref, o.render_sample(W, Hh, s.camera, sp, 1, transmission=0))


def test_traversal_edge_rays_against_bruteforce(hostsim, oracle):
    """The product's BVH traversal (fma slabs, padded boxes) against the oracle's brute-force loop over all triangles:
    exactly axis-parallel rays (d.k == +-0, what sun shadow rays at the zenith produce), rays starting on surfaces,
    rays along box faces of axis-aligned walls."""
    lib = C.CDLL(hostsim)
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, oracle.f32p, oracle.f32p, C.c_int32]
    rng = np.random.default_rng(3)
    for s, box in ((scenes.cornell_box(), 1.0), (scenes.random_triangles(3000, box=2.0, edge=0.4), 2.5)):
        o = oracle.OracleScene(s)
        ls = T.LightSamplingConfig()
        d = s.desc()
        hs = lib.hostsim_scene_create(C.byref(d), C.byref(ls))
        n = 6000
        q = np.zeros((n, 8), np.float32)
        q[:, 0:3] = rng.uniform(-box, box, (n, 3)) + (np.array([0, 1, 0]) if box == 1.0 else 0)
        q[:, 7] = 1e20
        axes = np.eye(3, dtype=np.float32)
        for i in range(n):
            k = i % 6
            v = axes[k % 3] * (1 if k < 3 else -1)
            if i % 4 == 1:
                v = v + axes[(k + 1) % 3] * np.float32(rng.uniform(-1, 1))  # parallel to one axis plane only
            if i % 4 == 2:
                v = np.where(v == 0, np.float32(-0.0), v)  # negative zeros
            q[i, 4:7] = v / np.linalg.norm(v)
        # a third of the origins snapped onto grid planes of the geometry (on walls / box faces)
        q[::3, 0] = np.round(q[::3, 0])
        ref, tref = o.trace_closest(q, bruteforce=True)
        got, tgot = np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
        lib.hostsim_trace(hs, q.ctypes.data, n, oracle._fp(got), oracle._fp(tgot), 0)
        assert (tref >= 0).mean() > 0.2
        assert np.array_equal(tref, tgot) and np.array_equal(ref.view(np.uint32), got.view(np.uint32))
        # oracle's own BVH too
        ob, tob = o.trace_closest(q)
        assert np.array_equal(tref, tob) and np.array_equal(ref.view(np.uint32), ob.view(np.uint32))


def test_c4_style_instanced_scene_matches_oracle(H, oracle):
    """BASELINE configs[3] at reduced size: one mesh instanced 25 times through .vks-style quantised similarity
    transforms, per-triangle u8 material ids, GGX + thick/thin transmission + emissive triangles -> binned tri-light NEE."""
    s = scenes.instanced_scene(1500, 25)
    sp = load_sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    o = oracle.OracleScene(s)
    assert len(o.lights()) >= 25 * 64
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    n = H.hostsim_num_lights(hs)
    arr = (T.TriLightData * n)()
    H.hostsim_get_lights(hs, arr)
    assert np.array_equal(np.frombuffer(arr, np.float32).reshape(-1, 12), o.lights())
    W, Hh = 200, 112
    ref = o.render_sample(W, Hh, s.camera, sp, 2, transmission=1)
    a = o._args(W, Hh, s.camera, sp, transmission=1, first_sample=2)  # view_params.frame_id seeds the alpha test of shadow rays
    img = np.zeros((Hh, W, 4), np.float32)
    H.hostsim_render_sample(hs, C.byref(a), 2, oracle._fp(img))
    H.hostsim_scene_destroy(hs)
    assert (ref[..., 3] > 0).mean() > 0.02
    assert np.array_equal(ref.view(np.uint32), img.view(np.uint32))


def test_host_bvh_does_not_depend_on_the_builder_thread_count(hostsim, monkeypatch):
    """build_bvh (csrc/rptr_host.cpp) builds the top of the tree serially and the sub-trees on worker threads: the 4-wide
    tree and the leaf order must be identical for any thread count (and the traversal result is independent of both)."""
    lib = C.CDLL(hostsim)
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_bvh_hash.restype = C.c_uint64
    lib.hostsim_bvh_hash.argtypes = [C.c_void_p]
    lib.hostsim_num_nodes.argtypes = [C.c_void_p]
    s = scenes.random_triangles(150000)  # large enough for the threaded path (> 65536 triangles)
    ls = T.LightSamplingConfig()
    seen = {}
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("RPTR_BUILD_THREADS", threads)
        d = s.desc()
        hs = lib.hostsim_scene_create(C.byref(d), C.byref(ls))
        seen[threads] = (lib.hostsim_bvh_hash(hs), lib.hostsim_num_nodes(hs))
        lib.hostsim_scene_destroy(hs)
    assert seen["1"] == seen["3"] == seen["8"] and seen["1"][1] > 30000


def test_wide_bvh_structure_and_conservative_boxes(hostsim):
    """The eight-wide nodes (csrc/rptr_bvh.cuh: 7-bit bounds on the node's own grid) as the traversal decodes them: every node
    and triangle referenced once, and every triangle strictly inside the decoded boxes of all its ancestors -- on a soup, on
    instances with large offsets, and on a scene whose triangles are tiny against its extent."""
    lib = C.CDLL(hostsim)
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_validate_bvh.argtypes = [C.c_void_p, C.c_void_p]
    ls = T.LightSamplingConfig()
    cases = [scenes.random_triangles(60000), scenes.instanced_scene(3000, 12), scenes.random_triangles(20000, box=10.0, edge=0.0008),
             scenes.cornell_box(), scenes.smooth_shaded_scene()]
    for s in cases:
        d = s.desc()
        hs = lib.hostsim_scene_create(C.byref(d), C.byref(ls))
        assert hs
        out = (C.c_int64 * 4)()
        lib.hostsim_validate_bvh(hs, out)
        lib.hostsim_scene_destroy(hs)
        nodes, tris, bad, depth = list(out)
        assert tris > 0 and nodes > 0 and bad == 0 and depth <= 32, (s.name, nodes, tris, bad, depth)


def test_vertex_normals_and_uvs_reach_the_shading(H, oracle):
    """The smooth-shaded scene really exercises hit.glsl:58-128: dropping the vertex normals, the uvs or the normal maps each
    changes the image, and the product code follows the oracle in every variant (has_normals / has_uvs combinations)."""
    sp = load_sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    W, Hh = 128, 72
    ls = T.LightSamplingConfig()
    images = {}
    for variant in ("full", "no_normals", "no_uvs", "no_normal_maps"):
        s = scenes.smooth_shaded_scene()
        for g in s.geometries:
            if variant == "no_normals":
                g.has_normals = False
            if variant == "no_uvs":
                g.has_uvs = False
        if variant == "no_normal_maps":
            for m in s.materials:
                m.normal_map = -1
        d = s.desc()
        hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
        assert hs
        o = oracle.OracleScene(s)
        ref = o.render_sample(W, Hh, s.camera, sp, 1)
        a = o._args(W, Hh, s.camera, sp)
        img = np.zeros((Hh, W, 4), np.float32)
        H.hostsim_render_sample(hs, C.byref(a), 1, oracle._fp(img))
        H.hostsim_scene_destroy(hs)
        assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), variant
        images[variant] = ref
    for variant in ("no_normals", "no_uvs", "no_normal_maps"):
        assert (images[variant] != images["full"]).any(-1).mean() > 0.02, variant


def random_path_queries(n, seed, box):
    rng = np.random.default_rng(seed)
    q = np.zeros((n, 8), np.float32)
    q[:, 0:3] = rng.uniform(-box, box, (n, 3))
    d = rng.uniform(-0.6 * box, 0.6 * box, (n, 3)) - q[:, 0:3]
    q[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    q[:, 7] = rng.choice([1e20, 2.0 * box, 0.25 * box], n)
    return q


def fold_query_layers(layers):
    """accumulate_query (vulkan/accumulate.glsl:32-42) applied layer after layer, float32."""
    r = None
    for k, x in enumerate(layers):
        accum = r.copy() if k > 0 else np.zeros_like(x)
        accum += (x - accum) / np.float32(k + 1)
        r = accum if k == 0 else r + accum
    return r


def test_ray_queries_through_the_integrator(H, hostsim, oracle):
    """render_ray_queries (vulkan/pt_megakernel.glsl:276-283, 327-334, accumulate.glsl:32-42): the product's shared
    code -- TileMap in query mode, generate_primary + ray override, shade_vertex -- against the oracle's own statement of the
    dispatch (virtual square, 32 x 16 workgroups, swizzled invocation id as the sampler's pixel), for a scene with alpha-tested
    materials (per-candidate shadow seeds use the query's pixel) and emissive triangles."""
    lib = C.CDLL(hostsim)
    lib.hostsim_ray_query_layer.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_void_p, C.c_int32, C.c_uint32, oracle.f32p]
    sp = load_sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    ls = T.LightSamplingConfig()
    for make, box, n in ((lambda: scenes.alpha_tested_soup(6000), 5.0, 3000), (emissive_soup, 4.0, 1537)):
        s = make()
        o = oracle.OracleScene(s)
        d = s.desc()
        hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
        q = random_path_queries(n, 21, box)
        W, Hh = 97, 61  # the frame size only enters through the sampler's linear pixel index
        a = o._args(W, Hh, s.camera, sp, frame_offset=5, first_sample=3)
        layers = []
        for k in range(3):
            out = np.zeros((n, 4), np.float32)
            lib.hostsim_ray_query_layer(hs, C.byref(a), q.ctypes.data, n, k, oracle._fp(out))
            layers.append(out)
        H.hostsim_scene_destroy(hs)
        ref1 = o.render_ray_queries(W, Hh, s.camera, sp, q, view_frame_id=3, frame_offset=5, batch_spp=1)
        assert np.isfinite(ref1).all() and (ref1[:, 3] > 0).mean() > 0.2 and (ref1[:, :3] > 0).any()
        assert np.array_equal(ref1.view(np.uint32), layers[0].view(np.uint32))
        ref3 = o.render_ray_queries(W, Hh, s.camera, sp, q, view_frame_id=3, frame_offset=5, batch_spp=3)
        assert np.array_equal(ref3.view(np.uint32), fold_query_layers(layers).view(np.uint32))
        # a query's sample is NOT the sample of the frame pixel with the same index: the invocation id is swizzled
        assert not np.array_equal(ref1, ref3)


def test_query_pixel_swizzle_is_a_permutation_of_the_dispatch():
    """setup_pixel_assignment.glsl:17-22 maps the invocations of a 32 x 16 workgroup onto its own pixels one-to-one."""
    lib_py = []
    for n in (1, 511, 512, 513, 5000):
        dim_x = int(np.ceil(np.sqrt(np.float32(n))))
        gx_n = (dim_x + 31) // 32
        q = np.arange(((n + 511) // 512) * 512, dtype=np.uint32)
        wg, l = q >> 9, q & 511
        ix, iy = (wg % gx_n) * 32 + (l & 31), (wg // gx_n) * 16 + (l >> 5)
        sx, sy = (ix & ~np.uint32(0x18)) + ((iy & 3) << 3), (iy & ~np.uint32(3)) + ((ix & 0x18) >> 3)
        assert len(set(zip(sx.tolist(), sy.tolist()))) == len(q)
        assert (sx // 32 == ix // 32).all() and (sy // 16 == iy // 16).all()
        lib_py.append((sx, sy))
