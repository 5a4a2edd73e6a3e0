"""1x1-texel mode + stochastic alpha candidate filter (8a-4, generate_candidate_hit,
vulkan/pt_megakernel.glsl:153-272): the product's host material resolution + traversal code (compiled for the CPU by
tests/hostsim) against the oracle, which looks textures up at shading time like the reference does."""
import ctypes as C

import numpy as np
import pytest

from realtimepathtracingresearchframework_b200 import load_pointset_tables, load_sky_fit, scenes, types as T


@pytest.fixture(scope="module")
def H(hostsim, oracle):
    lib = C.CDLL(hostsim)
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_render_sample.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_uint32, oracle.f32p]
    return lib


def render_both(H, oracle, s, W, Hh, sample, sky=None, **kw):
    sp = load_sky_fit(T.SceneConfig(**(sky or {})))
    o = oracle.OracleScene(s)
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    assert hs
    try:
        ref = o.render_sample(W, Hh, s.camera, sp, sample, **kw)
        a = o._args(W, Hh, s.camera, sp, first_sample=sample, **kw)
        img = np.zeros((Hh, W, 4), np.float32)
        H.hostsim_render_sample(hs, C.byref(a), sample, oracle._fp(img))
    finally:
        H.hostsim_scene_destroy(hs)
    return ref, img


def test_texture_handle_encoding():
    import struct
    h = T.texture_handle(5, 2)
    bits = struct.unpack("<I", struct.pack("<f", h))[0]
    assert bits == 0x80000000 | (2 << 29) | 5  # rendering/bsdfs/texture_channel_mask.h:20-27
    assert h < 0 or bits >> 31 == 1


@pytest.mark.parametrize("sample,frame_offset", [(0, 0), (3, 7)])
def test_alpha_tested_scene_matches_oracle_bit_for_bit(H, oracle, sample, frame_offset):
    s = scenes.alpha_tested_soup()
    ref, img = render_both(H, oracle, s, 200, 120, sample, sky=dict(sun_dir=(0.35, 0.8, 0.45)), frame_offset=frame_offset)
    assert np.isfinite(ref).all() and ref[..., :3].max() > 0
    assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), "%d pixels differ" % (ref != img).any(-1).sum()


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_alpha_test_with_the_qmc_pointsets_draws_from_its_own_lcg(H, oracle, variant):
    """RBO_rng_variant != UNIFORM: alpha_rng is a separate LCG seeded like the UNIFORM pointset (pt_megakernel.glsl:354-358),
    so the path's BN / Sobol dimensions do not shift when a candidate is alpha-tested."""
    s = scenes.alpha_tested_soup()
    kw = dict(frame_offset=3, rng_variant=variant, pointset_tables=load_pointset_tables())
    for sample in (0, 2):
        ref, img = render_both(H, oracle, s, 300, 140, sample, sky=dict(sun_dir=(0.35, 0.8, 0.45)), **kw)
        assert np.isfinite(ref).all() and ref[..., :3].max() > 0
        assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), "%d pixels differ" % (ref != img).any(-1).sum()
    # and the pointset is really in effect
    uni, _ = render_both(H, oracle, s, 300, 140, 2, sky=dict(sun_dir=(0.35, 0.8, 0.45)), frame_offset=3)
    assert not np.array_equal(ref[..., :3], uni[..., :3])


def test_alpha_changes_the_image_the_expected_way(H, oracle):
    """cut-out triangles vanish, half-transparent ones let about half of the primary rays through"""
    s = scenes.alpha_tested_soup()
    opaque = scenes.alpha_tested_soup()
    for m in opaque.materials:
        m.flags |= T.BASE_MATERIAL_NOALPHA
    W, Hh = 200, 120
    a, _ = render_both(H, oracle, s, W, Hh, 0)
    b, _ = render_both(H, oracle, opaque, W, Hh, 0)
    assert not np.array_equal(a, b)
    # alpha channel of the sample = "the path hit something" (pt_megakernel.glsl:736): fewer hits with cut-outs
    assert a[..., 3].sum() < b[..., 3].sum()


def test_textures_larger_than_one_texel_are_rejected(H, oracle):
    s = scenes.alpha_tested_soup(500)
    s.textures[0] = (np.zeros((2, 2, 4), np.uint8), T.COLOR_SPACE_SRGB)
    ls = T.LightSamplingConfig()
    d = s.desc()
    assert not H.hostsim_scene_create(C.byref(d), C.byref(ls))  # build_host_scene throws: "only 1 x 1 textures are supported"


def test_srgb_and_unorm_decoding_of_resolved_materials(H, oracle):
    """AOV channel 1 (base colour * throughput at the first hit) shows the resolved texel colours exactly."""
    s = scenes.alpha_tested_soup(3000)
    p = T.RenderParams()
    p.output_channel = 1
    ref, img = render_both(H, oracle, s, 160, 90, 0, params=p)
    assert np.array_equal(ref.view(np.uint32), img.view(np.uint32))

    def srgb(v):
        c = v / 255.0
        return np.float32(c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4)
    want = np.array([srgb(188), srgb(64), srgb(230)], np.float32)  # material 3: opaque sRGB texel, alpha 1
    px = ref[..., :3].reshape(-1, 3)
    assert (np.abs(px - want).max(axis=1) == 0).any(), "no pixel shows the decoded sRGB texel"


def test_normal_map_changes_shading_only_where_it_is_set(H, oracle):
    """one-texel normal maps (pt_megakernel.glsl:634-654): bit-exact against the oracle (covered by the tests above, the soup
    has two such materials) and really in effect: removing them changes pixels, but not the hit mask."""
    s = scenes.alpha_tested_soup()
    flat = scenes.alpha_tested_soup()
    for m in flat.materials:
        m.normal_map = -1
    W, Hh = 200, 120
    a, ia = render_both(H, oracle, s, W, Hh, 0)
    b, ib = render_both(H, oracle, flat, W, Hh, 0)
    assert np.array_equal(a.view(np.uint32), ia.view(np.uint32)) and np.array_equal(b.view(np.uint32), ib.view(np.uint32))
    assert not np.array_equal(a[..., :3], b[..., :3])
    assert np.array_equal(a[..., 3], b[..., 3])
    # AOV channel 2 = shading normal of the first hit: differs exactly where a normal-mapped material is hit
    p = T.RenderParams()
    p.output_channel = 2
    na, _ = render_both(H, oracle, s, W, Hh, 0, params=p)
    nb, _ = render_both(H, oracle, flat, W, Hh, 0, params=p)
    changed = (na[..., :3] != nb[..., :3]).any(-1)
    assert 0.02 < changed.mean() < 0.5


# ---- textures larger than 1 x 1 ------------------------------------------------------------------------------
def test_texture_unit_product_equals_oracle(H, hostsim, oracle):
    """The bilinear REPEAT sampler (our statement of the VkSampler the reference leaves to the hardware): product
    (csrc/rptr_shading.cuh sample_texture, RGBA8 device image) and oracle (TextureSet::sample on the caller's texels) bit for
    bit over wrap-around, negative and huge coordinates, texel centres, 1-4 channels, sRGB and linear."""
    lib = C.CDLL(hostsim)
    lib.hostsim_sample_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, oracle.f32p]
    L = oracle.lib()
    L.oracle_sample_texture.argtypes = [C.POINTER(T.TextureDesc), C.c_float, C.c_float, oracle.f32p]
    s = scenes.textured_scene()
    d = s.desc()
    ls = T.LightSamplingConfig()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    assert hs
    rng = np.random.default_rng(4)
    uvs = np.concatenate([rng.uniform(-3, 9, (300, 2)), rng.uniform(0, 1, (200, 2)), [[0, 0], [1, 1], [0.5, 0.5], [-0.0, 1e-9], [1e12, -1e12], [np.nan, 0.3],
                                                                                       [0.5 / 64, 0.5 / 32], [1 - 1e-7, 1 - 1e-7]]]).astype(np.float32)
    sizes = set()
    for tid, (px, _) in enumerate(s.textures):
        if px.shape[0] * px.shape[1] == 1:
            continue  # folded into the materials by the product
        sizes.add(px.shape)
        for u, v in uvs:
            a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
            lib.hostsim_sample_texture(hs, tid, float(u), float(v), a.ctypes.data_as(oracle.f32p))
            L.oracle_sample_texture(C.byref(d.textures[tid]), float(u), float(v), b.ctypes.data_as(oracle.f32p))
            assert a.tobytes() == b.tobytes(), (tid, u, v, a, b)
            assert np.isfinite(a).all() and (a >= 0).all() and (a <= 1).all()
    H.hostsim_scene_destroy(hs)
    assert len(sizes) >= 4
    # texel centres return the texel itself, sRGB through the transfer function; the sample between two texels is their mean
    img = np.array([[[0, 0, 0, 255], [255, 255, 255, 0]]], np.uint8)
    td = T.TextureDesc(width=2, height=1, channels=4, color_space=T.COLOR_SPACE_SRGB)
    td.texels = img.ctypes.data_as(C.POINTER(C.c_uint8))
    out = np.zeros(4, np.float32)
    L.oracle_sample_texture(C.byref(td), 0.25, 0.5, out.ctypes.data_as(oracle.f32p))
    assert out.tolist() == [0.0, 0.0, 0.0, 1.0]
    L.oracle_sample_texture(C.byref(td), 0.5, 0.5, out.ctypes.data_as(oracle.f32p))
    assert out.tolist() == [0.5, 0.5, 0.5, 0.5]
    L.oracle_sample_texture(C.byref(td), 1.0, 0.5, out.ctypes.data_as(oracle.f32p))  # REPEAT: half way between texel 1 and texel 0
    assert out.tolist() == [0.5, 0.5, 0.5, 0.5]


@pytest.mark.parametrize("sample,transmission", [(0, 0), (2, 1)])
def test_textured_scene_matches_oracle_bit_for_bit(H, oracle, sample, transmission):
    """uv lookups of base colour + alpha, specular / roughness / metallic channels, ior and the normal map at the hit, and of the
    alpha channel at traversal candidates (closest hit: the path's LCG, front to back; shadow rays: per-candidate seeds)."""
    s = scenes.textured_scene()
    ref, img = render_both(H, oracle, s, 192, 108, sample, sky=dict(sun_dir=(0.35, 0.8, 0.45)), frame_offset=2, transmission=transmission)
    assert np.isfinite(ref).all() and ref[..., :3].max() > 0
    assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), "%d pixels differ" % (ref != img).any(-1).sum()
    if sample == 0:
        # the textures matter: constant materials give another image, and so do opaque ones (the alpha channel cuts holes)
        flat = scenes.smooth_shaded_scene()
        ref_flat, _ = render_both(H, oracle, flat, 192, 108, sample, sky=dict(sun_dir=(0.35, 0.8, 0.45)), frame_offset=2)
        assert (ref_flat != ref).any(-1).mean() > 0.05
        opaque = scenes.textured_scene()
        for m in opaque.materials:
            m.flags |= T.BASE_MATERIAL_NOALPHA
        ref_opaque, _ = render_both(H, oracle, opaque, 192, 108, sample, sky=dict(sun_dir=(0.35, 0.8, 0.45)), frame_offset=2)
        assert (ref_opaque != ref).any(-1).mean() > 0.01


def test_textured_emitters_and_missing_uvs_are_refused(H):
    s = scenes.textured_scene()
    s.materials[1].emission_intensity = 3.0
    s.materials[1].base_color = s.materials[0].base_color
    d = s.desc()
    ls = T.LightSamplingConfig()
    assert not H.hostsim_scene_create(C.byref(d), C.byref(ls))  # the reference's own emitter collection reads the handle bits as a colour
    s = scenes.textured_scene()
    s.geometries[0].has_uvs = False  # alpha texture on a geometry without uvs
    d = s.desc()
    assert not H.hostsim_scene_create(C.byref(d), C.byref(ls))
