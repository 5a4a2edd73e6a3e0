"""Low-discrepancy samplers (RenderBackendOptions::rng_variant = BN / SOBOL / Z_SBL).

tests/golden/ref_pointsets.npz holds streams produced by EXECUTING the reference's own
rendering/pointsets/{sobol,sample_order,bn_rng}.glsl (oracle/gen_golden.py through oracle/_ref).  Checked bit for bit:
the oracle's restatement (oracle/pointsets_oracle.h), the product's rptr_pointsets.cuh compiled for the CPU
(tests/hostsim), and -- where oracle/_ref is present -- the reference again, live.  Whole-image parity of the product code
against the oracle with each sampler follows."""
import ctypes as C
import os

import numpy as np
import pytest

from realtimepathtracingresearchframework_b200 import load_pointset_tables, load_sky_fit, scenes, types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u32p, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_pointsets.npz"))


@pytest.fixture(scope="module")
def tables():
    return load_pointset_tables()


@pytest.fixture(scope="module")
def H(hostsim, oracle):
    lib = C.CDLL(hostsim)
    lib.hostsim_pointset_replay.argtypes = [C.c_int, C.POINTER(C.c_void_p)] + [C.c_uint32] * 6 + [i32p, i32p, C.c_int, oracle.f32p, u32p]
    lib.hostsim_morton_sample_id.restype = C.c_uint32
    lib.hostsim_morton_sample_id.argtypes = [C.c_uint32] * 5 + [C.c_int, C.c_int]
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(T.LightSamplingConfig)]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_render_sample.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_uint32, oracle.f32p]
    return lib


def replay(fn, variant, tables, oracle, gold, with_h=False):
    ops = np.ascontiguousarray(gold["ops"], np.int32)
    args = np.ascontiguousarray(gold["args"], np.int32)
    q = gold["v%d_in" % variant]
    n_draws = int((ops == 0).sum())
    draws = np.zeros((len(q), n_draws), np.float32)
    state = np.zeros((len(q), 2), np.uint32)
    ptrs, keep = oracle.table_ptrs(tables)
    for i, row in enumerate(q):
        pre = [variant, ptrs] if ptrs is not None else [variant]
        a = pre + [int(row[0]), int(row[1]), int(row[2]), int(row[3]), int(row[4]), int(row[5])]
        if with_h:
            a.append(1080)
        m = fn(*a, ops.ctypes.data_as(i32p), args.ctypes.data_as(i32p), len(ops), draws[i].ctypes.data_as(oracle.f32p),
               state[i].ctypes.data_as(u32p))
        assert m == n_draws
    return draws, state


def test_tables_have_the_reference_shapes(tables):
    assert [t.size for t in tables] == [1024 * 32, 256 * 256, 256 * 256, 128 * 128 * 8]
    assert all(t.dtype == np.uint32 for t in tables)
    # the inversion table is a permutation of the 65536 samples of one tile (sobol.glsl:113-133)
    assert np.array_equal(np.sort(tables[1]), np.arange(65536, dtype=np.uint32))
    assert tables[2].max() < 256 and tables[3].max() < 256


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_oracle_samplers_match_the_reference_streams(oracle, gold, tables, variant):
    draws, state = replay(oracle.lib().oracle_pointset_replay, variant, tables, oracle, gold)
    assert np.array_equal(state, gold["v%d_state" % variant])
    assert np.array_equal(draws.view(np.uint32), gold["v%d_draws" % variant].view(np.uint32))
    assert (draws >= 0).all() and (draws <= 1).all()


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_product_samplers_match_the_reference_streams(H, oracle, gold, tables, variant):
    draws, state = replay(H.hostsim_pointset_replay, variant, tables, oracle, gold)
    assert np.array_equal(state, gold["v%d_state" % variant])
    assert np.array_equal(draws.view(np.uint32), gold["v%d_draws" % variant].view(np.uint32))


def test_morton_sample_id_matches_the_reference(H, oracle, gold):
    mq = gold["morton_in"]
    for fn in (oracle.lib().oracle_morton_sample_id, H.hostsim_morton_sample_id):
        got = np.array([fn(*[int(x) for x in row]) for row in mq], np.uint32)
        assert np.array_equal(got, gold["morton_out"])


def test_fixture_is_what_the_reference_produces_now(oracle, gold, tables):
    """Only where oracle/_ref was built from /root/reference: the committed fixture and tables are reproduced live."""
    R = oracle.ref()
    if R is None or not hasattr(R, "ref_pointset_replay"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for which, t in enumerate(tables):
        buf = np.zeros(t.size, np.uint32)
        assert R.ref_pointset_table(which, buf.ctypes.data_as(u32p)) == t.size
        assert np.array_equal(buf, t)

    def fn(variant, _ptrs, *rest):
        return R.ref_pointset_replay(variant, *rest)
    for variant in (1, 2, 3):
        ops = np.ascontiguousarray(gold["ops"], np.int32)
        args = np.ascontiguousarray(gold["args"], np.int32)
        q = gold["v%d_in" % variant][:64]
        out = np.zeros(int((ops == 0).sum()), np.float32)
        for i, row in enumerate(q):
            R.ref_pointset_replay(variant, *[int(x) for x in row], 1080, ops.ctypes.data_as(i32p), args.ctypes.data_as(i32p), len(ops),
                                  out.ctypes.data_as(oracle.f32p), None)
            assert np.array_equal(out.view(np.uint32), gold["v%d_draws" % variant][i].view(np.uint32))


def test_sobol_points_are_a_scrambled_0_2_sequence(oracle, tables):
    """Property, not fixture: with the XOR scramble undone, dimensions (0, 1) of samples 0..255 of one pixel stratify into
    a 16 x 16 grid (Sobol' (0,2)-sequence) -- guards the table layout matrix[dim * 32 + bit]."""
    m = tables[0].reshape(1024, 32)
    pts = np.zeros((256, 2), np.uint32)
    for i in range(256):
        for d in range(2):
            r = 0
            for b in range(8):
                if (i >> b) & 1:
                    r ^= int(m[d, b])
            pts[i, d] = r
    cells = (pts[:, 0] >> 28) * 16 + (pts[:, 1] >> 28)
    assert len(set(cells.tolist())) == 256


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_product_code_with_each_sampler_matches_oracle_bit_for_bit(H, oracle, tables, variant):
    s = scenes.random_triangles(6000)
    sp = load_sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    o = oracle.OracleScene(s)
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    assert hs
    try:
        W, Hh = 300, 170  # wider than one 256-pixel Sobol tile, taller than one 128-pixel blue-noise tile
        uniform = o.render_sample(W, Hh, s.camera, sp, 2, frame_offset=5)
        for sample in (0, 2):
            kw = dict(frame_offset=5, rng_variant=variant, pointset_tables=tables)
            ref = o.render_sample(W, Hh, s.camera, sp, sample, **kw)
            a = o._args(W, Hh, s.camera, sp, first_sample=sample, **kw)  # BN seeds from view_params.frame_id
            img = np.zeros((Hh, W, 4), np.float32)
            H.hostsim_render_sample(hs, C.byref(a), sample, oracle._fp(img))
            assert np.isfinite(ref).all() and ref[..., :3].max() > 0
            assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), "%d pixels differ" % (ref != img).any(-1).sum()
        assert not np.array_equal(ref, uniform), "the sampler must change the image"
        # same estimator: the image means agree to Monte-Carlo accuracy
        assert abs(ref[..., :3].mean() / uniform[..., :3].mean() - 1.0) < 0.15
    finally:
        H.hostsim_scene_destroy(hs)


# ---- raster-TAA screen jitter (render_vulkan.cpp:2917-2926; librender/halton.h) -------------------------------------------
def test_halton_table_and_screen_jitter(H, oracle, gold):
    hal = gold["halton_23"]  # the reference's own table, all 64 entries
    assert hal.shape == (64, 2) and hal[0, 0] == np.float32(0.5)
    H.hostsim_halton_23.argtypes = [C.c_int32, oracle.f32p]
    H.hostsim_screen_jitter.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, oracle.f32p]
    for fn in (oracle.lib().oracle_halton_23, H.hostsim_halton_23):
        got = np.zeros((64, 2), np.float32)
        for k in range(64):
            fn(k, got[k].ctypes.data_as(oracle.f32p))
        assert np.array_equal(got.view(np.uint32), hal.view(np.uint32))
    # jitter = halton_23[(frame_offset + frame_id) % 16] * 2 / dims - 1 / dims, in float, left to right
    for fo, fid, w, h in ((0, 0, 1920, 1080), (7, 12, 333, 77), (2 ** 32 - 3, 9, 640, 480)):
        k = ((fo + fid) & 0xFFFFFFFF) % 16
        want = hal[k] * np.float32(2.0) / np.array([w, h], np.float32) - np.float32(1.0) / np.array([w, h], np.float32)
        for fn in (oracle.lib().oracle_screen_jitter, H.hostsim_screen_jitter):
            got = np.zeros(2, np.float32)
            fn(fo, fid, w, h, got.ctypes.data_as(oracle.f32p))
            assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))
    # and against the statements of update_view_parameters themselves, executed from the reference's host code
    for (fo, fid, w, h), want in zip(gold["jitter_in"].tolist(), gold["jitter_out"]):
        for fn in (oracle.lib().oracle_screen_jitter, H.hostsim_screen_jitter):
            got = np.zeros(2, np.float32)
            fn(fo, fid, w, h, got.ctypes.data_as(oracle.f32p))
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (fo, fid, w, h)


def test_raster_taa_frames_match_oracle(H, oracle):
    """enable_raster_taa: no pixel-filter draws, the frame's Halton jitter instead (pt_megakernel.glsl:316-320)"""
    s = scenes.random_triangles(5000)
    sp = load_sky_fit()
    o = oracle.OracleScene(s)
    ls = T.LightSamplingConfig()
    d = s.desc()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    W, Hh = 160, 90
    p = T.RenderParams(enable_raster_taa=1)
    imgs = []
    for sample in (0, 5):
        ref = o.render_sample(W, Hh, s.camera, sp, sample, params=p, frame_offset=3)
        a = o._args(W, Hh, s.camera, sp, params=p, frame_offset=3, first_sample=sample)
        img = np.zeros((Hh, W, 4), np.float32)
        H.hostsim_render_sample(hs, C.byref(a), sample, oracle._fp(img))
        assert np.array_equal(ref.view(np.uint32), img.view(np.uint32))
        imgs.append(ref)
    H.hostsim_scene_destroy(hs)
    assert not np.array_equal(imgs[0], imgs[1])
    # every pixel of a frame looks through the same sub-pixel offset: the hit mask of frame 0 equals the one rendered with the
    # jitter folded into a pinhole without any draw -- here simply: it differs from the box-filtered frame
    assert not np.array_equal(imgs[0], o.render_sample(W, Hh, s.camera, sp, 0, frame_offset=3))
