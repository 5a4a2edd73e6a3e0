"""The temporal passes of the reference's ENABLE_REALTIME_RESOLVE build: reproject_and_accumulate
(rendering/postprocess/reprojection.glsl, called from vulkan/process_samples.comp:106-113) and the TAA step
(vulkan/processing/process_taa.comp).  CPU part: the product's per-pixel code (csrc/rptr_post.cuh, compiled for the host by
tests/hostsim) against the oracle's statement-by-statement restatement (oracle/post_oracle.h) -- bit for bit, on synthetic
frames that reach every branch -- plus properties the shaders imply.  The CUDA kernels are held to the same oracle in
tests/test_gpu_parity.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


from temporal_util import synthetic_frame  # noqa: E402


@pytest.fixture(scope="module")
def passes(oracle, hostsim):
    L = C.CDLL(hostsim)
    L.hostsim_reproject.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
    L.hostsim_taa.argtypes = [C.c_int32] * 5 + [C.c_void_p] * 4
    return oracle, L


@pytest.mark.parametrize("seed,batch,window", [(1, 1, 8), (2, 4, 32), (3, 1, 1)])
def test_reprojection_product_code_equals_oracle(passes, seed, batch, window):
    oracle, L = passes
    rng = np.random.default_rng(seed)
    w, h = 61, 47
    cur, hist, nd_hist, nd, mj = synthetic_frame(rng, w, h)
    want = oracle.reproject_accumulate(cur, hist, nd_hist, nd, mj, 1.0 / window, batch)
    got = oracle.reproject_accumulate(cur, hist, nd_hist, nd, mj, 1.0 / window, batch, fn=L.hostsim_reproject)
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    stored, shown = want
    # what the pass promises: the stored alpha is 1 - sample weight in [0, 1 - 1/window]; the display colour keeps the sample's alpha
    assert np.array_equal(shown[..., 3], cur[..., 3]) and np.array_equal(shown[..., :3], stored[..., :3])
    wgt = 1.0 - stored[..., 3]
    assert (wgt >= np.float32(1.0 / window) - 1e-7).all() and (wgt <= 1.0).all()
    assert (wgt[cur[..., 3] > 1.0] >= np.float32(0.95)).all()  # 0.95, unless the geometry test rejects the history altogether
    # every branch was reached: history rejected (weight 1), blended, and clamped to the window
    assert (wgt == 1.0).any() and (window == 1 or ((wgt < 1.0) & (wgt > 1.0 / window + 1e-6)).any())
    # pixels whose weight is 1 show their own sample (h + (x - h) * 1: equal up to the rounding of the two operations)
    assert np.allclose(stored[wgt == 1.0][:, :3], cur[wgt == 1.0][:, :3], rtol=1e-6, atol=1e-6)


def test_reprojection_of_a_static_converged_view_is_the_weighted_mean(passes):
    """No motion, identical geometry, history weight 1/k everywhere: the pass must reduce to history + (x - history) * w with
    w = w_old / (1 + w_old * batch) -- the running mean the non-temporal build computes (process_samples.comp:121-125)."""
    oracle, _ = passes
    rng = np.random.default_rng(7)
    w, h = 32, 24
    cur, hist, _, nd, _ = synthetic_frame(rng, w, h)
    cur[..., 3] = 1.0
    nd = np.zeros((h, w, 4), np.float16); nd[..., 2] = 1.0; nd[..., 3] = 4.0   # a wall facing the camera
    mj = np.zeros((h, w, 4), np.float16)
    k = 4
    hist[..., 3] = 1.0 - 1.0 / k
    cur[..., :3] = np.array([0.8, 0.5, 0.25], np.float32)                         # flat colour: the bilateral mix equals the history,
    hist[..., :3] = np.array([0.6, 0.55, 0.3], np.float32)                        # so the projection test leaves the weight alone
    stored, _ = oracle.reproject_accumulate(cur, hist, nd, nd, mj, 1.0 / 64, 1)
    w_old = np.float32(1.0) - hist[..., 3]
    w_new = w_old / (np.float32(1.0) + w_old * np.float32(1.0))
    inner = (slice(1, -1), slice(1, -1))
    assert np.allclose(1.0 - stored[inner][..., 3], w_new[inner], rtol=0, atol=1e-6)
    want = hist[..., :3] + (cur[..., :3] - hist[..., :3]) * w_new[..., None]
    assert np.allclose(stored[inner][..., :3], want[inner], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("seed,upscale", [(11, 1), (12, 2)])
def test_taa_product_code_equals_oracle(passes, seed, upscale):
    oracle, L = passes
    rng = np.random.default_rng(seed)
    rw, rh = 40, 30
    w, h = rw * upscale, rh * upscale
    cur = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    his = np.clip(cur.astype(np.int32) + rng.integers(-40, 41, (h, w, 4)), 0, 255).astype(np.uint8)
    _, _, _, _, mj = synthetic_frame(rng, rw, rh, motion_scale=0.03)
    want = oracle.process_taa(cur, his, mj, upscale)
    got = oracle.process_taa(cur, his, mj, upscale, fn=L.hostsim_taa)
    assert np.array_equal(got, want)
    assert (want != cur).any()
    # reprojected outside the frame: weight 1, the pixel is left as process_samples wrote it
    mjf = mj.astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    rx = (xx + 0.5) / w + 0.5 * mjf[yy // upscale, xx // upscale, 0]
    ry = (yy + 0.5) / h + 0.5 * mjf[yy // upscale, xx // upscale, 1]
    outside = (rx < 0) | (ry < 0) | (rx > 1) | (ry > 1)
    assert outside.any() and np.array_equal(want[outside], cur[outside])


def test_taa_of_a_constant_image_is_the_identity(passes):
    oracle, _ = passes
    w, h = 24, 16
    img = np.full((h, w, 4), 137, np.uint8)
    mj = np.zeros((h, w, 4), np.float16)
    out = oracle.process_taa(img, img, mj, 1)
    inner = (slice(6, -6), slice(6, -6))   # away from the border, where the Lanczos window reads zeros outside the image
    assert np.array_equal(out[inner], img[inner])


# ---- pin: the oracle against the passes EXECUTED from the reference's shader sources (oracle/ref_shim/ref_post.cpp compiles
# rendering/postprocess/reprojection.glsl and vulkan/processing/process_taa.comp where they lie; outputs committed as
# tests/golden/ref_post.npz by `python oracle/gen_golden.py --post-only`) -------------------------------------------------------
@pytest.fixture(scope="module")
def golden_post():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_post.npz"))


@pytest.mark.parametrize("seed,batch,window", [(1, 1, 8), (2, 4, 32), (3, 1, 1)])
def test_oracle_reprojection_matches_the_reference_shader(oracle, golden_post, seed, batch, window):
    rng = np.random.default_rng(seed)
    cur, hist, nd_hist, nd, mj = synthetic_frame(rng, 61, 47)
    stored, shown = oracle.reproject_accumulate(cur, hist, nd_hist, nd, mj, 1.0 / window, batch)
    want_stored, want_shown = golden_post["reproject_%d_stored" % seed], golden_post["reproject_%d_shown" % seed]
    # libm exp / sqrt and unfused dot products on the reference side, RPTR-FP on ours: a few ulp, and no pixel takes another branch
    assert np.abs(stored[..., 3] - want_stored[..., 3]).max() <= 4e-6          # 1 - sample weight
    assert np.abs(stored[..., :3] - want_stored[..., :3]).max() <= 2e-6
    assert np.array_equal(shown[..., 3], want_shown[..., 3]) and np.abs(shown[..., :3] - want_shown[..., :3]).max() <= 2e-6
    assert np.array_equal(stored[..., 3] == 0.0, want_stored[..., 3] == 0.0)    # history rejected <=> weight exactly 1


@pytest.mark.parametrize("seed,upscale", [(11, 1), (12, 2)])
def test_oracle_taa_matches_the_reference_shader(oracle, golden_post, seed, upscale):
    rng = np.random.default_rng(seed)
    rw, rh = 40, 30
    w, h = rw * upscale, rh * upscale
    cur = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    his = np.clip(cur.astype(np.int32) + rng.integers(-40, 41, (h, w, 4)), 0, 255).astype(np.uint8)
    _, _, _, _, mj = synthetic_frame(rng, rw, rh, motion_scale=0.03)
    got = oracle.process_taa(cur, his, mj, upscale).astype(np.int32)
    want = golden_post["taa_%d" % seed].astype(np.int32)
    d = np.abs(got - want)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3   # libm sin vs the RPTR-FP kernel: a value on a rounding boundary may move by one code


def test_reference_executed_passes_reproduce_their_fixture(golden_post):
    """When oracle/_ref is here (the authoring container), the fixture is what the reference's code returns now."""
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    R = C.CDLL(so)
    if not hasattr(R, "ref_reproject_accumulate"):
        pytest.skip("oracle/_ref predates ref_post.cpp")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import pyoracle as po
    R.ref_reproject_accumulate.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(1)
    cur, hist, nd_hist, nd, mj = synthetic_frame(rng, 61, 47)
    stored, shown = po.reproject_accumulate(cur, hist, nd_hist, nd, mj, 1.0 / 8, 1, fn=R.ref_reproject_accumulate)
    assert np.array_equal(stored.view(np.uint32), golden_post["reproject_1_stored"].view(np.uint32))
    assert np.array_equal(shown.view(np.uint32), golden_post["reproject_1_shown"].view(np.uint32))
