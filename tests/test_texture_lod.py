"""f2: mip chains, block-compressed textures and the footprint level of detail (USE_MIPMAPPING, librender/render_params.glsl.h:8).
* footprint algebra (rendering/rt/footprint.glsl:10-61): product == oracle bit for bit, oracle pinned to the file itself executed as
  C++ (oracle/ref_shim/ref_footprint.cpp -> tests/golden/ref_textures.npz);
* unpack_material / get_material_alpha with textureGrad (rendering/rt/material_textures.glsl:37-135): the reference's code executed
  over our texture unit (ref_shim/ref_materials.cpp) == the oracle;
* the texture unit itself (ours: the reference leaves it to the sampler hardware): textureGrad / textureLod product == oracle, plus
  properties; BC1 / BC3 / BC5 decoding product == oracle, known blocks, round trips;
* a textured, mip-mapped, block-compressed scene: product code (tests/hostsim) == oracle on whole frames."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import texture_util as tu  # noqa: E402
from realtimepathtracingresearchframework_b200 import load_sky_fit, scenes, types as T  # noqa: E402

f32p = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def libs(oracle, hostsim):
    L = oracle.lib()
    H = C.CDLL(hostsim)
    for lib, pre in ((L, "oracle"), (H, "hostsim")):
        getattr(lib, pre + "_footprint_op").argtypes = [C.c_int32, f32p, f32p]
        getattr(lib, pre + "_log2").restype = C.c_float
        getattr(lib, pre + "_log2").argtypes = [C.c_float]
        getattr(lib, pre + "_decode_texture").restype = C.c_int64
        getattr(lib, pre + "_decode_texture").argtypes = [C.POINTER(T.TextureDesc), C.c_void_p, C.c_int64]
    L.oracle_sample_texture_grad.argtypes = [C.POINTER(T.TextureDesc), C.c_float, C.c_float, f32p, f32p, f32p]
    L.oracle_sample_texture_lod.argtypes = [C.POINTER(T.TextureDesc), C.c_float, C.c_float, C.c_int32, f32p]
    L.oracle_unpack_material_at.argtypes = [C.POINTER(T.BaseMaterial), C.POINTER(T.TextureDesc), C.c_int, f32p, f32p, f32p]
    H.hostsim_scene_create.restype = C.c_void_p
    H.hostsim_scene_create.argtypes = [C.c_void_p, C.c_void_p]
    H.hostsim_scene_destroy.argtypes = [C.c_void_p]
    H.hostsim_sample_texture_grad.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, f32p, f32p, f32p]
    H.hostsim_sample_texture_lod.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_int32, f32p]
    return L, H


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_textures.npz"))


def run_op(lib, name, op, inp, n_out):
    out = np.zeros(n_out, np.float32)
    getattr(lib, name)(op, np.ascontiguousarray(inp, np.float32).ctypes.data_as(f32p), out.ctypes.data_as(f32p))
    return out


def test_footprint_algebra_product_oracle_reference(libs, golden):
    L, H = libs
    op0, d2 = tu.footprint_cases(1500, 99)
    worst = np.zeros(3)
    for i in range(len(op0)):
        F = run_op(L, "oracle_footprint_op", 0, op0[i], 4)
        assert F.tobytes() == run_op(H, "hostsim_footprint_op", 0, op0[i], 4).tobytes()
        # the chain continues from the reference's own intermediate results, so each operation is compared on identical inputs
        in1 = np.concatenate([d2[i], op0[i, :3], golden["fp_to_footprint"][i]])
        F2 = run_op(L, "oracle_footprint_op", 1, in1, 4)
        assert F2.tobytes() == run_op(H, "hostsim_footprint_op", 1, in1, 4).tobytes()
        in2 = np.concatenate([d2[i], golden["fp_reflect"][i]])
        dp = run_op(L, "oracle_footprint_op", 2, in2, 6)
        assert dp.tobytes() == run_op(H, "hostsim_footprint_op", 2, in2, 6).tobytes()
        for k, (got, want) in enumerate(((F, golden["fp_to_footprint"][i]), (F2, golden["fp_reflect"][i]), (dp, golden["fp_to_dpdxy"][i]))):
            if np.isfinite(want).all():
                assert np.isfinite(got).all()
                worst[k] = max(worst[k], np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
    # glm-style unfused products there, RPTR-FP fma chains here; the eigen-decomposition of footprint_to_dpdxy is the least conditioned
    assert worst[0] <= 2e-6 and worst[1] <= 8e-6 and worst[2] <= 2e-4, worst


def test_log2_of_the_contract(libs):
    L, H = libs
    xs = np.concatenate([2.0 ** np.arange(-20, 21), np.random.default_rng(1).uniform(1e-6, 1e6, 4000), [1.0, np.sqrt(2), 0.7071, 3.0e-39, 0.0]]).astype(np.float32)
    for x in xs:
        a, b = L.oracle_log2(float(x)), H.hostsim_log2(float(x))
        assert np.float32(a).tobytes() == np.float32(b).tobytes()
        if x >= 1.2e-38:
            assert abs(a - np.log2(np.float64(x))) <= 3e-7 * max(1.0, abs(np.log2(np.float64(x))))
        else:
            assert a == -127.0   # zero / subnormal: below every level
    assert L.oracle_log2(8.0) == 3.0 and L.oracle_log2(0.25) == -2.0


def decode(lib, name, desc):
    n = getattr(lib, name)(C.byref(desc), None, 0)
    assert n > 0
    out = np.zeros(n, np.uint8)
    assert getattr(lib, name)(C.byref(desc), out.ctypes.data, n) == n
    return out


def test_block_compressed_textures_decode_identically_and_plausibly(libs):
    L, H = libs
    tset = tu.texture_set()
    descs, keep = tu.texture_descs(tset)
    for i, (levels, cs, bc) in enumerate(tset):
        a, b = decode(L, "oracle_decode_texture", descs[i]), decode(H, "hostsim_decode_texture", descs[i])
        assert np.array_equal(a, b), "texture %d (bc %d)" % (i, bc)
        assert len(a) == 4 * sum(l.shape[0] * l.shape[1] for l in levels)   # every level at its size, padding of the blocks dropped
        if bc == 0:
            want = []
            for l in levels:
                full = np.zeros(l.shape[:2] + (4,), np.uint8); full[..., 3] = 255; full[..., :l.shape[2]] = l
                want.append(full.reshape(-1))
            assert np.array_equal(a, np.concatenate(want))
    # a smooth image survives each format with the error block compression is expected to have (the decoders mean what the encoder meant)
    yy, xx = np.mgrid[0:16, 0:24]
    smooth = np.stack([xx * 10, yy * 15, 255 - xx * 9, np.where(xx < 12, 255, 40 + yy * 8)], -1).astype(np.uint8)
    for bc in (1, -1, 3, 5):
        d = T.TextureDesc(width=24, height=16, channels=4, color_space=0, bc_format=bc, mip_levels=1)
        buf = scenes.encode_bc(smooth, bc).copy()
        d.texels = buf.ctypes.data_as(C.POINTER(C.c_uint8))
        got = decode(H, "hostsim_decode_texture", d).reshape(16, 24, 4).astype(np.int32)
        assert np.array_equal(got.reshape(-1), decode(L, "oracle_decode_texture", d))
        want = smooth.astype(np.int32)
        chans = {1: [0, 1, 2], -1: [0, 1, 2], 3: [0, 1, 2, 3], 5: [0, 1]}[bc]
        err = np.abs(got[..., chans] - want[..., chans])
        if bc == -1:
            err = err[want[..., 3] >= 128]   # punched-out texels decode to transparent black
        assert err.max() <= 32 and err.mean() <= 10, (bc, err.max(), err.mean())   # two gradients per block against one colour line
        if bc in (1, 5):
            assert (got[..., 3] == 255).all()
        if bc == 5:
            assert (got[..., 2] == 0).all()
        if bc == -1:   # punch-through alpha: 0 or 255 only, following the source's alpha
            assert set(np.unique(got[..., 3])) <= {0, 255} and ((got[..., 3] == 0) == (want[..., 3] < 128)).all()
    # known blocks (Khronos Data Format Specification, S3TC / RGTC): red / blue endpoints, the four indices in the first four texels
    blk = bytes([0x00, 0xF8, 0x1F, 0x00]) + (0b11100100).to_bytes(4, "little")   # c0 = 0xF800 > c1 = 0x001F: four-colour mode
    d = T.TextureDesc(width=4, height=4, channels=4, color_space=0, bc_format=1, mip_levels=1)
    buf = np.frombuffer(blk, np.uint8).copy()
    d.texels = buf.ctypes.data_as(C.POINTER(C.c_uint8))
    px = decode(H, "hostsim_decode_texture", d).reshape(4, 4, 4)
    assert px[0, 0].tolist() == [255, 0, 0, 255] and px[0, 1].tolist() == [0, 0, 255, 255]
    assert px[0, 2].tolist() == [170, 0, 85, 255] and px[0, 3].tolist() == [85, 0, 170, 255]
    blk = bytes([0x1F, 0x00, 0x00, 0xF8]) + (0b11100100).to_bytes(4, "little")   # c0 < c1: three colours + transparent (BC1 RGBA) / black (BC1 RGB)
    buf = np.frombuffer(blk, np.uint8).copy()
    d.texels = buf.ctypes.data_as(C.POINTER(C.c_uint8))
    for fmt, last in ((-1, [0, 0, 0, 0]), (1, [0, 0, 0, 255])):
        d.bc_format = fmt
        px = decode(H, "hostsim_decode_texture", d).reshape(4, 4, 4)
        assert px[0, 2].tolist() == [128, 0, 128, 255] and px[0, 3].tolist() == last
        assert np.array_equal(px.reshape(-1), decode(L, "oracle_decode_texture", d))
    a8 = bytes([255, 0]) + int(sum(k << (3 * k) for k in range(8))).to_bytes(6, "little")    # BC4: indices 0..7 in the first eight texels
    blk = a8 + bytes([0x00, 0xF8, 0x1F, 0x00, 0, 0, 0, 0])
    buf = np.frombuffer(blk, np.uint8).copy()
    d.texels = buf.ctypes.data_as(C.POINTER(C.c_uint8))
    d.bc_format = 3
    px = decode(H, "hostsim_decode_texture", d).reshape(16, 4)
    assert px[:8, 3].tolist() == [255, 0, 219, 182, 146, 109, 73, 36] and px[0, :3].tolist() == [255, 0, 0]
    assert np.array_equal(px.reshape(-1), decode(L, "oracle_decode_texture", d))
    d.bc_format = 2   # BC2 is not a format .vks scenes use: refused, not mis-decoded
    assert H.hostsim_decode_texture(C.byref(d), None, 0) < 0 and L.oracle_decode_texture(C.byref(d), None, 0) < 0


def textured_mip_scene(bc=True):
    """The textured test scene with mip chains on every image and the block formats of a .vks scene (colour BC1 RGBA / BC3, ORM BC1,
    normal map BC5)."""
    s = scenes.textured_scene()
    fmts = {0: 3, 1: 1, 2: 5} if bc else {}
    for i, t in enumerate(s.textures):
        px = t[0]
        if px.shape[0] * px.shape[1] > 1:
            s.textures[i] = (px, t[1], scenes.mip_chain(px), fmts.get(i, 0))
    return s


def test_texture_unit_product_equals_oracle_and_behaves(libs):
    L, H = libs
    s = textured_mip_scene()
    d = s.desc()
    ls = T.LightSamplingConfig()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    assert hs
    rng = np.random.default_rng(8)
    checked = 0
    for tid in range(d.n_textures):
        td = d.textures[tid]
        if td.width * td.height == 1:
            continue
        for _ in range(150):
            u, v = (float(x) for x in rng.uniform(-2, 3, 2))
            sc = 10.0 ** rng.uniform(-4, 0)
            dx, dy = (rng.normal(size=2) * sc).astype(np.float32), (rng.normal(size=2) * sc * rng.uniform(0.02, 1)).astype(np.float32)
            a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
            H.hostsim_sample_texture_grad(hs, tid, u, v, dx.ctypes.data_as(f32p), dy.ctypes.data_as(f32p), a.ctypes.data_as(f32p))
            L.oracle_sample_texture_grad(C.byref(td), u, v, dx.ctypes.data_as(f32p), dy.ctypes.data_as(f32p), b.ctypes.data_as(f32p))
            assert a.tobytes() == b.tobytes(), (tid, u, v, dx, dy)
            assert np.isfinite(a).all() and (a >= -1e-6).all() and (a <= 1 + 1e-6).all()
            lvl = int(rng.integers(0, 8))
            H.hostsim_sample_texture_lod(hs, tid, u, v, lvl, a.ctypes.data_as(f32p))
            L.oracle_sample_texture_lod(C.byref(td), u, v, lvl, b.ctypes.data_as(f32p))
            assert a.tobytes() == b.tobytes()
            checked += 1
        # zero footprint = the base level with one tap; a footprint larger than the image = the 1 x 1 level (the image's mean, about)
        z = np.zeros(2, np.float32)
        big = np.array([4.0, 0.0], np.float32), np.array([0.0, 4.0], np.float32)
        base, grad0, top, gradbig = (np.zeros(4, np.float32) for _ in range(4))
        L.oracle_sample_texture_lod(C.byref(td), 0.3, 0.6, 0, base.ctypes.data_as(f32p))
        L.oracle_sample_texture_grad(C.byref(td), 0.3, 0.6, z.ctypes.data_as(f32p), z.ctypes.data_as(f32p), grad0.ctypes.data_as(f32p))
        assert base.tobytes() == grad0.tobytes()
        L.oracle_sample_texture_lod(C.byref(td), 0.3, 0.6, 99, top.ctypes.data_as(f32p))
        L.oracle_sample_texture_grad(C.byref(td), 0.3, 0.6, big[0].ctypes.data_as(f32p), big[1].ctypes.data_as(f32p), gradbig.ctypes.data_as(f32p))
        assert top.tobytes() == gradbig.tobytes()
    H.hostsim_scene_destroy(hs)
    assert checked >= 400


def test_reference_material_glue_over_our_texture_unit(libs, golden):
    """unpack_material + get_material_alpha as the reference's own code computes them when every textureGrad goes to our texture unit
    (fixture: oracle/gen_golden.py --textures-only) == the oracle's restatement, on raw and block-compressed mip-mapped textures."""
    L, _ = libs
    tset = tu.texture_set()
    descs, keep = tu.texture_descs(tset)
    mats, uv, duvdxy = tu.random_textured_materials(300, len(tset), 123)
    want = golden["mat_grad"]
    for i, m in enumerate(mats):
        got = np.zeros(17, np.float32)
        L.oracle_unpack_material_at(C.byref(m), descs, len(tset), uv[i].ctypes.data_as(f32p), duvdxy[i].ctypes.data_as(f32p), got.ctypes.data_as(f32p))
        assert np.allclose(got, want[i], rtol=2e-6, atol=1e-7), (i, got, want[i])
    assert (want[:, 16] < 1.0).any() and (want[:, 7] > 0).any() and np.unique(np.round(want[:, 0], 3)).size > 50


@pytest.mark.parametrize("bc", [False, True])
def test_textured_mip_scene_product_code_equals_oracle(libs, oracle, bc):
    """Whole frames of the mip-mapped (and block-compressed) textured scene: footprints from the camera, their reflection at every
    bounce, duvdxy at the hits, anisotropic textureGrad on every textured parameter, the normal map at level = bounce, alpha at
    candidates from the base level -- product code == oracle bit for bit."""
    L, H = libs
    H.hostsim_render_sample.argtypes = [C.c_void_p, C.POINTER(oracle.OracleRenderArgs), C.c_uint32, oracle.f32p]
    s = textured_mip_scene(bc)
    W, Hh = 80, 48
    sp = load_sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    o = oracle.OracleScene(s)
    d = s.desc()
    ls = T.LightSamplingConfig()
    hs = H.hostsim_scene_create(C.byref(d), C.byref(ls))
    assert hs
    last = None
    for pr, sample in ((1.0, 0), (3.0, 2)):
        p = T.RenderParams(pixel_radius=pr)
        ref = o.render_sample(W, Hh, s.camera, sp, sample, transmission=1, params=p)
        a = o._args(W, Hh, s.camera, sp, first_sample=sample, transmission=1, params=p)   # frame_id seeds the shadow-ray alpha tests
        img = np.zeros((Hh, W, 4), np.float32)
        H.hostsim_render_sample(hs, C.byref(a), sample, oracle._fp(img))
        assert np.isfinite(ref).all() and ref[..., :3].max() > 0
        assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)), "pixel_radius %g: %d pixels differ" % (pr, (ref != img).any(-1).sum())
        last = ref
    H.hostsim_scene_destroy(hs)
    single = scenes.textured_scene()
    ref1 = oracle.OracleScene(single).render_sample(W, Hh, single.camera, sp, 2, transmission=1, params=T.RenderParams(pixel_radius=3.0))
    assert not np.array_equal(ref1, last)   # the mip chain changes the image (coarser levels are read)
