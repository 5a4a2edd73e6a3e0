"""Generates the committed fixtures from the reference's OWN sources (through oracle/_ref/libref.so, built from
/root/reference by oracle/Makefile).  Run in the authoring container:  python oracle/gen_golden.py

  realtimepathtracingresearchframework_b200/data/sky_fits.json   update_sky_light fits for the SceneConfigs used by tests/bench
  tests/golden/ref_vectors.npz                                   input/output vectors of the reference's shading functions
  realtimepathtracingresearchframework_b200/data/pointset_tables.npz  the reference's Sobol / blue-noise sampler tables (data)
  tests/golden/ref_pointsets.npz                                 sampler streams of rendering/pointsets/{sobol,bn_rng,sample_order}.glsl
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from realtimepathtracingresearchframework_b200 import types as T  # noqa: E402

SKY_CONFIGS = [
    dict(),                                                        # SceneConfig defaults (render_params.glsl.h:157-162)
    dict(sun_dir=(0.35, 0.8, 0.45)),
    dict(sun_dir=(1.0, 0.25, 0.3), turbidity=5.0, albedo=(0.3, 0.3, 0.3)),
    dict(sun_dir=(0.0, -0.2, 1.0)),                                # sun below the horizon
]


def fa(*v):
    return (C.c_float * len(v))(*v)


def sky_key(cfg):
    return "%.6g|%.6g,%.6g,%.6g|%.6g|%.6g,%.6g,%.6g" % (cfg.bump_scale, *cfg.sun_dir, cfg.turbidity, *cfg.albedo)


def gen_sky():
    table = {}
    for kw in SKY_CONFIGS:
        cfg = T.SceneConfig(**kw)
        sp = po.sky_fit(cfg)
        table[sky_key(cfg)] = dict(sky_configs=[[float(x) for x in row] for row in sp.sky_configs], sky_radiances=[float(x) for x in sp.sky_radiances],
                                   sun_dir=[float(x) for x in sp.sun_dir], sun_cos_angle=float(sp.sun_cos_angle),
                                   sun_radiance=[float(x) for x in sp.sun_radiance], normal_z_scale=float(sp.normal_z_scale))
    path = os.path.join(ROOT, "realtimepathtracingresearchframework_b200", "data", "sky_fits.json")
    with open(path, "w") as f:
        json.dump(table, f, indent=1)
    print("wrote", path, len(table), "fits")


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def gen_vectors(n=512, seed=1234):
    R = po.ref()
    rng = np.random.default_rng(seed)
    out = {}
    # --- LCG / murmur: integer exact ---
    idx = rng.integers(0, 2 ** 32, (n, 3), dtype=np.uint64).astype(np.uint32)
    seeds = np.array([R.ref_lcg_seed(int(a), int(b), int(c)) for a, b, c in idx], np.uint32)
    draws = np.zeros((n, 4), np.float32)
    states = np.zeros((n, 4), np.uint32)
    for i in range(n):
        st = C.c_uint32(int(seeds[i]))
        for k in range(4):
            draws[i, k] = R.ref_lcg_randomf(C.byref(st))
            states[i, k] = st.value
    out.update(lcg_in=idx, lcg_seed=seeds, lcg_draws=draws, lcg_states=states)
    # --- glTF BSDF ---
    mats = np.zeros((n, 20), np.float32)  # BaseMaterial as 20 words
    mviews = mats.view(np.uint32)
    nrm = unit(rng.normal(size=(n, 3))).astype(np.float32)
    wo = unit(nrm + 0.9 * unit(rng.normal(size=(n, 3)))).astype(np.float32)
    wi = unit(nrm + 0.9 * unit(rng.normal(size=(n, 3)))).astype(np.float32)
    flip = rng.random(n) < 0.25
    wi[flip] = (wi[flip] - 2 * nrm[flip] * np.sum(wi[flip] * nrm[flip], -1, keepdims=True)).astype(np.float32)
    u4 = rng.random((n, 4)).astype(np.float32)
    res = {k: [] for k in ("bsdf", "wpdf", "sample", "bsdf_tr", "wpdf_tr", "sample_tr", "vx", "vy")}
    for i in range(n):
        m = T.BaseMaterial(base_color=tuple(rng.random(3)), roughness=float(rng.random()), metallic=float(rng.random() < 0.3) * float(rng.random()),
                           ior=float(1.0 if rng.random() < 0.15 else 1.05 + rng.random()), specular=0.5,
                           specular_transmission=float(rng.random() < 0.5) * float(rng.random()), clearcoat_gloss=float(rng.random()),
                           flags=int(rng.integers(0, 4)))
        mats[i] = np.frombuffer(bytes(m), np.float32)
        vx, vy = np.zeros(3, np.float32), np.zeros(3, np.float32)
        R.ref_ortho_basis(fa(*nrm[i]), vx.ctypes.data_as(po.f32p), vy.ctypes.data_as(po.f32p))
        res["vx"].append(vx), res["vy"].append(vy)
        for tr, suf in ((0, ""), (1, "_tr")):
            o3 = np.zeros(3, np.float32)
            R.ref_gltf_bsdf(C.byref(m), fa(*nrm[i]), fa(*wo[i]), fa(*wi[i]), tr, o3.ctypes.data_as(po.f32p))
            res["bsdf" + suf].append(o3)
            res["wpdf" + suf].append(R.ref_gltf_wpdf(C.byref(m), fa(*nrm[i]), fa(*wo[i]), fa(*wi[i]), tr))
            o8 = np.zeros(8, np.float32)
            R.ref_gltf_sample(C.byref(m), fa(*nrm[i]), fa(*wo[i]), fa(*vx), fa(*vy), fa(*u4[i, :2]), fa(*u4[i, 2:]), tr, o8.ctypes.data_as(po.f32p))
            res["sample" + suf].append(o8)
    out.update(mat=mviews.copy(), n=nrm, wo=wo, wi=wi, u4=u4, **{"gltf_" + k: np.array(v, np.float32) for k, v in res.items()})
    # --- triangle lights ---
    tri = (rng.normal(size=(n, 3, 3)) + rng.normal(size=(n, 1, 3)) * 2).astype(np.float32)
    d = unit(tri).astype(np.float32)
    sa = np.zeros((n, 4), np.float32)
    smp = np.zeros((n, 3), np.float32)
    u2 = rng.random((n, 2)).astype(np.float32)
    for i in range(n):
        R.ref_triangle_solid_angle(fa(*d[i, 0]), fa(*d[i, 1]), fa(*d[i, 2]), sa[i].ctypes.data_as(po.f32p))
        R.ref_sample_solid_angle_polygon(fa(*d[i, 0]), fa(*d[i, 1]), fa(*d[i, 2]), fa(*u2[i]), smp[i].ctypes.data_as(po.f32p))
    atan_in = np.concatenate([rng.normal(size=n // 2) * 3, rng.random(n // 2) * 40]).astype(np.float32)
    atan_out = np.array([R.ref_fast_positive_atan(float(x)) for x in atan_in], np.float32)
    out.update(tri_dirs=d, tri_solid_angle=sa, tri_sample=smp, tri_u2=u2, atan_in=atan_in, atan_out=atan_out)
    # binned RIS light selection
    n_l = 40
    lights = np.zeros((n_l, 12), np.float32)
    c = rng.normal(size=(n_l, 1, 3)) * 3 + np.array([0, 4, 0])
    lights[:, :9] = (c + rng.normal(size=(n_l, 3, 3)) * 0.5).reshape(n_l, 9)
    lights[:, 9:] = rng.random((n_l, 3)) * 10
    larr = (T.TriLightData * n_l).from_buffer_copy(lights.tobytes())
    hp = (rng.normal(size=(n, 3)) * 2).astype(np.float32)
    hn = unit(rng.normal(size=(n, 3))).astype(np.float32)
    ris = np.zeros((n, 9), np.float32)
    for i in range(n):
        R.ref_sample_tri_lights(larr, n_l, 16, fa(*hp[i]), fa(*hn[i]), fa(*u4[i, :2]), fa(*u4[i, 2:]), ris[i].ctypes.data_as(po.f32p))
    out.update(ris_lights=lights, ris_p=hp, ris_n=hn, ris_out=ris)
    # --- quantisation + hit attributes ---
    qv = rng.integers(0, 2 ** 63, (n, 3), dtype=np.uint64)
    qn = rng.integers(0, 2 ** 63, (n, 3), dtype=np.uint64)
    scale = np.array([2.0 ** -16, 2.0 ** -15, 3e-5], np.float32)
    offs = np.array([-16.0, 1.5, 0.25], np.float32)
    deq = np.zeros((n, 3, 3), np.float32)
    dn = np.zeros((n, 3, 3), np.float32)
    duv = np.zeros((n, 3, 2), np.float32)
    hit = np.zeros((n, 2, 14), np.float32)
    w2o = np.array([0.5, 0.1, 0, -0.1, 0.6, 0.2, 0.05, -0.2, 0.7], np.float32)
    bary = rng.random((n, 2)).astype(np.float32) * 0.5
    id4 = rng.integers(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32)
    for i in range(n):
        for k in range(3):
            R.ref_dequantize_position(C.c_uint64(int(qv[i, k])), fa(*scale), fa(*offs), deq[i, k].ctypes.data_as(po.f32p))
            R.ref_dequantize_normal(C.c_uint32(int(qn[i, k] & 0xFFFFFFFF)), dn[i, k].ctypes.data_as(po.f32p))
            R.ref_dequantize_uv(C.c_uint32(int(qn[i, k] >> 32)), duv[i, k].ctypes.data_as(po.f32p))
        qva = (C.c_uint64 * 3)(*[int(x) for x in qv[i]])
        qna = (C.c_uint64 * 3)(*[int(x) for x in qn[i]])
        for j, (hn_, hu_, mid) in enumerate(((0, 0, 3), (1, 1, -2))):
            R.ref_hit_attributes(qva, qna, fa(*scale), fa(*offs), hn_, hu_, fa(*w2o), mid, id4.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(i % 16),
                                 C.c_float(1.5), C.c_float(float(bary[i, 0])), C.c_float(float(bary[i, 1])), hit[i, j].ctypes.data_as(po.f32p))
    out.update(q_verts=qv, q_nuv=qn, q_scale=scale, q_offset=offs, deq_pos=deq, deq_normal=dn, deq_uv=duv, hit_w2o=w2o, hit_bary=bary,
               hit_id4=id4, hit_out=hit)
    # --- host light binning ---
    for tag, cnt in (("small", 2), ("mid", 37), ("big", 300)):
        em = np.zeros((cnt, 12), np.float32)
        cc = rng.normal(size=(cnt, 1, 3)) * 4
        em[:, :9] = (cc + rng.normal(size=(cnt, 3, 3)) * rng.random((cnt, 1, 1)) * 1.5).reshape(cnt, 9)
        em[:, 9:] = rng.random((cnt, 3)) * (10.0 ** rng.uniform(-1, 2, (cnt, 1)))
        ls = T.LightSamplingConfig()
        outl = (T.TriLightData * (2 * cnt + 64))()
        m = R.ref_bin_emitters((T.TriLightData * cnt).from_buffer_copy(em.tobytes()), cnt, C.byref(ls), outl, 2 * cnt + 64)
        assert m > 0
        out["bin_%s_in" % tag] = em
        out["bin_%s_out" % tag] = np.frombuffer(outl, np.float32).reshape(-1, 12)[:m].copy()
    # --- sky model + sun sampling: rendering/lights/sky_model_arhosek/sky_model.glsl:40-59, rendering/lights/sun.glsl:9-20 ---
    sky_dirs = unit(rng.normal(size=(n, 3))).astype(np.float32)
    sky_dirs[: n // 2, 1] = np.abs(sky_dirs[: n // 2, 1])
    sky_dirs[:8, 1] = np.array([0.0, 1e-4, 1.0, 0.999, 0.5, 0.01, 0.25, 0.75], np.float32)
    sky_dirs = unit(sky_dirs).astype(np.float32)
    for ci, kw in enumerate(SKY_CONFIGS):
        sp = po.sky_fit(T.SceneConfig(**kw))
        sd = np.array(list(sp.sun_dir), np.float32)
        rad = np.zeros((n, 3), np.float32)
        for i in range(n):
            R.ref_skymodel_radiance(C.byref(sp), fa(*sd), fa(*sky_dirs[i]), rad[i].ctypes.data_as(po.f32p))
        out["sky%d_radiance" % ci] = rad
        sun = np.zeros((n, 4), np.float32)
        for i in range(n):
            R.ref_sample_sun_dir(fa(*sd), C.c_float(float(sp.sun_cos_angle)), fa(*u4[i, :2]), sun[i].ctypes.data_as(po.f32p))
        out["sky%d_sun_samples" % ci] = sun
        # compute_sky_illum (vulkan/pt_megakernel.glsl:113-149) executed from the reference: the sky dirs plus directions inside
        # and around the sun disc, previous-bounce pdfs from "camera ray" (2e16) down to diffuse-like values, p_sun = 1 and 0.5
        rs = np.random.default_rng(977 + ci)  # own stream: the sections below keep their inputs
        dirs = sky_dirs.copy()
        dirs[8:72] = unit(sd + 0.006 * rs.normal(size=(64, 3))).astype(np.float32)
        dirs[72] = sd
        pdfs = np.where(rs.random(n) < 0.25, 2.0e16, 10.0 ** rs.uniform(-2, 4, n)).astype(np.float32)
        ill = np.zeros((2, n, 3), np.float32)
        for k, p_sun in enumerate((1.0, 0.5)):
            sp2 = T.SceneParams.from_buffer_copy(sp)
            sp2.sun_radiance[3] = p_sun
            for i in range(n):
                R.ref_compute_sky_illum(C.byref(sp2), fa(0.0, 0.0, 0.0), fa(*dirs[i]), C.c_float(float(pdfs[i])), ill[k, i].ctypes.data_as(po.f32p))
        out["sky%d_illum_dirs" % ci] = dirs
        out["sky%d_illum_pdfs" % ci] = pdfs
        out["sky%d_illum" % ci] = ill
    out["sky_dirs"] = sky_dirs
    # --- two blocks of main_spp executed from the reference (oracle/ref_shim/ref_loop.cpp): bounce prologue + Russian roulette ---
    rs = np.random.default_rng(4242)  # own stream: the sections below keep their inputs
    n_pl = 1024
    pl = np.zeros((n_pl, 20), np.float32)
    gn = unit(rs.normal(size=(n_pl, 3)))
    nrm = unit(gn + 0.4 * rs.normal(size=(n_pl, 3)))
    tan = unit(np.cross(nrm, unit(rs.normal(size=(n_pl, 3))))) * rs.uniform(0.2, 3.0, (n_pl, 1))
    pl[:, 0:3] = nrm
    pl[:, 3] = 10.0 ** rs.uniform(-2, 2, n_pl)
    pl[:, 4:7] = gn * 10.0 ** rs.uniform(-4, 1, (n_pl, 1))  # area-scaled geometric normal
    pl[:, 7:10] = tan
    pl[:, 10] = rs.uniform(0.2, 3.0, n_pl)
    pl[:, 11:14] = rs.normal(size=(n_pl, 3)) * 3
    rd = unit(-gn + 0.9 * unit(rs.normal(size=(n_pl, 3))))
    rd[::3] = unit(rs.normal(size=(len(rd[::3]), 3)))  # back-facing and grazing hits, shading normal on the far side
    pl[:, 14:17] = rd
    tex8 = rs.integers(0, 256, (n_pl, 3)).astype(np.uint8)
    tex8[::5] = (127, 127, 255)  # the loader's default normal texel (librender/scene.cpp:870-941)
    pl[:, 17:20] = tex8.astype(np.float32) / np.float32(255.0)
    flags = rs.choice([0, T.BASE_MATERIAL_ONESIDED, T.BASE_MATERIAL_VOLUME, T.BASE_MATERIAL_NOALPHA], n_pl).astype(np.uint32)
    has_map = (rs.random(n_pl) < 0.5).astype(np.int32)
    zscale = np.where(rs.random(n_pl) < 0.5, 1.0, rs.uniform(0.25, 4.0, n_pl)).astype(np.float32)
    plo = np.zeros((n_pl, 17), np.float32)
    for i in range(n_pl):
        R.ref_bounce_prologue(pl[i].ctypes.data_as(po.f32p), int(flags[i]), 0 if has_map[i] else -1, C.c_float(float(zscale[i])), plo[i].ctypes.data_as(po.f32p))
    out.update(pl_in=pl, pl_tex8=tex8, pl_flags=flags, pl_has_map=has_map, pl_zscale=zscale, pl_out=plo)
    n_rr = 2048
    thr = (10.0 ** rs.uniform(-3, 0.5, (n_rr, 3))).astype(np.float32)
    thr[::7] = rs.uniform(0.9, 1.1, (len(thr[::7]), 3)).astype(np.float32)
    rr_bounce = rs.integers(0, 12, n_rr).astype(np.int32)
    rr_depth = rs.choice([0, 2, 5], n_rr).astype(np.int32)
    rr_u = rs.random(n_rr).astype(np.float32)
    rr_out = np.zeros((n_rr, 4), np.float32)
    for i in range(n_rr):
        t = thr[i].copy()
        rr_out[i, 3] = R.ref_russian_roulette(int(rr_bounce[i]), int(rr_depth[i]), t.ctypes.data_as(po.f32p), C.c_float(float(rr_u[i])))
        rr_out[i, :3] = t
    out.update(rr_thr=thr, rr_bounce=rr_bounce, rr_depth=rr_depth, rr_u=rr_u, rr_out=rr_out)
    # ray-generation head of main_spp, geometry_scale_to_tmin, the resolve's running mean (same shim, own stream)
    from realtimepathtracingresearchframework_b200 import scenes
    cam = scenes.random_triangles(16).camera
    n_rg = 1024
    Wd, Hd = 1920, 1080
    vp = po.view_params(cam, Wd, Hd)  # du, dv, top_left (itself pinned by construction: SURVEY 8a-1)
    camv = np.concatenate([np.array(list(cam.pos), np.float32), vp]).astype(np.float32)
    rg = np.zeros((n_rg, 5), np.uint32)  # px, py, sample_index, frame_offset, enable_raster_taa
    rg[:, 0] = rs.integers(0, Wd, n_rg)
    rg[:, 1] = rs.integers(0, Hd, n_rg)
    rg[:8, 0] = [0, Wd - 1, 0, Wd - 1, 1, 2, 3, 4]
    rg[:8, 1] = [0, 0, Hd - 1, Hd - 1, 1, 2, 3, 4]
    rg[:, 2] = rs.integers(0, 5000, n_rg)
    rg[:, 3] = rs.integers(0, 100, n_rg)
    rg[:, 4] = rs.random(n_rg) < 0.3
    rg_j = np.zeros((n_rg, 2), np.float32)
    rg_o = np.zeros((n_rg, 9), np.float32)
    for i in range(n_rg):
        if rg[i, 4]:
            po.lib().oracle_screen_jitter(int(rg[i, 3]), int(rg[i, 2]), Wd, Hd, rg_j[i].ctypes.data_as(po.f32p))  # table pinned in ref_pointsets.npz
        R.ref_camera_ray(camv.ctypes.data_as(po.f32p), Wd, Hd, int(rg[i, 0]), int(rg[i, 1]), int(rg[i, 2]), int(rg[i, 3]), int(rg[i, 4]),
                         rg_j[i].ctypes.data_as(po.f32p), rg_o[i].ctypes.data_as(po.f32p))
    out.update(rg_cam=camv, rg_in=rg, rg_jitter=rg_j, rg_out=rg_o)
    n_tm = 512
    tm_in = np.zeros((n_tm, 4), np.float32)
    tm_in[:, :3] = rs.normal(size=(n_tm, 3)) * 10.0 ** rs.uniform(-2, 3, (n_tm, 1))
    tm_in[:, 3] = 10.0 ** rs.uniform(-3, 3, n_tm)
    tm_in[::9, 3] = 0.0
    out["tmin_in"] = tm_in
    out["tmin_out"] = np.array([R.ref_geometry_scale_to_tmin(tm_in[i, :3].copy().ctypes.data_as(po.f32p), C.c_float(float(tm_in[i, 3]))) for i in range(n_tm)],
                               np.float32)
    n_rm = 1024
    rm_x = (10.0 ** rs.uniform(-3, 3, (n_rm, 4))).astype(np.float32)
    rm_h = (10.0 ** rs.uniform(-3, 3, (n_rm, 4))).astype(np.float32)
    rm_n = np.stack([rs.integers(1, 5000, n_rm), rs.choice([1, 1, 1, 4, 64], n_rm)], 1).astype(np.uint32)
    rm_o = rm_h.copy()
    for i in range(n_rm):
        R.ref_running_mean(rm_x[i].ctypes.data_as(po.f32p), rm_o[i].ctypes.data_as(po.f32p), int(rm_n[i, 0]), int(rm_n[i, 1]))
    out.update(rm_x=rm_x, rm_hist=rm_h, rm_n=rm_n, rm_out=rm_o)
    # camera basis of update_view_parameters (vulkan/render_vulkan.cpp:2887-2895), cut out of the reference's host code
    n_cb = 256
    cb_in = np.zeros((n_cb, 12), np.float32)  # pos3, dir3, up3, fovy, w, h
    cb_in[:, 0:3] = rs.normal(size=(n_cb, 3)) * 10
    cb_in[:, 3:6] = unit(rs.normal(size=(n_cb, 3))) * rs.choice([1.0, 1.0, 0.5, 3.0], (n_cb, 1))  # dir is NOT normalised by the function
    cb_in[:, 6:9] = unit(rs.normal(size=(n_cb, 3)))
    cb_in[:, 9] = rs.uniform(10, 120, n_cb)
    cb_in[:, 10:12] = np.array([(1920, 1080), (1280, 720), (640, 480), (333, 777)], np.float32)[rs.integers(0, 4, n_cb)]
    cb_out = np.zeros((n_cb, 9), np.float32)
    for i in range(n_cb):
        c = T.RenderCameraParams()
        c.pos[:], c.dir[:], c.up[:], c.fovy = cb_in[i, 0:3].tolist(), cb_in[i, 3:6].tolist(), cb_in[i, 6:9].tolist(), float(cb_in[i, 9])
        R.ref_view_params(C.byref(c), int(cb_in[i, 10]), int(cb_in[i, 11]), cb_out[i].ctypes.data_as(po.f32p))
    out.update(cb_in=cb_in, cb_out=cb_out)
    # display chain: tonemap() + linear_to_srgb() executed from the reference (own stream)
    rt = np.random.default_rng(90210)
    tm_rgb = (10.0 ** rt.uniform(-4, 2, (768, 3))).astype(np.float32)
    tm_rgb[::11] = 0.0
    tm_mode = rt.integers(0, 3, 768).astype(np.int32)
    tm_out = np.zeros((768, 6), np.float32)
    for i in range(768):
        R.ref_tonemap_srgb(int(tm_mode[i]), tm_rgb[i].ctypes.data_as(po.f32p), tm_out[i].ctypes.data_as(po.f32p))
    out.update(tm_rgb=tm_rgb, tm_mode=tm_mode, tm_out=tm_out)
    # the reference's srgb_to_linear over all 256 code values (cross-check of the texel decode)
    out["srgb_decode"] = np.array([R.ref_srgb_to_linear(C.c_float(float(np.float32(v) / np.float32(255.0)))) for v in range(256)], np.float32)
    # rt_intersect.comp:main over scripted ray queries (own stream)
    rq = np.random.default_rng(31337)
    n_rq = 512
    rq_q = np.zeros((n_rq, 8), np.float32)
    rq_q[:, 0:3] = rq.normal(size=(n_rq, 3)) * 10.0 ** rq.uniform(-2, 3, (n_rq, 1))
    rq_mode = rq.choice([0, 0, 0, 1, 7, -1, -5], n_rq).astype(np.int32)
    rq_q[:, 3] = rq_mode.view(np.float32)
    rq_q[:, 4:7] = unit(rq.normal(size=(n_rq, 3)))
    rq_q[:, 7] = rq.choice([1e20, 5.0, 0.5], n_rq)
    rq_hit = np.zeros((n_rq, 6), np.float32)  # hit, bary2, custom index, geometry index, prim
    rq_hit[:, 0] = rq.random(n_rq) < 0.6
    rq_hit[:, 1:3] = rq.random((n_rq, 2)) * 0.5
    rq_hit[:, 3] = rq.integers(0, 1000, n_rq)
    rq_hit[:, 4] = rq.integers(0, 8, n_rq)
    rq_hit[:, 5] = rq.integers(0, 2 ** 24, n_rq)
    rq_res = np.full((n_rq, 4), 123.0, np.float32)  # what the caller passed in: a skipped query must leave it alone
    rq_info = np.zeros((n_rq, 4), np.float32)
    for i in range(n_rq):
        R.ref_ray_query(rq_q[i].ctypes.data_as(po.f32p), int(rq_hit[i, 0]), rq_hit[i, 1:3].copy().ctypes.data_as(po.f32p), int(rq_hit[i, 3]),
                        int(rq_hit[i, 4]), int(rq_hit[i, 5]), rq_res[i].ctypes.data_as(po.f32p), rq_info[i].ctypes.data_as(po.f32p))
    out.update(rq_q=rq_q, rq_hit=rq_hit, rq_res=rq_res, rq_info=rq_info)
    # raytrace_test_visibility over scripted ray queries + the alpha test of generate_candidate_hit
    n_vis = 768
    vis_in = np.zeros((n_vis, 14), np.float32)  # from3, dir3, dist, geometry_scale, frame_id, frame_offset, px, py, n_cands, opaque_hit
    vis_c = np.zeros((n_vis, 6, 4), np.float32)  # (t, prim, inst, accept) per candidate
    vis_o = np.zeros((n_vis, 5 + 6), np.float32)
    vis_in[:, 0:3] = rs.normal(size=(n_vis, 3)) * 10.0 ** rs.uniform(-1, 2, (n_vis, 1))
    vis_in[:, 3:6] = unit(rs.normal(size=(n_vis, 3)))
    vis_in[:, 6] = 10.0 ** rs.uniform(-5, 3, n_vis)       # light distance, down to below the ray epsilon
    vis_in[:, 7] = 10.0 ** rs.uniform(-2, 3, n_vis)       # geometry_scale (total_t)
    vis_in[:, 8] = rs.integers(0, 5000, n_vis)
    vis_in[:, 9] = rs.integers(0, 100, n_vis)
    vis_in[:, 10] = rs.integers(0, Wd, n_vis)
    vis_in[:, 11] = rs.integers(0, Hd, n_vis)
    vis_in[:, 12] = rs.integers(0, 7, n_vis)
    vis_in[:, 13] = rs.random(n_vis) < 0.15
    for i in range(n_vis):
        k = int(vis_in[i, 12])
        vis_c[i, :k, 0] = np.sort(rs.uniform(0, 1, k)) * vis_in[i, 6]
        vis_c[i, :k, 1] = rs.choice(100000, k, replace=False)
        vis_c[i, :k, 2] = rs.integers(0, 100, k)
        vis_c[i, :k, 3] = rs.random(k) < 0.25
        R.ref_test_visibility(vis_in[i, 0:3].copy().ctypes.data_as(po.f32p), vis_in[i, 3:6].copy().ctypes.data_as(po.f32p), C.c_float(float(vis_in[i, 6])),
                              C.c_float(float(vis_in[i, 7])), int(vis_in[i, 8]), int(vis_in[i, 9]), int(vis_in[i, 10]), int(vis_in[i, 11]), Wd, Hd,
                              np.ascontiguousarray(vis_c[i]).ctypes.data_as(po.f32p), k, int(vis_in[i, 13]), vis_o[i].ctypes.data_as(po.f32p))
    out.update(vis_in=vis_in, vis_cands=vis_c, vis_out=vis_o)
    n_af = 1024
    af_alpha = rs.choice([0.0, 1.0, 0.5, 0.25, 128.0 / 255.0, -0.0], n_af).astype(np.float32)
    af_alpha[::3] = rs.random(len(af_alpha[::3])).astype(np.float32)
    af_flags = rs.choice([0, T.BASE_MATERIAL_NOALPHA, T.BASE_MATERIAL_ONESIDED], n_af).astype(np.uint32)
    af_state = rs.integers(0, 2 ** 32, n_af, dtype=np.uint64).astype(np.uint32)
    af_out = np.zeros((n_af, 2), np.uint32)
    for i in range(n_af):
        st = C.c_uint32(int(af_state[i]))
        af_out[i, 0] = R.ref_alpha_filter(C.c_float(float(af_alpha[i])), int(af_flags[i]), C.byref(st))
        af_out[i, 1] = st.value
    out.update(af_alpha=af_alpha, af_flags=af_flags, af_state=af_state, af_out=af_out)
    # --- next-event estimation: rendering/mc/nee.glsl:32-90 (sample_direct_light) executed from the reference ---
    n_nee = 768
    nn = unit(rng.normal(size=(n_nee, 3))).astype(np.float32)
    gnn = unit(nn + 0.25 * rng.normal(size=(n_nee, 3))).astype(np.float32)
    woo = unit(nn + 0.9 * unit(rng.normal(size=(n_nee, 3)))).astype(np.float32)
    hpp = (rng.normal(size=(n_nee, 3)) * 2).astype(np.float32)
    uu = rng.random((n_nee, 4)).astype(np.float32)
    nee_mats = np.zeros((n_nee, 20), np.uint32)
    nee_frames = np.zeros((n_nee, 6), np.float32)
    sp = po.sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    sun_dir = np.array(list(sp.sun_dir), np.float32)
    nee_out = np.zeros((2, n_nee, 16), np.float32)
    for i in range(n_nee):
        m = T.BaseMaterial(base_color=tuple(float(np.float32(x)) for x in rng.random(3)), roughness=float(np.float32(rng.uniform(0.03, 1.0))),
                           metallic=float(rng.random() < 0.3) * float(np.float32(rng.random())),
                           ior=float(1.0 if rng.random() < 0.2 else np.float32(1.05 + rng.random())), specular=float(np.float32(rng.random())),
                           flags=T.BASE_MATERIAL_NOALPHA)
        nee_mats[i] = np.frombuffer(bytes(m), np.uint32)
        vx, vy = np.zeros(3, np.float32), np.zeros(3, np.float32)
        R.ref_ortho_basis(fa(*nn[i]), vx.ctypes.data_as(po.f32p), vy.ctypes.data_as(po.f32p))
        nee_frames[i] = np.concatenate([vx, vy])
        for mode, (p_sun, nl) in enumerate(((1.0, 0), (0.5, n_l))):
            sr = np.array([sp.sun_radiance[0], sp.sun_radiance[1], sp.sun_radiance[2], p_sun], np.float32)
            R.ref_sample_direct_light(C.byref(m), fa(*hpp[i]), fa(*gnn[i]), fa(*nn[i]), fa(*vx), fa(*vy), fa(*woo[i]), fa(*uu[i]), fa(*sun_dir),
                                      C.c_float(float(sp.sun_cos_angle)), fa(*sr), C.cast(larr, C.c_void_p), nl, 16,
                                      nee_out[mode, i].ctypes.data_as(po.f32p))
    out.update(nee_mat=nee_mats, nee_p=hpp, nee_gn=gnn, nee_n=nn, nee_wo=woo, nee_u4=uu, nee_frames=nee_frames, nee_sun_dir=sun_dir,
               nee_sun_cos=np.float32(sp.sun_cos_angle), nee_sun_rgb=np.array(list(sp.sun_radiance)[:3], np.float32), nee_out=nee_out)
    # --- the whole per-vertex shading function: rendering/mc/shade_base_material.glsl:14-96 executed from the reference ---
    n_sh = 1024
    sh_n = unit(rng.normal(size=(n_sh, 3))).astype(np.float32)
    sh_gn = unit(sh_n + 0.2 * rng.normal(size=(n_sh, 3))).astype(np.float32)
    sh_wo = unit(sh_n + 0.9 * unit(rng.normal(size=(n_sh, 3)))).astype(np.float32)
    sh_p = (rng.normal(size=(n_sh, 3)) * 2).astype(np.float32)
    sh_ia = np.zeros((n_sh, 15), np.float32)
    sh_mat = np.zeros((n_sh, 20), np.uint32)
    sh_state = np.zeros((n_sh, 12), np.float32)   # bounce, output_channel, prev_pdf, illum3, thr3, approx_sa, max_depth, glossy_only
    sh_rng = rng.integers(0, 2 ** 32, n_sh, dtype=np.uint64).astype(np.uint32)
    sh_out = np.zeros((2, n_sh, 19), np.float32)
    for i in range(n_sh):
        emissive = rng.random() < 0.15
        m = T.BaseMaterial(base_color=tuple(float(np.float32(x)) for x in rng.random(3)), roughness=float(np.float32(rng.uniform(0.03, 1.0))),
                           metallic=float(rng.random() < 0.3) * float(np.float32(rng.random())),
                           ior=float(1.0 if rng.random() < 0.2 else np.float32(1.05 + rng.random())), specular=float(np.float32(rng.random())),
                           emission_intensity=float(np.float32(rng.uniform(1, 30))) if emissive else 0.0, flags=T.BASE_MATERIAL_NOALPHA)
        sh_mat[i] = np.frombuffer(bytes(m), np.uint32)
        vx, vy = np.zeros(3, np.float32), np.zeros(3, np.float32)
        R.ref_ortho_basis(fa(*sh_n[i]), vx.ctypes.data_as(po.f32p), vy.ctypes.data_as(po.f32p))
        sh_ia[i] = np.concatenate([sh_p[i], sh_gn[i], sh_n[i], vx, vy])
        bounce = int(rng.integers(0, 5))
        channel = int(rng.integers(1, 4)) if rng.random() < 0.1 else 0
        prev_pdf = float(np.float32(2.e16 if bounce == 0 else 10.0 ** rng.uniform(-2, 3)))
        il = (rng.random(3) * (bounce > 0)).astype(np.float32)
        thr = (rng.random(3) ** 2).astype(np.float32) if bounce else np.ones(3, np.float32)
        sa = float(np.float32(10.0 ** rng.uniform(-5, -1)))
        max_depth = int(rng.choice([9, 9, 9, bounce + 1, bounce + 2]))
        glossy = int(rng.random() < 0.1)
        sh_state[i] = [bounce, channel, prev_pdf, *il, *thr, sa, max_depth, glossy]
        for mode, (p_sun, nl) in enumerate(((1.0, 0), (0.5, n_l))):
            sr = np.array([sp.sun_radiance[0], sp.sun_radiance[1], sp.sun_radiance[2], p_sun], np.float32)
            R.ref_shade_base_material(C.byref(m), bounce, channel, C.c_float(prev_pdf), fa(*il), fa(*thr), C.c_float(sa), fa(*sh_wo[i]), fa(*sh_ia[i]),
                                      int(sh_rng[i]), max_depth, glossy, fa(*sun_dir), C.c_float(float(sp.sun_cos_angle)), fa(*sr),
                                      C.cast(larr, C.c_void_p), nl, 16, sh_out[mode, i].ctypes.data_as(po.f32p))
    out.update(sh_mat=sh_mat, sh_ia=sh_ia, sh_wo=sh_wo, sh_state=sh_state, sh_rng=sh_rng, sh_out=sh_out)
    # --- material decode with texture handles (rendering/rt/material_textures.glsl:95-145 + gltf_bsdf.glsl:38-62) ---
    # The reference's unpack_material / get_material_alpha run on 1 x 1 textures; the sampler's return value for an 8-bit texel
    # is OUR statement of the texture unit (UNORM8 -> v / 255 in float, sRGB colour channels through the transfer function in
    # double, rounded once), so the fixture holds both the 8-bit texels and the floats handed to the reference.
    n_tex, n_mat = 12, 320
    tex8 = rng.integers(0, 256, (n_tex, 4)).astype(np.uint8)
    tex8[0] = (255, 255, 255, 255)
    tex8[1] = (10, 200, 30, 0)       # alpha 0: no premultiplied-alpha divide
    tex8[2] = (128, 64, 32, 1)       # alpha 1/255 > 0.001: divide
    tex_srgb = (np.arange(n_tex) % 2).astype(np.int32)
    tex_ch = np.where(np.arange(n_tex) % 5 == 4, 3, 4).astype(np.int32)  # some RGB-only textures: alpha reads 1
    texf = np.zeros((n_tex, 4), np.float32)
    for t in range(n_tex):
        for k in range(4):
            v = int(tex8[t, k])
            if k == 3:
                texf[t, k] = np.float32(v) / np.float32(255.0) if tex_ch[t] == 4 else np.float32(1.0)
            elif tex_srgb[t]:
                c = v / 255.0
                texf[t, k] = np.float32(c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4)
            else:
                texf[t, k] = np.float32(v) / np.float32(255.0)
    mat_words = np.zeros((n_mat, 20), np.uint32)
    res = np.zeros((2, n_mat, 17), np.float32)
    for i in range(n_mat):
        def scalar(lo, hi, p_tex=0.3):
            if rng.random() < p_tex:
                return T.texture_handle(int(rng.integers(0, n_tex)), int(rng.integers(0, 4)))
            return float(np.float32(rng.uniform(lo, hi)))
        bc = tuple(float(np.float32(x)) for x in rng.random(3))
        if rng.random() < 0.5:
            bc = (T.texture_handle(int(rng.integers(0, n_tex))), bc[1], bc[2])
        m = T.BaseMaterial(base_color=bc, roughness=scalar(0.02, 1.0), metallic=scalar(0.0, 1.0), specular=scalar(0.0, 1.0),
                           ior=(1.0 if rng.random() < 0.15 else scalar(1.05, 2.2, 0.1)),
                           specular_transmission=(0.0 if rng.random() < 0.4 else scalar(0.0, 1.0)), clearcoat_gloss=scalar(0.0, 1.0),
                           emission_intensity=(float(np.float32(rng.uniform(0.5, 30))) if rng.random() < 0.25 else 0.0),
                           flags=int(rng.integers(0, 16)))
        mat_words[i] = np.frombuffer(bytes(m), np.uint32)
        for tr in (0, 1):
            R.ref_unpack_material(C.byref(m), texf.ctypes.data_as(po.f32p), n_tex, tr, res[tr, i].ctypes.data_as(po.f32p))
    out.update(mt_tex8=tex8, mt_tex_srgb=tex_srgb, mt_tex_channels=tex_ch, mt_tex_float=texf, mt_materials=mat_words, mt_unpacked=res)
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


# The sampler calls one path makes (vulkan/pt_megakernel.glsl:314-317,423,719; rendering/mc/shade_base_material.glsl:59-82):
# op 0 = RANDOM_FLOAT1(rng, arg), 1 = RANDOM_SET_DIM, 2 = RANDOM_SHIFT_DIM
def path_ops(n_vertices):
    ops = [(0, 0), (0, 1)]
    for v in range(n_vertices):
        ops += [(1, 6 + 8 * v), (0, 2), (0, 3), (0, 0), (0, 1), (2, 4), (0, 2), (0, 3), (0, 0), (0, 1), (2, 4), (0, -1)]
    return ops


def gen_pointsets(n=384, seed=4321):
    R = po.ref()
    rng = np.random.default_rng(seed)
    # --- tables: the reference's data, stored in the smallest exact integer type ---
    tabs = []
    for which in range(4):
        cnt = R.ref_pointset_table(which, None)
        buf = np.zeros(cnt, np.uint32)
        R.ref_pointset_table(which, buf.ctypes.data_as(C.POINTER(C.c_uint32)))
        tabs.append(buf)
    assert tabs[1].max() < 65536 and tabs[2].max() < 256 and tabs[3].max() < 256
    path = os.path.join(ROOT, "realtimepathtracingresearchframework_b200", "data", "pointset_tables.npz")
    np.savez_compressed(path, sobol_matrix=tabs[0], sobol_tile_invert=tabs[1].astype(np.uint16), bn_sobol=tabs[2].astype(np.uint8),
                        bn_scrambling_1spp=tabs[3].astype(np.uint8))
    print("wrote", path, os.path.getsize(path), "bytes")
    # --- sampler streams ---
    ops = path_ops(4) + [(1, 1019), (0, 0), (0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 2047), (0, 1), (0, 2)]
    op_a = np.array([o for o, _ in ops], np.int32)
    arg_a = np.array([a for _, a in ops], np.int32)
    n_draws = int((op_a == 0).sum())
    out = dict(ops=op_a, args=arg_a)
    for variant in (1, 2, 3):
        q = np.zeros((n, 6), np.uint32)  # sample_index, frame_id, frame_offset, px, py, width
        q[:, 0] = rng.integers(0, 5000, n)
        q[: n // 8, 0] = rng.integers(0, 65536, n // 8)      # large sample indices (index bits up to 2^32 for Z_SBL)
        q[:, 1] = rng.integers(0, 5000, n)
        q[:, 2] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        q[: n // 4, 2] = rng.integers(0, 4, n // 4)
        q[:, 5] = rng.choice([1920, 1280, 256, 3840, 333], n)
        q[:, 3] = rng.integers(0, 4096, n) % q[:, 5]
        q[:, 4] = rng.integers(0, 2160, n)
        draws = np.zeros((n, n_draws), np.float32)
        state = np.zeros((n, 2), np.uint32)
        for i in range(n):
            m = R.ref_pointset_replay(variant, int(q[i, 0]), int(q[i, 1]), int(q[i, 2]), int(q[i, 3]), int(q[i, 4]), int(q[i, 5]), 1080,
                                      op_a.ctypes.data_as(C.POINTER(C.c_int32)), arg_a.ctypes.data_as(C.POINTER(C.c_int32)), len(ops),
                                      draws[i].ctypes.data_as(po.f32p), state[i].ctypes.data_as(C.POINTER(C.c_uint32)))
            assert m == n_draws
        out["v%d_in" % variant] = q
        out["v%d_draws" % variant] = draws
        out["v%d_state" % variant] = state
    # --- morton_sample_id (sample_order.glsl:21-73) with every flag combination and non-square / non-power-of-two tiles ---
    mq = np.zeros((n, 7), np.uint32)
    mq[:, 0] = rng.integers(0, 1000, n)
    mq[:, 1] = rng.integers(0, 4096, n)
    mq[:, 2] = rng.integers(0, 4096, n)
    tiles = np.array([(256, 256), (8, 8), (16, 64), (128, 32), (100, 60), (1920, 1080), (2, 2)], np.uint32)
    mq[:, 3:5] = tiles[rng.integers(0, len(tiles), n)]
    mq[:, 5] = rng.integers(0, 2, n)
    mq[:, 6] = rng.integers(0, 2, n)
    hal = np.zeros((R.ref_halton_23(None), 2), np.float32)
    R.ref_halton_23(hal.ctypes.data_as(po.f32p))
    out["halton_23"] = hal
    # the screen-jitter statements of update_view_parameters executed from the reference's host code
    rj = np.random.default_rng(611)
    jin = np.zeros((256, 4), np.uint32)
    jin[:, 0] = rj.integers(0, 2 ** 32, 256, dtype=np.uint64).astype(np.uint32)
    jin[:64, 0] = np.arange(64)
    jin[:, 1] = rj.integers(0, 5000, 256)
    jin[:, 2:4] = np.array([(1920, 1080), (1280, 720), (333, 77), (640, 480)], np.uint32)[rj.integers(0, 4, 256)]
    jout = np.zeros((256, 2), np.float32)
    for i in range(256):
        R.ref_screen_jitter(int(jin[i, 0]), int(jin[i, 1]), int(jin[i, 2]), int(jin[i, 3]), jout[i].ctypes.data_as(po.f32p))
    out["jitter_in"] = jin
    out["jitter_out"] = jout
    R.ref_morton_sample_id.restype = C.c_uint32
    out["morton_in"] = mq
    out["morton_out"] = np.array([R.ref_morton_sample_id(*[int(x) for x in row]) for row in mq], np.uint32)
    path = os.path.join(ROOT, "tests", "golden", "ref_pointsets.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def gen_queries(seed=77):
    """render_ray_queries: the reference's dispatch size, invocation index / swizzled invocation id of every query of a few
    dispatches, and accumulate_query chains -- all executed from the reference's files (oracle/ref_shim/ref_queries.cpp)."""
    R = po.ref()
    R.ref_query_invocation.argtypes = [C.c_uint32] * 5 + [C.POINTER(C.c_uint32)]
    R.ref_accumulate_query.argtypes = [po.f32p, po.f32p, C.c_uint32]
    R.ref_query_dispatch.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    out = {}
    counts = np.array([1, 2, 31, 32, 33, 511, 512, 513, 1000, 1024, 1025, 4097, 65536, 250000, 262144, 1000003], np.int32)
    disp = np.zeros((len(counts), 5), np.int32)
    for i, n in enumerate(counts):
        R.ref_query_dispatch(int(n), 3, disp[i].ctypes.data_as(C.POINTER(C.c_int32)))
    out["q_counts"], out["q_dispatch"] = counts, disp
    # every invocation of the dispatches of 513 and 4097 queries: (swizzled x, swizzled y, invocation index)
    for n, tag in ((513, "513"), (4097, "4097")):
        d = disp[list(counts).index(n)]
        nwx, nwy = int(d[2]), int(d[3])
        rows = np.zeros((nwx * nwy * 512, 3), np.uint32)
        k = 0
        for wy in range(nwy):
            for wx in range(nwx):
                for l in range(512):
                    R.ref_query_invocation(wx, wy, nwx, nwy, l, rows[k].ctypes.data_as(C.POINTER(C.c_uint32)))
                    k += 1
        out["q_invocations_" + tag] = rows
    rng = np.random.default_rng(seed)
    xs = rng.uniform(0.0, 4.0, (64, 6, 4)).astype(np.float32)  # 64 chains of 6 layers
    res = np.zeros((64, 6, 4), np.float32)
    for c in range(64):
        r = rng.uniform(-1, 1, 4).astype(np.float32)  # stale content of the result buffer: layer 0 must overwrite it
        for k in range(6):
            R.ref_accumulate_query(r.ctypes.data_as(po.f32p), xs[c, k].ctypes.data_as(po.f32p), k)
            res[c, k] = r
    out["q_accum_in"], out["q_accum_out"] = xs, res
    path = os.path.join(ROOT, "tests", "golden", "ref_queries.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def gen_textures(n=1500, n_mats=300):
    """f2: the footprint algebra of rendering/rt/footprint.glsl executed from the reference's file (oracle/ref_shim/ref_footprint.cpp), and
    unpack_material / get_material_alpha of rendering/rt/material_textures.glsl in the USE_MIPMAPPING configuration (textureGrad) executed
    from the reference's files (ref_shim/ref_materials.cpp) over OUR texture unit (the oracle's textureGrad on tests/texture_util.py's
    texture set: the reference leaves that part to the hardware)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import texture_util as tu
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    L = po.lib()
    f32p = C.POINTER(C.c_float)
    R.ref_footprint_op.argtypes = [C.c_int32, f32p, f32p]
    L.oracle_footprint_op.argtypes = [C.c_int32, f32p, f32p]
    out = {}
    op0, d2 = tu.footprint_cases(n, 99)
    F = np.zeros((n, 4), np.float32)
    F2 = np.zeros((n, 4), np.float32)
    dp = np.zeros((n, 6), np.float32)
    for i in range(n):
        R.ref_footprint_op(0, op0[i].ctypes.data_as(f32p), F[i].ctypes.data_as(f32p))
        inp = np.concatenate([d2[i], op0[i, :3], F[i]]).astype(np.float32)
        R.ref_footprint_op(1, inp.ctypes.data_as(f32p), F2[i].ctypes.data_as(f32p))
        inp2 = np.concatenate([d2[i], F2[i]]).astype(np.float32)
        R.ref_footprint_op(2, inp2.ctypes.data_as(f32p), dp[i].ctypes.data_as(f32p))
    out["fp_to_footprint"], out["fp_reflect"], out["fp_to_dpdxy"] = F, F2, dp
    # material glue over our texture unit
    tset = tu.texture_set()
    descs, keep = tu.texture_descs(tset)
    L.oracle_sample_texture_grad.argtypes = [C.POINTER(T.TextureDesc), C.c_float, C.c_float, f32p, f32p, f32p]
    UNIT = C.CFUNCTYPE(None, C.c_uint32, f32p, f32p, f32p, f32p)

    def unit(tex_id, uv, dx, dy, rgba):
        L.oracle_sample_texture_grad(C.byref(descs[tex_id]), uv[0], uv[1], dx, dy, rgba)
    cb = UNIT(unit)
    R.ref_unpack_material_grad.argtypes = [C.POINTER(T.BaseMaterial), f32p, f32p, UNIT, f32p]
    mats, uv, duvdxy = tu.random_textured_materials(n_mats, len(tset), 123)
    res = np.zeros((n_mats, 17), np.float32)
    for i, m in enumerate(mats):
        R.ref_unpack_material_grad(C.byref(m), uv[i].ctypes.data_as(f32p), duvdxy[i].ctypes.data_as(f32p), cb, res[i].ctypes.data_as(f32p))
    out["mat_grad"] = res
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_textures.npz"), **out)
    print("wrote tests/golden/ref_textures.npz:", {k: v.shape for k, v in out.items()})


def gen_vks():
    """What the reference's own reader (ext/libvkr/src/vkr.c in oracle/_ref, through ref_shim/ref_vkr.c) reports for the .vks scene and
    texture directory tests/vks_util.py writes with our writer."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vks_util
    from realtimepathtracingresearchframework_b200 import vks
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    i64p, f32p = C.POINTER(C.c_int64), C.POINTER(C.c_float)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        path, _ = vks_util.write_test_scene(d)
        p = path.encode()
        counts = np.zeros(12, np.int64)
        assert R.ref_vkr_scene_counts(p, counts.ctypes.data_as(i64p)) == 0, "vkr_open_scene refused the file"
        out["counts"] = counts
        n_meshes, n_inst, n_mat = int(counts[1]), int(counts[2]), int(counts[3])
        ints, floats, segs, names = np.zeros((n_meshes, 10), np.int64), np.zeros((n_meshes, 6), np.float32), np.zeros((n_meshes, 16), np.int64), []
        for i in range(n_meshes):
            name = C.create_string_buffer(128)
            assert R.ref_vkr_mesh(p, C.c_int64(i), ints[i].ctypes.data_as(i64p), floats[i].ctypes.data_as(f32p), segs[i].ctypes.data_as(i64p), C.c_int64(8), name) == 0
            names.append(name.value.decode())
        out["mesh_ints"], out["mesh_floats"], out["mesh_segs"], out["mesh_names"] = ints, floats, segs, np.array(names)
        inst = np.zeros((n_inst, 3), np.int64)
        for i in range(n_inst):
            assert R.ref_vkr_instance(p, C.c_int64(i), inst[i].ctypes.data_as(i64p)) == 0
        out["instances"] = inst
        mf, mt, mn = np.zeros((n_mat, 8), np.float32), np.zeros((n_mat, 21), np.int64), []
        for i in range(n_mat):
            name = C.create_string_buffer(128)
            assert R.ref_vkr_material(p, C.c_int64(i), name, mf[i].ctypes.data_as(f32p), mt[i].ctypes.data_as(i64p)) == 0
            mn.append(name.value.decode())
        out["material_floats"], out["material_tex"], out["material_names"] = mf, mt, np.array(mn)
        c = vks.read_vks_container(path)
        tr = np.zeros((int(counts[7]), 12), np.float32)
        for k in range(len(tr)):
            rec = np.ascontiguousarray(c["transform_table"][24 * k:24 * k + 24])
            R.ref_vkr_dequantize_transform(rec.ctypes.data_as(C.POINTER(C.c_ubyte)), tr[k].ctypes.data_as(f32p))
        out["transforms"] = tr
        R.ref_vkr_transform_offset.restype = C.c_int64
        R.ref_vkr_transform_offset.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
        out["transform_offsets"] = np.array([R.ref_vkr_transform_offset(i, 3, 5, f) for i, f in ((0, 0), (2, 7), (3, 0), (4, 2), (7, 3))], np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_vks.npz"), **out)
    print("wrote tests/golden/ref_vks.npz:", {k: v.shape for k, v in out.items()})


POST_CASES = [(1, 1, 8), (2, 4, 32), (3, 1, 1)]   # (seed, batch, spp_accumulation_window) on 61 x 47 frames
TAA_CASES = [(11, 1), (12, 2)]                     # (seed, upscale) on 40 x 30 render frames


def gen_post():
    """The temporal passes executed from the reference's shader sources (oracle/ref_shim/ref_post.cpp) on the synthetic frames of
    tests/temporal_util.py: accumulator + display colour of reproject_and_accumulate, LDR target of process_taa."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from temporal_util import synthetic_frame
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    R.ref_reproject_accumulate.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
    R.ref_process_taa.argtypes = [C.c_int32] * 5 + [C.c_void_p] * 4
    out = {}
    for seed, batch, window in POST_CASES:
        rng = np.random.default_rng(seed)
        cur, hist, nd_hist, nd, mj = synthetic_frame(rng, 61, 47)
        stored, shown = po.reproject_accumulate(cur, hist, nd_hist, nd, mj, 1.0 / window, batch, fn=R.ref_reproject_accumulate)
        out["reproject_%d_stored" % seed], out["reproject_%d_shown" % seed] = stored, shown
    for seed, upscale in TAA_CASES:
        rng = np.random.default_rng(seed)
        rw, rh = 40, 30
        w, h = rw * upscale, rh * upscale
        cur = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        his = np.clip(cur.astype(np.int32) + rng.integers(-40, 41, (h, w, 4)), 0, 255).astype(np.uint8)
        _, _, _, _, mj = synthetic_frame(rng, rw, rh, motion_scale=0.03)
        out["taa_%d" % seed] = po.process_taa(cur, his, mj, upscale, fn=R.ref_process_taa)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_post.npz"), **out)
    print("wrote tests/golden/ref_post.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if "--post-only" in sys.argv:
        gen_post()
        sys.exit(0)
    if "--textures-only" in sys.argv:
        gen_textures()
        sys.exit(0)
    if "--vks-only" in sys.argv:
        gen_vks()
        sys.exit(0)
    if "--queries-only" in sys.argv:
        gen_queries()
        sys.exit(0)
    if po.ref() is None:
        sys.exit("oracle/_ref/libref.so missing: run `make -C oracle` in a container that has /root/reference")
    if "--pointsets-only" not in sys.argv:
        gen_sky()
        gen_vectors()
    gen_pointsets()
    gen_queries()
    gen_post()
    gen_textures()
    gen_vks()
