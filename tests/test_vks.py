"""f2: `.vks` scene files with their `.vkt` texture directory (realtimepathtracingresearchframework_b200/vks.py) -- the standalone host
path's counterpart of Scene::load_vkrs (librender/scene.cpp:544-1006) + ext/libvkr/src/vkr.c.
* container: what OUR reader reports for a file written by OUR writer == what the REFERENCE's reader (vkr.c compiled into oracle/_ref,
  answers committed as tests/golden/ref_vks.npz) reports for the same bytes: counts, offsets, mesh / instance / material tables, texture
  headers, parameter files, the quantised-transform table;
* mapping: the loaded scene renders to the same image (oracle) as the same scene built directly with scenes.Scene."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import vks_util  # noqa: E402
from realtimepathtracingresearchframework_b200 import load_sky_fit, scenes, types as T, vks  # noqa: E402


@pytest.fixture(scope="module")
def yard(tmp_path_factory):
    d = tmp_path_factory.mktemp("vks")
    path, desc = vks_util.write_test_scene(str(d))
    return path, desc


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vks.npz"))


def test_container_tables_equal_the_reference_readers(yard, golden):
    path, _ = yard
    c = vks.read_vks_container(path)
    an = c["animation"]
    counts = [c["version"], len(c["meshes"]), len(c["instances"]), len(c["materials"]), c["n_triangles"], len(c["lod_groups"]), an["n_frames"],
              an["n_static"], an["n_animated"], an["offset"]]
    assert counts == golden["counts"][:10].tolist()
    for i, m in enumerate(c["meshes"]):
        ints = golden["mesh_ints"][i]
        nuv = m["vertex_buffer_offset"] + 24 * m["n_tris"]
        assert [len(m["segment_tris"]), m["n_tris"], m["material_id_base"], m["n_materials_in_range"], m["lod_group"], m["vertex_buffer_offset"], nuv,
                nuv + 24 * m["n_tris"], m["material_id_size"], m["flags"]] == ints.tolist()
        assert np.array_equal(np.float32(list(m["scale"]) + list(m["offset"])), golden["mesh_floats"][i])
        segs = golden["mesh_segs"][i].reshape(-1, 2)[:len(m["segment_tris"])]
        assert segs[:, 0].tolist() == m["segment_tris"] and segs[:, 1].tolist() == m["segment_material_base"]
        assert m["name"] == str(golden["mesh_names"][i])
    for i, inst in enumerate(c["instances"]):
        assert [inst["mesh_id"], inst["transform_index"], inst["flags"]] == golden["instances"][i].tolist()
    assert c["materials"] == [str(n) for n in golden["material_names"]]
    # the transform table: our dequantisation of every record == vkr_dequantize_transform, bit for bit
    tab = c["transform_table"]
    for k in range(an["n_static"]):
        rec = tab[24 * k:24 * k + 24]
        ours = scenes.vks_instance_transform(rec[:12].view(np.float32), rec[12:16].view(np.float32)[0], rec[16:24].view(np.uint16), flip=False)
        assert np.array_equal(ours.reshape(-1).view(np.uint32), golden["transforms"][k].view(np.uint32))
    assert [vks.transform_offset(i, 3, 5, f) for i, f in ((0, 0), (2, 7), (3, 0), (4, 2), (7, 3))] == golden["transform_offsets"].tolist()


def test_material_files_equal_the_reference_readers(yard, golden):
    path, _ = yard
    tdir = vks.texture_dir(path)
    for i, name in enumerate(vks_util.MATERIALS):
        f, tex = golden["material_floats"][i], golden["material_tex"][i].reshape(3, 7)
        for k, kind in enumerate(("BaseColor", "Normal", "Specular")):
            t = vks.read_vkt(os.path.join(tdir, "%s_%s.vkt" % (name, kind)))
            assert (t is not None) == bool(tex[k, 0])
            if t is not None:
                header = 4 * 6 + 8 + 24 * len(t["levels"])
                assert [t["width"], t["height"], t["format"], len(t["levels"]), len(t["blob"]), header] == tex[k, 1:].tolist()
    s = vks.load_vks(path)
    for i, m in enumerate(s.materials):
        f = golden["material_floats"][i]
        assert m.emission_intensity == f[0] and m.specular_transmission == f[4] and m.ior == f[5]
        if f[0] > 0:
            assert np.array_equal(np.float32(list(m.base_color)), f[1:4])
    # scene.cpp:836-1003: flags, handles, defaults
    wall, glass, lamp, leaf, plain = s.materials
    assert wall.flags & T.BASE_MATERIAL_NOALPHA and glass.flags & T.BASE_MATERIAL_NOALPHA       # BC1 RGB / default texture: no alpha
    assert not (leaf.flags & T.BASE_MATERIAL_NOALPHA) and not (plain.flags & T.BASE_MATERIAL_NOALPHA)   # BC3 / RGBA8
    assert glass.flags & T.BASE_MATERIAL_ONESIDED and not (leaf.flags & T.BASE_MATERIAL_ONESIDED)         # transmissive, name without "doublesided"
    assert len(s.textures) == 3 * len(s.materials) and all(m.normal_map == 3 * i + 1 for i, m in enumerate(s.materials))
    d = s.desc()
    assert [d.textures[k].bc_format for k in range(3)] == [1, 5, 1] and d.textures[0].mip_levels == 5
    assert d.textures[9].bc_format == 3 and d.textures[12].bc_format == 0 and d.textures[12].mip_levels == 4
    assert (d.textures[3].width, d.textures[3].height, d.textures[3].mip_levels) == (1, 1, 1)          # defaults are 1 x 1


def directly_built(desc):
    """The same scene through scenes.Scene, without the file: what load_vks must produce"""
    path_unused = None
    s = scenes.Scene()
    meshes, instances, (trans, scal, quats) = desc["meshes"], desc["instances"], desc["transforms"]
    for m in meshes:
        geoms, base = [], 0
        for n, _ in m["segments"]:
            if n:
                geoms.append(scenes.Geometry(m["qverts"][3 * base:3 * (base + n)], m["scale"], m["offset"], qnormal_uv=m["qnuv"][3 * base:3 * (base + n)],
                                             has_normals=True, has_uvs=True))
            base += n
        mid = s.add_mesh(geoms)
        if len(m["segments"]) == 1 and m["n_materials_in_range"] > 1:
            s.add_pmesh(mid, [m["material_id_base"]], m["material_ids"])
        else:
            s.add_pmesh(mid, [b for n, b in m["segments"] if n])
    for name, mesh_id, tidx in instances:
        if mesh_id == 2:
            continue   # LoD level 1 of the tree: not a base level
        rec = np.frombuffer(vks.quantize_transform(trans[tidx], scal[tidx], quats[tidx]), np.uint8)
        s.add_instance(mesh_id, scenes.vks_instance_transform(rec[:12].view(np.float32), rec[12:16].view(np.float32)[0], rec[16:24].view(np.uint16)))
    return s


def test_loaded_scene_renders_like_the_directly_built_one(yard, oracle):
    path, desc = yard
    s = vks.load_vks(path)
    ref = directly_built(desc)
    ref.materials, ref.textures = s.materials, s.textures   # materials / textures come from the files in both cases (checked above)
    assert s.total_tris() == ref.total_tris() == 2 * 60 + 2 * 50
    cam = scenes.look_at_camera((0, 2, 14), (0, 0, 0), fovy=50.0)
    sp = load_sky_fit(T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)))
    W, H = 96, 54
    a, _ = oracle.OracleScene(s).render(W, H, cam, sp, spp=2, transmission=1)
    b, _ = oracle.OracleScene(ref).render(W, H, cam, sp, spp=2, transmission=1)
    assert np.isfinite(a).all() and a[..., :3].max() > 0 and (a[..., 3] > 0).mean() > 0.05
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # the emitter of the parameter file lights the scene: without it the image is darker
    dark = vks.load_vks(path)
    dark.materials[2].emission_intensity = 0.0
    c, _ = oracle.OracleScene(dark).render(W, H, cam, sp, spp=2, transmission=1)
    assert c[..., :3].sum() < a[..., :3].sum()


def test_malformed_files_are_refused(yard, tmp_path):
    path, _ = yard
    data = open(path, "rb").read()
    bad = tmp_path / "bad.vks"
    for mutate in (lambda b: b"\0\0\0\0" + b[4:], lambda b: b[:4] + (9).to_bytes(4, "little") + b[8:], lambda b: b[:200],
                   lambda b: b[:16] + (17).to_bytes(8, "little") + b[24:]):
        bad.write_bytes(mutate(data))
        with pytest.raises(vks.VksError):
            vks.read_vks_container(str(bad))
    (tmp_path / "x.vkt").write_bytes(b"\1\2\3\4" * 10)
    with pytest.raises(vks.VksError):
        vks.read_vkt(str(tmp_path / "x.vkt"))
    assert vks.read_vkt(str(tmp_path / "missing.vkt")) is None


def test_reference_reader_reproduces_its_fixture(yard, golden):
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    R = C.CDLL(so)
    if not hasattr(R, "ref_vkr_scene_counts"):
        pytest.skip("oracle/_ref predates ref_vkr.c")
    path, _ = yard
    out = np.zeros(12, np.int64)
    assert R.ref_vkr_scene_counts(path.encode(), out.ctypes.data_as(C.POINTER(C.c_int64))) == 0
    assert out.tolist() == golden["counts"].tolist()
