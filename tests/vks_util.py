"""A small .vks scene with its texture directory, written by realtimepathtracingresearchframework_b200/vks.py: shared by
tests/test_vks.py and oracle/gen_golden.py --vks-only (which asks the reference's own reader, ext/libvkr/src/vkr.c, what it makes of it)."""
import os

import numpy as np

from realtimepathtracingresearchframework_b200 import scenes, types as T, vks

MATERIALS = ["wall", "glass", "lamp", "leaf_doublesided", "plain"]


def mesh_streams(n, seed, with_uv=True):
    """n triangles on the 21-bit grid + their normal / uv stream"""
    g = scenes.random_triangle_grid(n, seed, box=2.0, edge=0.5, scale=2.0 ** -16, base=-4.0)
    qv = scenes.pack_qverts(g.reshape(-1, 3))
    rng = np.random.default_rng(seed + 1)
    nrm = rng.normal(size=(3 * n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = rng.uniform(0, 2, (3 * n, 2)).astype(np.float32)
    return qv, scenes.pack_qnormal_uv(nrm, uv)


def write_test_scene(directory):
    """Returns (path of the .vks file, description dict used to build the same scene directly)."""
    path = os.path.join(directory, "yard.vks")
    tdir = vks.texture_dir(path)
    os.makedirs(tdir, exist_ok=True)
    qa, na = mesh_streams(60, 11)
    qb, nb = mesh_streams(50, 12)
    qc, nc = mesh_streams(20, 13)
    ids_a = (np.arange(60) % 3).astype(np.uint8)
    off = (-4.0 + 2.0 ** -17,) * 3
    meshes = [
        dict(name="terrace", scale=(2.0 ** -16,) * 3, offset=off, qverts=qa, qnuv=na, segments=[(60, 0)], material_ids=ids_a, material_id_base=0,
             n_materials_in_range=3, lod_group=0),
        dict(name="tree", scale=(2.0 ** -16,) * 3, offset=off, qverts=qb, qnuv=nb, segments=[(30, 3), (0, 1), (20, 4)], material_id_base=3,
             n_materials_in_range=2, lod_group=1),
        dict(name="tree_lod1", scale=(2.0 ** -16,) * 3, offset=off, qverts=qc, qnuv=nc, segments=[(20, 3)], material_id_base=3,
             n_materials_in_range=1, lod_group=1),
    ]
    quats = [(0.0, 0.0, 0.0, 1.0), (0.0, 0.3826834, 0.0, 0.9238795), (0.5, 0.5, 0.5, 0.5), (0.0, 0.0, 0.7071068, 0.7071068)]
    trans = [(0.0, 0.0, 0.0), (3.0, 0.5, -1.0), (-2.5, 1.0, 2.0), (0.0, -1.0, 0.0)]
    scal = [1.0, 0.75, 1.25, 2.0]
    transforms = [vks.quantize_transform(t, s, q) for t, s, q in zip(trans, scal, quats)]
    instances = [("terrace0", 0, 0), ("terrace1", 0, 1), ("tree0", 1, 2), ("tree0_lod1", 2, 2), ("tree1", 1, 3)]
    vks.write_vks(path, meshes, instances, MATERIALS, transforms, lod_groups=[[(1, 0.0), (2, 0.6)]])
    # texture directory: colour (BC3 with alpha / BC1 / raw), normal (BC5), specular-roughness-metalness (BC1); some materials have none
    col = scenes.procedural_image(32, 16, 4, 21)
    yy, xx = np.mgrid[0:16, 0:32]
    col[..., 3] = np.where((xx // 4 + yy // 4) % 2 == 0, 255, 60).astype(np.uint8)
    vks.write_vkt(os.path.join(tdir, "wall_BaseColor.vkt"), scenes.mip_chain(scenes.procedural_image(16, 16, 3, 22)), 132)
    vks.write_vkt(os.path.join(tdir, "leaf_doublesided_BaseColor.vkt"), scenes.mip_chain(col), 138)
    vks.write_vkt(os.path.join(tdir, "plain_BaseColor.vkt"), scenes.mip_chain(scenes.procedural_image(8, 8, 4, 23)), 37)
    nrm = scenes.procedural_image(16, 16, 2, 24).astype(np.int32)
    nrm = np.stack([128 + (nrm[..., 0] - 128) // 3, 128 + (nrm[..., 1] - 128) // 3], -1).astype(np.uint8)
    vks.write_vkt(os.path.join(tdir, "wall_Normal.vkt"), scenes.mip_chain(nrm), 141)
    vks.write_vkt(os.path.join(tdir, "wall_Specular.vkt"), scenes.mip_chain(scenes.procedural_image(16, 8, 3, 25)), 131)
    open(os.path.join(tdir, "lamp_EmissionIntensity.txt"), "w").write("12.5\n1.0\n0.8\n0.6\n")
    open(os.path.join(tdir, "glass_SpecularTransmission.txt"), "w").write("0.9\n1.45\n")
    open(os.path.join(tdir, "leaf_doublesided_SpecularTransmission.txt"), "w").write("0.3\n1.3\n0.0\n0.2\n")
    open(os.path.join(tdir, "glass_Ex.txt"), "w").write("glass_SHADERMATERIAL_thin")
    return path, dict(meshes=meshes, instances=instances, transforms=(trans, scal, quats))
