"""Inputs shared by tests/test_texture_lod.py and oracle/gen_golden.py --textures-only: a small texture set with mip chains (raw and
block-compressed) and random materials that refer to it."""
import ctypes as C

import numpy as np

from realtimepathtracingresearchframework_b200 import scenes, types as T


def texture_set():
    """[(levels, color_space, bc_format)]: sRGB RGBA with alpha holes, linear RGB, BC1 / BC3 / BC5 compressed, a non-square odd-sized one"""
    a = scenes.procedural_image(32, 16, 4, 3)
    yy, xx = np.mgrid[0:16, 0:32]
    a[..., 3] = np.where(((xx // 4 + yy // 4) % 3) == 0, 0, np.where(((xx // 4 + yy // 4) % 3) == 1, 150, 255)).astype(np.uint8)
    specs = [(a, T.COLOR_SPACE_SRGB, 0), (scenes.procedural_image(16, 16, 3, 4), T.COLOR_SPACE_LINEAR, 0),
             (scenes.procedural_image(24, 12, 3, 5), T.COLOR_SPACE_LINEAR, 1), (a.copy(), T.COLOR_SPACE_SRGB, 3),
             (scenes.procedural_image(16, 16, 2, 6), T.COLOR_SPACE_LINEAR, 5), (scenes.procedural_image(13, 7, 4, 7), T.COLOR_SPACE_SRGB, -1)]
    return [(scenes.mip_chain(px), cs, bc) for px, cs, bc in specs]


def texture_descs(tset):
    """ctypes array of rptr_texture_desc for texture_set() (+ the arrays that must stay alive)"""
    descs = (T.TextureDesc * len(tset))()
    keep = []
    for i, (levels, cs, bc) in enumerate(tset):
        blob = np.concatenate([(scenes.encode_bc(l, bc) if bc else np.ascontiguousarray(l, np.uint8)).reshape(-1) for l in levels])
        h, w, ch = levels[0].shape
        descs[i].width, descs[i].height, descs[i].channels, descs[i].color_space = w, h, ch, cs
        descs[i].bc_format, descs[i].mip_levels = bc, len(levels)
        descs[i].texels = blob.ctypes.data_as(C.POINTER(C.c_uint8))
        keep.append(blob)
    return descs, keep


def random_textured_materials(n, n_textures, seed):
    """BaseMaterials whose parameters are constants or handles into the texture set, with random uv and duvdxy per case"""
    rng = np.random.default_rng(seed)
    mats = []
    for _ in range(n):
        m = T.BaseMaterial(base_color=tuple(rng.uniform(0, 1, 3)), roughness=float(rng.uniform(0.05, 1)), metallic=float(rng.uniform(0, 1)),
                           specular=float(rng.uniform(0, 1)), ior=float(rng.choice([1.0, 1.5])), specular_transmission=float(rng.choice([0.0, 0.6])),
                           flags=0)
        if rng.uniform() < 0.8:
            m.base_color = (T.texture_handle(int(rng.integers(n_textures))), 0.0, 0.0)
        for f in ("specular", "roughness", "metallic", "specular_transmission"):
            if rng.uniform() < 0.5:
                setattr(m, f, T.texture_handle(int(rng.integers(n_textures)), int(rng.integers(4))))
        mats.append(m)
    uv = rng.uniform(-2, 3, (n, 2)).astype(np.float32)
    scale = 10.0 ** rng.uniform(-4, -0.5, (n, 1))
    duvdxy = (rng.normal(size=(n, 4)) * scale).astype(np.float32)
    duvdxy[::11] = 0.0   # zero footprint: base level
    return mats, uv, duvdxy


def footprint_cases(n, seed):
    """inputs of the three footprint operations (hostsim_footprint_op layouts)"""
    rng = np.random.default_rng(seed)
    def unit(v):
        return v / np.linalg.norm(v, axis=-1, keepdims=True)
    d = unit(rng.normal(size=(n, 3)))
    d[::17] = np.eye(3)[rng.integers(3, size=len(d[::17]))] * rng.choice([-1.0, 1.0], size=(len(d[::17]), 1))  # axis-aligned directions
    s = 10.0 ** rng.uniform(-5, -1, (n, 1))
    op0 = np.concatenate([d, rng.normal(size=(n, 3)) * s, rng.normal(size=(n, 3)) * s], 1).astype(np.float32)
    d2 = unit(rng.normal(size=(n, 3))).astype(np.float32)
    return op0, d2
