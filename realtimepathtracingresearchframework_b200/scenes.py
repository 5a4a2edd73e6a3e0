"""Procedural scenes for BASELINE.json's configs, built directly as the reference's in-memory
scene model: quantised unrolled vertex streams per Geometry, Mesh / ParameterizedMesh / Instance tables and
BaseMaterial blocks (librender/mesh.h:10-129, librender/scene.h:48-72).

Positions are generated ON the 21-bit grid with a power-of-two quantized_scaling so that the float vertices the BVH
sees and the ones the shader re-dequantises (rendering/rt/hit.glsl:41-46) are the same numbers (hard part 5).
"""
import ctypes as C

import numpy as np

from . import types as T

_MASK21 = np.uint64(0x1FFFFF)


def pack_qverts(grid):
    """grid: (..., 3) integer array of 21-bit coordinates -> uint64 words (librender/quantize.h:7-11)."""
    g = np.asarray(grid).astype(np.uint64)
    return (g[..., 0] & _MASK21) | ((g[..., 1] & _MASK21) << np.uint64(21)) | ((g[..., 2] & _MASK21) << np.uint64(42))


def quantize_normal(n):
    """float normals (..., 3) -> uint32 oct encoding (librender/quantize.h:21-35), float32 arithmetic."""
    n = np.asarray(n, dtype=np.float32)
    nl1 = np.abs(n[..., 0]) + np.abs(n[..., 1]) + np.abs(n[..., 2])
    pn = n[..., :2] / nl1[..., None]
    neg = n[..., 2] <= 0
    sx = np.where(pn[..., 0] >= 0, np.float32(1), np.float32(-1))
    sy = np.where(pn[..., 1] >= 0, np.float32(1), np.float32(-1))
    fx = (np.float32(1) - np.abs(pn[..., 1])) * sx
    fy = (np.float32(1) - np.abs(pn[..., 0])) * sy
    px = np.where(neg, fx, pn[..., 0]) * np.float32(0x8000)
    py = np.where(neg, fy, pn[..., 1]) * np.float32(0x8000)
    ix = np.clip(px.astype(np.int32), -0x7FFF, 0x7FFF)
    iy = np.clip(py.astype(np.int32), -0x7FFF, 0x7FFF)
    ux = (ix + 0x8000).astype(np.uint32)
    uy = (iy + 0x8000).astype(np.uint32)
    return ux | (uy << np.uint32(16))


def quantize_uv(uv):
    """uv (..., 2) -> uint32 (librender/quantize.h:38-42 with a zero safety offset)."""
    uv = np.asarray(uv, dtype=np.float32)
    s = np.float32(0xFFFF) / np.float32(8.0)
    a = (np.float32(0) + uv[..., 0]) * s
    b = ((np.float32(1) + np.float32(0)) - uv[..., 1]) * s
    ux = (np.float32(0.5) + a).astype(np.int64).astype(np.uint32) & np.uint32(0xFFFF)
    uy = (np.float32(0.5) + b).astype(np.int64).astype(np.uint32) & np.uint32(0xFFFF)
    return ux | (uy << np.uint32(16))


class Geometry:
    def __init__(self, qverts, scaling, offset, qnormal_uv=None, has_normals=False, has_uvs=False):
        self.qverts = np.ascontiguousarray(qverts, dtype=np.uint64).reshape(-1)
        assert self.qverts.size % 3 == 0
        self.qnormal_uv = None if qnormal_uv is None else np.ascontiguousarray(qnormal_uv, dtype=np.uint64).reshape(-1)
        self.scaling = tuple(float(x) for x in scaling)
        self.offset = tuple(float(x) for x in offset)
        self.has_normals, self.has_uvs = bool(has_normals), bool(has_uvs)

    @property
    def n_tris(self):
        return self.qverts.size // 3

    def positions(self):
        """Dequantised float32 vertices, (n_tris, 3, 3) (librender/dequantize.glsl:8-21)."""
        q = self.qverts
        u = np.stack([q & _MASK21, (q >> np.uint64(21)) & _MASK21, (q >> np.uint64(42)) & _MASK21], -1).astype(np.float32)
        p = u * np.asarray(self.scaling, np.float32) + np.asarray(self.offset, np.float32)
        return p.reshape(-1, 3, 3)


def mip_chain(img):
    """All mip levels of an (h, w, c) uint8 image down to 1 x 1: level k + 1 = (max(w // 2, 1), max(h // 2, 1)) of level k
    (util/image.cpp:11-21), each texel the rounded mean of the texels it covers (a 2 x 2 box for even sizes)."""
    levels = [np.ascontiguousarray(img, np.uint8)]
    while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:
        a = levels[-1].astype(np.float64)
        h, w = a.shape[:2]
        nh, nw = max(h // 2, 1), max(w // 2, 1)
        ys = (np.arange(nh + 1) * h) // nh
        xs = (np.arange(nw + 1) * w) // nw
        out = np.zeros((nh, nw, a.shape[2]))
        for j in range(nh):
            for i in range(nw):
                out[j, i] = a[ys[j]:ys[j + 1], xs[i]:xs[i + 1]].mean((0, 1))
        levels.append(np.floor(out + 0.5).astype(np.uint8))
    return levels


def _bc1_block(tile, rgba_alpha):
    """One 4 x 4 RGBA tile -> 8 bytes of BC1: endpoints = the tile's per-channel min / max quantised to 565, nearest palette entry per
    texel.  rgba_alpha: three-colour mode with a transparent index for texels with alpha < 128 (BC1 RGBA)."""
    px = tile.reshape(16, 4).astype(np.int32)
    opaque = px[:, 3] >= 128 if rgba_alpha else np.ones(16, bool)
    src = px[opaque][:, :3] if opaque.any() else px[:, :3]
    lo, hi = src.min(0), src.max(0)
    def q565(c):
        return ((int(c[0]) * 31 + 127) // 255) << 11 | ((int(c[1]) * 63 + 127) // 255) << 5 | ((int(c[2]) * 31 + 127) // 255)
    def expand(v):
        r, g, b = v >> 11, (v >> 5) & 63, v & 31
        return np.array([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], np.int32)
    c0, c1 = q565(hi), q565(lo)
    punch = rgba_alpha and not opaque.all()
    if punch:
        if c0 > c1:
            c0, c1 = c1, c0
    else:
        if c0 < c1:
            c0, c1 = c1, c0
        if c0 == c1:  # four-colour mode needs c0 > c1; with equal endpoints every index decodes to (nearly) the same colour anyway
            if c0 == 0:
                c0 = 1
            else:
                c1 = c0 - 1
    e0, e1 = expand(c0), expand(c1)
    if c0 > c1:
        pal = [e0, e1, (2 * e0 + e1 + 1) // 3, (e0 + 2 * e1 + 1) // 3]
    else:
        pal = [e0, e1, (e0 + e1 + 1) // 2, np.zeros(3, np.int32)]
    idx = 0
    for i in range(16):
        if punch and not opaque[i]:
            k = 3
        else:
            cand = pal if c0 > c1 else pal[:3]
            k = int(np.argmin([((px[i, :3] - p) ** 2).sum() for p in cand]))
        idx |= k << (2 * i)
    return bytes([c0 & 255, c0 >> 8, c1 & 255, c1 >> 8]) + int(idx).to_bytes(4, "little")


def _bc4_block(vals):
    """16 values -> 8 bytes of BC4 UNORM (eight-value mode: endpoints max / min, nearest palette entry)."""
    v = vals.reshape(16).astype(np.int32)
    a, b = int(v.max()), int(v.min())
    if a == b:
        pal = [a, b, 0, 0, 0, 0, 0, 255] if a <= b else None
        pal = [a, b] + [((5 - i) * a + i * b + 2) // 5 for i in range(1, 5)] + [0, 255]
    else:
        pal = [a, b] + [((7 - i) * a + i * b + 3) // 7 for i in range(1, 7)]
    idx = 0
    for i in range(16):
        k = int(np.argmin([abs(int(v[i]) - p) for p in pal]))
        idx |= k << (3 * i)
    return bytes([a, b]) + int(idx).to_bytes(6, "little")


def encode_bc(img, bc_format):
    """Block-compresses one mip level ((h, w, c) uint8, c expanded to RGBA) the way a .vks exporter would: 4 x 4 blocks, the level
    padded to whole blocks by repeating its last row / column.  A simple encoder (endpoints = min / max), enough to produce valid
    streams for the decoders."""
    px = np.ascontiguousarray(img, np.uint8)
    h, w, c = px.shape
    rgba = np.zeros((h, w, 4), np.uint8)
    rgba[..., 3] = 255
    rgba[..., :c] = px
    H, W = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    pad = np.pad(rgba, ((0, H - h), (0, W - w), (0, 0)), mode="edge")
    out = bytearray()
    for by in range(0, H, 4):
        for bx in range(0, W, 4):
            tile = pad[by:by + 4, bx:bx + 4]
            if bc_format == 1:
                out += _bc1_block(tile, False)
            elif bc_format == -1:
                out += _bc1_block(tile, True)
            elif bc_format == 3:
                out += _bc4_block(tile[..., 3]) + _bc1_block(tile, False)
            elif bc_format == 5:
                out += _bc4_block(tile[..., 0]) + _bc4_block(tile[..., 1])
            else:
                raise ValueError("bc_format %r" % bc_format)
    return np.frombuffer(bytes(out), np.uint8)


class Scene:
    """In-memory scene; `desc()` yields the rptr_scene_desc consumed by rptr_cuda_set_scene (and by the oracle)."""

    def __init__(self):
        self.geometries = []   # Geometry
        self.meshes = []       # (first_geometry, n_geometries)
        self.pmeshes = []      # dict(mesh_id, material_offsets, tri_material_ids)
        self.instances = []    # (pmesh_id, 3x4 row-major transform)
        self.materials = []    # T.BaseMaterial
        self.binned_lights = None
        self.textures = []     # (texels uint8[h, w, channels], T.COLOR_SPACE_*): single-level images of any size
        self._keep = []

    def add_texture(self, texels, color_space=T.COLOR_SPACE_LINEAR, mips=False, bc_format=0):
        """A texture from one texel (1-4 channels, 8 bit) or an (h, w, channels) image; returns the texture id for T.texture_handle().
        mips=True appends a box-filtered mip chain down to 1 x 1 (util/image.h: all levels of an Image back to back); bc_format != 0
        block-compresses every level (BC1 RGB 1, BC1 RGBA -1, BC3 3, BC5 5: the formats of a .vks scene) with encode_bc()."""
        px = np.asarray(texels, np.uint8)
        px = px.reshape(1, 1, -1) if px.ndim == 1 else (px[..., None] if px.ndim == 2 else px)
        px = np.ascontiguousarray(px)
        # entries: (base level, colour space) -- or, with a mip chain / block compression, (base level, colour space, levels, bc_format)
        self.textures.append((px, color_space, mip_chain(px), bc_format) if (mips or bc_format) else (px, color_space))
        return len(self.textures) - 1

    def add_mesh(self, geometries):
        first = len(self.geometries)
        self.geometries.extend(geometries)
        self.meshes.append((first, len(geometries)))
        return len(self.meshes) - 1

    def add_pmesh(self, mesh_id, material_offsets, tri_material_ids=None):
        tm = None if tri_material_ids is None else np.ascontiguousarray(tri_material_ids, dtype=np.uint8)
        self.pmeshes.append(dict(mesh_id=mesh_id, material_offsets=np.ascontiguousarray(material_offsets, dtype=np.int32),
                                 tri_material_ids=tm))
        return len(self.pmeshes) - 1

    def add_instance(self, pmesh_id, transform=None):
        t = np.eye(4, dtype=np.float32)[:3] if transform is None else np.asarray(transform, dtype=np.float32).reshape(3, 4)
        self.instances.append((pmesh_id, np.ascontiguousarray(t)))
        return len(self.instances) - 1

    def total_tris(self):
        n = 0
        for pm_id, _ in self.instances:
            first, cnt = self.meshes[self.pmeshes[pm_id]["mesh_id"]]
            n += sum(g.n_tris for g in self.geometries[first:first + cnt])
        return n

    def desc(self):
        keep = []
        geoms = (T.GeometryDesc * max(1, len(self.geometries)))()
        for i, g in enumerate(self.geometries):
            geoms[i].qverts = g.qverts.ctypes.data
            geoms[i].qnormal_uv = None if g.qnormal_uv is None else g.qnormal_uv.ctypes.data
            for k in range(3):
                geoms[i].quantized_scaling[k] = g.scaling[k]
                geoms[i].quantized_offset[k] = g.offset[k]
            geoms[i].n_tris = g.n_tris
            geoms[i].has_normals = int(g.has_normals)
            geoms[i].has_uvs = int(g.has_uvs)
        meshes = (T.MeshDesc * max(1, len(self.meshes)))()
        for i, (f, n) in enumerate(self.meshes):
            meshes[i].first_geometry, meshes[i].n_geometries = f, n
        pms = (T.PMeshDesc * max(1, len(self.pmeshes)))()
        for i, pm in enumerate(self.pmeshes):
            pms[i].mesh_id = pm["mesh_id"]
            pms[i].n_material_offsets = pm["material_offsets"].size
            pms[i].material_offsets = pm["material_offsets"].ctypes.data
            tm = pm["tri_material_ids"]
            pms[i].tri_material_ids = None if tm is None else tm.ctypes.data
            pms[i].n_tri_material_ids = 0 if tm is None else tm.size
        insts = (T.InstanceDesc * max(1, len(self.instances)))()
        for i, (pm, t) in enumerate(self.instances):
            insts[i].pmesh_id = pm
            for k, x in enumerate(t.reshape(-1)):
                insts[i].transform[k] = float(x)
        mats = (T.BaseMaterial * max(1, len(self.materials)))(*self.materials)
        d = T.SceneDesc()
        d.geometries, d.n_geometries = geoms, len(self.geometries)
        d.meshes, d.n_meshes = meshes, len(self.meshes)
        d.pmeshes, d.n_pmeshes = pms, len(self.pmeshes)
        d.instances, d.n_instances = insts, len(self.instances)
        d.materials, d.n_materials = mats, len(self.materials)
        if self.binned_lights is not None:
            bl = (T.TriLightData * len(self.binned_lights))(*self.binned_lights)
            d.binned_lights, d.n_binned_lights = bl, len(self.binned_lights)
            keep.append(bl)
        if getattr(self, "textures", None):
            texs = (T.TextureDesc * len(self.textures))()
            for i, t in enumerate(self.textures):
                if len(t) == 4 and isinstance(t[0], str):  # ("vkt", colour space, vks.read_vkt() record, bcFormat): the file's bytes as they are
                    blob, bc, w, h, ch, n_levels = np.ascontiguousarray(t[2]["blob"]), t[3], t[2]["width"], t[2]["height"], 4, len(t[2]["levels"])
                else:
                    levels, bc = (t[2], t[3]) if len(t) == 4 else ([np.ascontiguousarray(t[0], np.uint8)], 0)
                    if len(t) == 4 and not np.array_equal(levels[0], t[0]):
                        levels = mip_chain(t[0])  # the base level was edited after add_texture
                    h, w, ch = levels[0].shape
                    n_levels = len(levels)
                    blob = np.concatenate([(encode_bc(l, bc) if bc else np.ascontiguousarray(l, np.uint8)).reshape(-1) for l in levels])
                texs[i].width, texs[i].height, texs[i].channels, texs[i].color_space = w, h, ch, t[1]
                texs[i].bc_format, texs[i].mip_levels = bc, n_levels
                texs[i].texels = blob.ctypes.data_as(C.POINTER(C.c_uint8))
                keep.append(blob)
            d.textures, d.n_textures = texs, len(self.textures)
            keep.append(texs)
        keep += [geoms, meshes, pms, insts, mats]
        self._keep = keep  # ctypes arrays must outlive the descriptor
        return d


def look_at_camera(eye, center, up=(0.0, 1.0, 0.0), fovy=65.0):
    """RenderCameraParams{eye, normalised view direction, up, fovy} as app.cpp:357 hands it to the backend."""
    eye = np.asarray(eye, np.float32)
    d = np.asarray(center, np.float32) - eye
    d = d / np.float32(np.sqrt(np.float32(np.dot(d, d))))
    return T.RenderCameraParams(pos=tuple(eye), dir=tuple(d), up=tuple(up), fovy=fovy)


# ---------------------------------------------------------------------------------------------------------------
# C1: Cornell-style box, 12 triangles (SURVEY 8d "C1 synthetic input")
# ---------------------------------------------------------------------------------------------------------------
def _quad(a, b, c, d):
    """two triangles (a,b,c), (a,c,d)"""
    return [a, b, c, a, c, d]


def cornell_box():
    s = Scene()
    base = np.array([-1.0, 0.0, -1.0])
    scale = 2.0 ** -20

    def grid(points):
        g = np.floor((np.asarray(points, np.float64) - base) / scale)
        return np.clip(g, 0, 0x1FFFFF).astype(np.int64)

    offset = tuple(base + 2.0 ** -21)
    floor = _quad((-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1))
    ceil = _quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1))
    back = _quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1))
    left = _quad((-1, 0, -1), (-1, 2, -1), (-1, 2, 1), (-1, 0, 1))
    right = _quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1))
    # light quad just under the ceiling, wound so that cross(v1-v0, v2-v0) points down (-y): tri lights only
    # illuminate points they face (is_tri_facing_forward, rendering/lights/tri.glsl:23-25)
    light = _quad((-0.25, 1.998, -0.25), (0.25, 1.998, -0.25), (0.25, 1.998, 0.25), (-0.25, 1.998, 0.25))
    groups = [floor + ceil + back, left, right, light]
    geoms = [Geometry(pack_qverts(grid(g)), (scale,) * 3, offset) for g in groups]
    mesh = s.add_mesh(geoms)
    noalpha = T.BASE_MATERIAL_NOALPHA
    lam = dict(ior=1.0, roughness=1.0, metallic=0.0, flags=noalpha)
    s.materials = [T.BaseMaterial(base_color=(0.73, 0.73, 0.73), **lam), T.BaseMaterial(base_color=(0.63, 0.065, 0.05), **lam),
                   T.BaseMaterial(base_color=(0.14, 0.45, 0.091), **lam),
                   T.BaseMaterial(base_color=(1.0, 0.85, 0.6), emission_intensity=17.0, **lam)]
    pm = s.add_pmesh(mesh, [0, 1, 2, 3])
    s.add_instance(pm)
    s.camera = look_at_camera((0, 1, 3.4), (0, 1, 0), fovy=40.0)
    s.name = "cornell12"
    return s


# ---------------------------------------------------------------------------------------------------------------
# C2/C3/C5: N random triangles (default 1 000 000) in 16 geometries, diffuse + GGX, sun + sky only
# ---------------------------------------------------------------------------------------------------------------
_PALETTE = [(0.80, 0.80, 0.80), (0.90, 0.35, 0.25), (0.25, 0.65, 0.90), (0.95, 0.85, 0.35), (0.35, 0.80, 0.45), (0.75, 0.40, 0.85),
            (0.95, 0.60, 0.20), (0.30, 0.35, 0.85), (0.60, 0.60, 0.60), (0.85, 0.25, 0.45), (0.20, 0.75, 0.75), (0.70, 0.75, 0.30),
            (0.90, 0.90, 0.95), (0.55, 0.35, 0.25), (0.40, 0.55, 0.35), (0.50, 0.50, 0.70)]


def splitmix64_uniform(seed, n):
    """n doubles in [0,1): the i-th value is the splitmix64 output for state seed + (i+1)*golden, top 53 bits."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def random_triangle_grid(n_tris, seed=0x5EED1A7B200, box=10.0, edge=0.15, scale=2.0 ** -16, base=-16.0):
    u = splitmix64_uniform(seed, n_tris * 9).reshape(n_tris, 9)
    c = (u[:, 0:3] * 2.0 - 1.0) * box
    e1 = (u[:, 3:6] * 2.0 - 1.0) * edge
    e2 = (u[:, 6:9] * 2.0 - 1.0) * edge
    v = np.stack([c, c + e1, c + e2], 1)
    return np.clip(np.floor((v - base) / scale), 0, 0x1FFFFF).astype(np.int64)  # (n, 3, 3)


def random_triangles(n_tris=1_000_000, n_geometries=16, seed=0x5EED1A7B200, box=10.0, edge=0.15):
    s = Scene()
    scale, base = 2.0 ** -16, -16.0
    g = random_triangle_grid(n_tris, seed, box=box, edge=edge, scale=scale, base=base)
    per = (n_tris + n_geometries - 1) // n_geometries
    offset = (base + 2.0 ** -17,) * 3
    geoms = []
    for j in range(n_geometries):
        part = g[j * per:(j + 1) * per]
        if len(part) == 0:
            break
        geoms.append(Geometry(pack_qverts(part.reshape(-1, 3)), (scale,) * 3, offset))
    mesh = s.add_mesh(geoms)
    s.materials = [T.BaseMaterial(base_color=_PALETTE[j % 16], roughness=0.1 + 0.05 * (j % 16), metallic=float(j & 1), ior=1.5,
                                  specular=0.5, flags=T.BASE_MATERIAL_NOALPHA) for j in range(len(geoms))]
    pm = s.add_pmesh(mesh, list(range(len(geoms))))
    s.add_instance(pm)
    s.camera = look_at_camera((0, 0, 30), (0, 0, 0), fovy=65.0)
    s.name = "random%d" % n_tris
    return s


def alpha_tested_soup(n_tris=6000, seed=99):
    """Random triangles whose materials exercise the 1x1-texel mode and the stochastic alpha candidate
    filter (8a-4): cut-out (alpha 0), half transparent (alpha 128/255), faint (alpha 32/255), an opaque sRGB texel, a
    material that reads roughness / metallic from texture channels, alpha-textured but flagged NOALPHA, constants, and
    two materials with a (one-texel) normal map."""
    s = random_triangles(n_tris, n_geometries=8, seed=seed, box=5.0, edge=0.9)
    t_half = s.add_texture((200, 120, 60, 128), T.COLOR_SPACE_SRGB)
    t_cut = s.add_texture((255, 255, 255, 0), T.COLOR_SPACE_SRGB)
    t_faint = s.add_texture((40, 180, 220, 32), T.COLOR_SPACE_SRGB)
    t_srgb = s.add_texture((188, 64, 230), T.COLOR_SPACE_SRGB)           # 3 channels: alpha reads 1
    t_orm = s.add_texture((255, 90, 200, 255), T.COLOR_SPACE_LINEAR)     # (specular, roughness, metallic) like the glTF ORM slot
    m = s.materials
    m[0].base_color, m[0].flags = (T.texture_handle(t_half), 0.0, 0.0), 0
    m[1].base_color, m[1].flags = (T.texture_handle(t_cut), 0.0, 0.0), 0
    m[2].base_color, m[2].flags = (T.texture_handle(t_faint), 0.0, 0.0), 0
    m[3].base_color, m[3].flags = (T.texture_handle(t_srgb), 0.0, 0.0), 0
    m[4].roughness, m[4].metallic, m[4].flags = T.texture_handle(t_orm, 1), T.texture_handle(t_orm, 2), 0
    m[5].base_color = (T.texture_handle(t_half), 0.0, 0.0)               # NOALPHA stays set: colour from the texel, never alpha-tested
    m[6].flags = 0                                                        # constants without NOALPHA: alpha 1, no draw
    t_nrm = s.add_texture((170, 96, 255), T.COLOR_SPACE_LINEAR)          # tilted tangent-space normal
    m[7].normal_map = t_nrm                                               # normal map (pt_megakernel.glsl:634-654), one texel
    m[3].normal_map = t_nrm
    s.camera = look_at_camera((0, 0, 16), (0, 0, 0), fovy=55.0)
    s.name = "alpha_soup%d" % n_tris
    return s


# ---------------------------------------------------------------------------------------------------------------------
# C4: one base mesh instanced many times; GGX + transmission (thick / thin) + emissive per-triangle materials
# ---------------------------------------------------------------------------------------------------------------------
def vks_flip(m43):
    """The vks axis flip of AnimationData::dequantize (librender/scene.cpp:36-40) applied to a float matrix[4][3] of vkr.c (column i of the
    glm::mat4x3 = matrix[i]): 3 x 4 row-major object-to-world matrix."""
    M = np.zeros((3, 4), np.float32)
    for c in range(4):
        M[:, c] = np.asarray(m43, np.float32)[c]
    flip = np.array([[-1, 0, 0], [0, 0, 1], [0, 1, 0]], np.float32)  # (x, y, z) -> (-x, z, y)
    return (flip @ M).astype(np.float32)


def vks_instance_transform(translation, scaling, quat_codes, flip=True):
    """The 3x4 object-to-world matrix a .vks instance yields: vkr_dequantize_transform of the 24-byte record
    (translation 3 x f32, scaling f32, quaternion 4 x u16; ext/libvkr/src/vkr.c:1381-1408) followed by the vks axis
    flip of AnimationData::dequantize (librender/scene.cpp:22-41).  float32 arithmetic throughout."""
    f = np.float32
    q = (np.asarray(quat_codes, np.uint16).astype(np.float32) * (f(2.0) / f(0xffff)) - f(1.0)).astype(np.float32)
    q[3] = -q[3]
    xx, xy, xz, xw = q[0] * q[0], q[0] * q[1], q[0] * q[2], q[0] * q[3]
    yy, yz, yw = q[1] * q[1], q[1] * q[2], q[1] * q[3]
    zz, zw = q[2] * q[2], q[2] * q[3]
    m = np.zeros((4, 3), np.float32)  # float matrix[4][3] of vkr.c
    m[0] = [f(1) - f(2) * (yy + zz), f(2) * (xy - zw), f(2) * (xz + yw)]
    m[1] = [f(2) * (xy + zw), f(1) - f(2) * (xx + zz), f(2) * (yz - xw)]
    m[2] = [f(2) * (xz - yw), f(2) * (yz + xw), f(1) - f(2) * (xx + yy)]
    m[:3] *= f(scaling)
    m[3] = np.asarray(translation, np.float32)
    return vks_flip(m) if flip else m


def quantize_quaternion(q):
    """vkr_quantize_transform's 16-bit quaternion code (ext/libvkr/src/vkr.c:1366-1370)."""
    q = np.asarray(q, np.float32)
    return np.floor((q * np.float32(0.5) + np.float32(0.5)) * np.float32(0xffff) - np.float32(0.5)).astype(np.uint16)


def instanced_scene(n_base_tris=100_000, n_instances=100, seed=0xC4C4C4, edge=None):
    """BASELINE configs[3]: the first n_base_tris triangles of the C2 generator scaled x0.1 (edge 0.015 in a +-1 box),
    instanced on a 5 x 5 x k lattice of pitch 6.  Reduced test sizes keep the surface density by growing the edges."""
    s = Scene()
    scale, base = 2.0 ** -19, -2.0
    if edge is None:
        edge = min(0.3, 0.015 * (100_000 / n_base_tris) ** 0.5)
    g = random_triangle_grid(n_base_tris, seed=0x5EED1A7B200, box=1.0, edge=edge, scale=scale, base=base)
    geo = Geometry(pack_qverts(g.reshape(-1, 3)), (scale,) * 3, (base + 2.0 ** -20,) * 3)
    mesh = s.add_mesh([geo])
    na = T.BASE_MATERIAL_NOALPHA
    mats = []
    for j in range(8):  # GGX opaque
        mats.append(T.BaseMaterial(base_color=_PALETTE[j], roughness=0.15 + 0.1 * j, metallic=float(j & 1), ior=1.5, flags=na))
    for j in range(4):  # transmissive: two thick (ONESIDED), two thin
        fl = na | T.BASE_MATERIAL_EXTENDED | (T.BASE_MATERIAL_ONESIDED if j < 2 else 0)
        mats.append(T.BaseMaterial(base_color=_PALETTE[8 + j], roughness=0.05 + 0.1 * j, ior=1.33 + 0.1 * j, specular_transmission=1.0,
                                   clearcoat_gloss=0.02 + 0.05 * j, flags=fl))
    for j in range(2):  # alpha-tested: 1 x 1 sRGB base-colour texture with alpha 128/255 -> stochastic candidates (SURVEY 8d, C4 input)
        rgb8 = [int(round(255.0 * c ** (1.0 / 2.2))) for c in _PALETTE[12 + j]]
        tex = s.add_texture(rgb8 + [128], T.COLOR_SPACE_SRGB)
        mats.append(T.BaseMaterial(base_color=(T.texture_handle(tex), 0.0, 0.0), roughness=0.6, ior=1.5, flags=0))
    for j in range(2):  # emissive
        mats.append(T.BaseMaterial(base_color=(1.0, 0.8 - 0.3 * j, 0.5 + 0.4 * j), emission_intensity=20.0, flags=na))
    s.materials = mats
    ids = (np.arange(n_base_tris) % 14).astype(np.uint8)
    n_em = min(64, n_base_tris)
    ids[:n_em] = 14 + (np.arange(n_em) % 2)
    pm = s.add_pmesh(mesh, [0], tri_material_ids=ids)
    u = splitmix64_uniform(seed, n_instances * 8).reshape(n_instances, 8)
    for i in range(n_instances):
        cell = np.array([i % 5, (i // 5) % 5, i // 25], np.float32) * np.float32(6.0)
        scl = np.float32(0.5 + u[i, 0])
        q = (u[i, 1:5] * 2.0 - 1.0).astype(np.float32)
        q = q / np.float32(np.sqrt(np.float32(np.dot(q, q))))
        s.add_instance(pm, vks_instance_transform(cell, scl, quantize_quaternion(q)))
    pos = np.array([t[:, 3] for _, t in s.instances], np.float32)  # instance origins after the (x, y, z) -> (-x, z, y) flip
    centre = 0.5 * (pos.min(0) + pos.max(0))
    radius = float(np.abs(pos - centre).max()) + 3.0
    s.camera = look_at_camera(centre + np.array([0.0, 0.2 * radius, 2.4 * radius], np.float32), centre, fovy=50.0)
    s.name = "instanced%dx%d" % (n_base_tris, n_instances)
    return s


# ---------------------------------------------------------------------------------------------------------------------
# Smooth-shaded, uv-mapped geometry (rendering/rt/hit.glsl:58-128): vertex normals, uvs, uv-derivative tangents
# ---------------------------------------------------------------------------------------------------------------------
def pack_qnormal_uv(normals, uvs):
    """(..., 3) normals + (..., 2) uvs -> uint64 words: oct-encoded normal in the low half, quantised uv in the high half
    (librender/quantize.h:21-42; librender/dequantize.glsl:23-48)."""
    return quantize_normal(normals).astype(np.uint64) | (quantize_uv(uvs).astype(np.uint64) << np.uint64(32))


def smooth_shaded_scene(n_u=24, n_v=12, n_soup=1500, seed=11):
    """Tessellated unit spheres with smooth vertex normals and a spherical uv map (geometry 0: normals + uvs), a soup of
    random triangles with arbitrary vertex normals -- many of them on the far side of the geometric normal, so the flip of
    hit.glsl:71-72 and the incident-direction fix of pt_megakernel.glsl:657-668 are exercised -- and per-triangle constant
    uvs (zero uv derivatives: the tangent falls back to cross(e2, gn), hit.glsl:117-121) (geometry 1: normals + uvs), and a
    soup with uvs only (geometry 2) and normals only (geometry 3).  Instanced three times: identity, a rotation with
    non-uniform scale (normals go through the inverse transpose) and a mirrored instance (negative determinant).
    Half of the materials carry a one-texel normal map, whose tangent frame comes from the uv derivatives."""
    s = Scene()
    scale, base = 2.0 ** -19, -2.0
    offset = (base + 2.0 ** -20,) * 3

    def snap(p):
        g = np.clip(np.floor((np.asarray(p, np.float64) - base) / scale), 0, 0x1FFFFF).astype(np.int64)
        return g, (g.astype(np.float32) * np.float32(scale) + np.float32(offset[0]))

    # lat-long sphere, unrolled vertices
    iu, iv = np.meshgrid(np.arange(n_u), np.arange(n_v), indexing="ij")
    quads = []
    for du, dv in ((0, 0), (1, 0), (1, 1), (0, 0), (1, 1), (0, 1)):
        quads.append(np.stack([(iu + du).ravel(), (iv + dv).ravel()], -1))
    uvi = np.stack(quads, 1).reshape(-1, 2).astype(np.float64)  # (n_tris * 3, 2) lattice coordinates
    phi, theta = uvi[:, 0] / n_u * 2 * np.pi, uvi[:, 1] / n_v * np.pi
    pos = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1)
    g0, p0 = snap(pos)
    nrm0 = p0 / np.maximum(np.linalg.norm(p0, axis=1, keepdims=True), 1e-9)
    uv0 = np.stack([uvi[:, 0] / n_u * 3.0, uvi[:, 1] / n_v], -1)  # u repeats three times around the sphere
    geo0 = Geometry(pack_qverts(g0), (scale,) * 3, offset, qnormal_uv=pack_qnormal_uv(nrm0, uv0), has_normals=True, has_uvs=True)

    u = splitmix64_uniform(seed, n_soup * 3 * 14).reshape(n_soup, 3, 14)
    c = (u[:, :1, 0:3] * 2.0 - 1.0) * 1.2
    tri = c + (u[:, :, 3:6] * 2.0 - 1.0) * 0.25
    g1, p1 = snap(tri.reshape(-1, 3))
    rn = u[:, :, 6:9].reshape(-1, 3) * 2.0 - 1.0
    rn = rn / np.maximum(np.linalg.norm(rn, axis=1, keepdims=True), 1e-9)
    const_uv = np.repeat(u[:, :1, 9:11], 3, 1).reshape(-1, 2)         # one uv per triangle: zero derivatives
    rand_uv = (u[:, :, 11:13].reshape(-1, 2) * 2.0)                   # arbitrary uvs
    third = n_soup // 3
    sl = [slice(0, 3 * third), slice(3 * third, 6 * third), slice(6 * third, 3 * n_soup)]
    geo1 = Geometry(pack_qverts(g1[sl[0]]), (scale,) * 3, offset, qnormal_uv=pack_qnormal_uv(rn[sl[0]], const_uv[sl[0]]), has_normals=True, has_uvs=True)
    geo2 = Geometry(pack_qverts(g1[sl[1]]), (scale,) * 3, offset, qnormal_uv=pack_qnormal_uv(rn[sl[1]], rand_uv[sl[1]]), has_normals=False, has_uvs=True)
    geo3 = Geometry(pack_qverts(g1[sl[2]]), (scale,) * 3, offset, qnormal_uv=pack_qnormal_uv(rn[sl[2]], rand_uv[sl[2]]), has_normals=True, has_uvs=False)
    mesh = s.add_mesh([geo0, geo1, geo2, geo3])
    na = T.BASE_MATERIAL_NOALPHA
    t_n0 = s.add_texture((170, 96, 255), T.COLOR_SPACE_LINEAR)
    t_n1 = s.add_texture((90, 150, 255), T.COLOR_SPACE_LINEAR)
    s.materials = [T.BaseMaterial(base_color=_PALETTE[1], roughness=0.35, ior=1.5, flags=na, normal_map=t_n0),
                   T.BaseMaterial(base_color=_PALETTE[2], roughness=0.2, metallic=1.0, ior=1.5, flags=na),
                   T.BaseMaterial(base_color=_PALETTE[3], roughness=0.6, ior=1.5, flags=na, normal_map=t_n1),
                   T.BaseMaterial(base_color=_PALETTE[4], roughness=1.0, ior=1.0, flags=na)]
    pm = s.add_pmesh(mesh, [0, 1, 2, 3])
    s.add_instance(pm)
    rot = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]], np.float32)  # orthonormal
    t1 = np.zeros((3, 4), np.float32)
    t1[:, :3] = rot * np.array([1.6, 0.7, 1.1], np.float32)  # columns scaled: non-uniform scale then rotation
    t1[:, 3] = (3.2, 0.4, -0.5)
    s.add_instance(pm, t1)
    t2 = np.zeros((3, 4), np.float32)
    t2[:, :3] = np.diag([-1.2, 1.0, 0.9]).astype(np.float32)  # mirrored
    t2[:, 3] = (-3.1, -0.3, 0.4)
    s.add_instance(pm, t2)
    s.camera = look_at_camera((0.3, 1.5, 7.5), (0.0, 0.0, 0.0), fovy=50.0)
    s.name = "smooth_shaded"
    return s


# ---------------------------------------------------------------------------------------------------------------------
# Textures larger than 1 x 1 (rendering/rt/material_textures.glsl:26-145): uv lookups at the hit and at alpha candidates
# ---------------------------------------------------------------------------------------------------------------------
def procedural_image(w, h, channels, seed):
    """A deterministic 8-bit test image: smooth gradients + a checker + hashed noise, so that bilinear weights, wrap-around and
    every channel matter."""
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    img = np.zeros((h, w, channels), np.uint8)
    for c in range(channels):
        v = (x * (37 + 11 * c) + y * (59 + 7 * c) + seed * 101) % 256
        checker = (((x * 4 // max(w, 1)) + (y * 4 // max(h, 1)) + c) % 2) * 96
        img[..., c] = ((v // 2) + checker + ((x * y + seed) % 23)) % 256
    return img


def textured_scene(seed=5):
    """The smooth-shaded, uv-mapped scene with textures of several sizes on every textured parameter: an sRGB RGBA base colour whose
    alpha channel cuts holes (stochastic alpha at candidates with interpolated uv: closest-hit and shadow rays), a linear
    RGB(A) image read channel by channel for specular / roughness / metallic, a tangent-space normal map, a non-square image, a
    1 x 1 texture beside them (folded on the host), and a NOALPHA material that reads colour but is never alpha-tested."""
    s = smooth_shaded_scene()
    t_color = s.add_texture(procedural_image(64, 32, 4, seed), T.COLOR_SPACE_SRGB)
    alpha = t_color_img = s.textures[t_color][0]
    yy, xx = np.mgrid[0:32, 0:64]
    alpha[..., 3] = np.where(((xx // 8 + yy // 8) % 3) == 0, 0, np.where(((xx // 8 + yy // 8) % 3) == 1, 140, 255)).astype(np.uint8)
    t_orm = s.add_texture(procedural_image(16, 16, 3, seed + 1), T.COLOR_SPACE_LINEAR)
    nrm = procedural_image(32, 32, 3, seed + 2).astype(np.int32)
    nrm = np.stack([128 + (nrm[..., 0] - 128) // 3, 128 + (nrm[..., 1] - 128) // 3, np.full(nrm.shape[:2], 255)], -1).astype(np.uint8)
    t_nrm = s.add_texture(nrm, T.COLOR_SPACE_LINEAR)
    t_gray = s.add_texture(procedural_image(5, 3, 1, seed + 3), T.COLOR_SPACE_SRGB)  # one channel: (r, 0, 0, 1)
    t_one = s.add_texture((90, 200, 60, 255), T.COLOR_SPACE_SRGB)
    m = s.materials
    m[0].base_color, m[0].flags = (T.texture_handle(t_color), 0.0, 0.0), 0           # alpha-tested, textured colour, one-texel normal map stays
    m[1].roughness, m[1].metallic, m[1].specular = T.texture_handle(t_orm, 1), T.texture_handle(t_orm, 2), T.texture_handle(t_orm, 0)
    m[2].base_color, m[2].normal_map = (T.texture_handle(t_color), 0.0, 0.0), t_nrm  # NOALPHA stays set: colour from the image, no alpha test
    m[3].base_color, m[3].ior = (T.texture_handle(t_gray), 0.0, 0.0), 1.5
    m[3].roughness = T.texture_handle(t_orm, 0)
    s.materials.append(T.BaseMaterial(base_color=(T.texture_handle(t_one), 0.0, 0.0), roughness=0.5, flags=0))
    s.name = "textured"
    return s
