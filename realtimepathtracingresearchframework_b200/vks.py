"""Reading (and, for tests and exports, writing) the reference's scene files: `.vks` scenes with their `<scene>_textures/` directory of
`.vkt` textures and parameter text files.

Behind the backend boundary the reference's own loader does this (librender/scene.cpp:544-1006 `Scene::load_vkrs` on top of
ext/libvkr/src/vkr.c) and hands the backend a `Scene`; this module is the same step for the standalone host path (scenes.Scene ->
rptr_scene_desc).  The container layout follows vkr.c (`vkr_load_scene` :771-1144, `vkr_open_texture` :211-306,
`vkr_load_material` :521-625, `vkr_get_transform_offset` :197-209): file version 3 and 4 are read, version 4 is written.  The mapping
to a scene follows scene.cpp: one Geometry per non-empty mesh segment over the mesh's quantised vertex / normal+uv streams, one
parameterized mesh per mesh (per-triangle 8-bit material ids for single-segment meshes with several materials, per-segment material
offsets otherwise), base-level instances only (meshes of higher LoD levels are skipped), transforms from the quantised table at
frame 0 with the vks axis flip, three textures per material (base colour sRGB, normal BC5, specular / roughness / metallic) with
1 x 1 defaults, roughness / metallic read from channels 1 / 2 of the third, emission and transmission from the parameter files.
tests/test_vks.py pins the container to vkr.c itself (compiled into oracle/_ref) in both directions.
"""
import os
import struct

import numpy as np

from . import scenes
from . import types as T

VKR_MAGIC, VKT_MAGIC = 0xABCABC, 0xBC1BC1
QUANTIZED_TRANSFORM_SIZE = 24  # 3 x f32 translation, f32 scaling, 4 x u16 quaternion (vkr.h:15, vkr.c:1346-1408)
# VkFormat codes of .vkt files (vkr.h:51-68) -> (Image::bcFormat, has an alpha channel), librender/scene.cpp:836-860
VK_FORMATS = {131: (1, False), 132: (1, False), 133: (-1, True), 134: (-1, True), 137: (3, True), 138: (3, True), 141: (5, False), 37: (0, True),
              43: (0, True)}
BLOCK_BYTES = {1: 8, -1: 8, 3: 16, 5: 16}


class VksError(Exception):
    pass


def texture_dir(scene_file):
    """buildTextureDir (vkr.c:80-107): the scene file's name without its extension + "_textures/" """
    base = scene_file[:scene_file.rfind(".")] if "." in os.path.basename(scene_file) else scene_file
    return base + "_textures/"


# ---------------------------------------------------------------------------------------------------------------------------------
# reading
# ---------------------------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, data, name):
        self.d, self.pos, self.name = data, 0, name

    def take(self, fmt):
        size = struct.calcsize("<" + fmt)
        if self.pos + size > len(self.d):
            raise VksError("%s: truncated file" % self.name)
        v = struct.unpack_from("<" + fmt, self.d, self.pos)
        self.pos += size
        return v if len(v) > 1 else v[0]

    def string(self):
        n = self.take("Q")
        if self.pos + n + 1 > len(self.d):
            raise VksError("%s: truncated string" % self.name)
        s = bytes(self.d[self.pos:self.pos + n]).decode("utf-8", "replace")
        self.pos += n + 1  # the terminating zero is stored (vkr_load_string, vkr.c:317-347)
        return s


def read_vkt(path):
    """A .vkt texture: dict(width, height, format (VkFormat), levels=[(w, h, bytes)]) or None when the file does not exist
    (textures are optional, vkr.c:222-233)."""
    if not os.path.exists(path):
        return None
    data = np.fromfile(path, np.uint8)
    r = _Reader(memoryview(data), path)
    if r.take("i") != VKT_MAGIC:
        raise VksError("%s is not a .vkt file" % path)
    version = r.take("i")
    if version != 1:
        raise VksError("%s: unsupported texture version %d" % (path, version))
    n_mips, width, height, fmt = r.take("iiii")
    data_size = r.take("Q")
    mips = [r.take("iiQq") for _ in range(n_mips)]
    data_offset = r.pos
    if data_offset + data_size > len(data):
        raise VksError("%s: truncated texel data" % path)
    levels = [(w, h, data[off:off + size]) for (w, h, size, off) in mips]
    return dict(width=width, height=height, format=fmt, levels=levels, blob=data[data_offset:data_offset + data_size])


def _param_file(path, max_values):
    """vkr_parse_material_param_file (vkr.c:411-453): up to max_values floats, one per line; [] when the file does not exist"""
    if not os.path.exists(path):
        return []
    vals = []
    for line in open(path).read().split("\n"):
        if len(vals) == max_values or not line.strip():
            break
        vals.append(float(line.strip()))
    return vals


def read_vks_container(path):
    """The tables of a .vks file as vkr_open_scene fills them (no scene mapping yet)."""
    data = np.fromfile(path, np.uint8)
    r = _Reader(memoryview(data), path)
    if r.take("i") != VKR_MAGIC:
        raise VksError("%s is not a .vks file" % path)
    version = r.take("i")
    if version not in (3, 4):
        raise VksError("%s: file version %d is not supported (3 and 4 are)" % (path, version))
    flags, header_size, data_offset = r.take("QQQ")
    n_meshes, n_instances, n_materials, n_triangles, n_groups = r.take("QQQQQ")
    out = dict(version=version, flags=flags, n_triangles=n_triangles)
    n_lod_groups, lod_offset = 1, 0
    anim = dict(n_frames=1, n_static=n_instances, n_animated=0, offset=0, start=0.0, step=0.0)
    if version >= 4:
        n_lod_groups, lod_offset, n_bone, bone_offset = r.take("QqQq")
        start, step = r.take("ff")
        n_frames, n_static, n_animated, anim_offset = r.take("QQQq")
        anim = dict(n_frames=n_frames, n_static=n_static, n_animated=n_animated, offset=anim_offset, start=start, step=step)
    if not (n_meshes and n_instances and n_groups and n_lod_groups):
        raise VksError("%s: invalid object counts" % path)
    if header_size != r.pos:
        raise VksError("%s: mismatching header size" % path)
    meshes = []
    for i in range(n_meshes):
        scale, offset = r.take("fff"), r.take("fff")
        mflags, header_end, vb_offset = r.take("QQQ")
        n_segments, n_tris, mat_base, n_mats_in_range = r.take("QQII")
        lod_group = 0
        if version >= 4:
            lod_group = r.take("q")
            r.take("4Q")
        else:
            r.take("5Q")
        seg_tris = [r.take("Q") for _ in range(n_segments)]
        seg_mats = [r.take("i") for _ in range(n_segments)]
        name = r.string()
        if header_end != r.pos:
            raise VksError("%s: mismatching header offset for mesh %d" % (path, i))
        if lod_group >= n_lod_groups:
            raise VksError("%s: invalid LoD group for mesh %d" % (path, i))
        meshes.append(dict(name=name, scale=scale, offset=offset, flags=mflags, vertex_buffer_offset=vb_offset, n_tris=n_tris,
                           material_id_base=mat_base, n_materials_in_range=n_mats_in_range, lod_group=lod_group, segment_tris=seg_tris,
                           segment_material_base=seg_mats))
    instances = []
    transforms_inline = []
    for g in range(n_groups):
        iflags, mesh_id = r.take("Ii")
        header_end, group_data, n_in_group = r.take("QQQ")
        name = r.string()
        if group_data != r.pos:
            raise VksError("%s: mismatching data offset for instance group %d" % (path, g))
        for _ in range(n_in_group):
            if version >= 4:
                instances.append(dict(name=name, mesh_id=mesh_id, flags=iflags, transform_index=r.take("I")))
            else:  # version 3 stores float[4][3] per instance; vkr.c quantises them into the table (vkr_quantize_transform)
                transforms_inline.append(np.array(r.take("12f"), np.float32).reshape(4, 3))
                instances.append(dict(name=name, mesh_id=mesh_id, flags=iflags, transform_index=len(transforms_inline) - 1))
        if header_end != r.pos:
            raise VksError("%s: mismatching header offset for instance group %d" % (path, g))
    if len(instances) != n_instances:
        raise VksError("%s: instance count does not match the groups" % path)
    lod_groups = [dict(mesh_ids=[], detail=[])]
    if version >= 4:
        if lod_offset != r.pos:
            raise VksError("%s: invalid LoD group offset" % path)
        lod_groups = []
        for _ in range(n_lod_groups):
            n = r.take("Q")
            ids = [r.take("q") for _ in range(n)]
            lod_groups.append(dict(mesh_ids=ids, detail=[r.take("f") for _ in range(n)]))
    if data_offset != r.pos:
        raise VksError("%s: mismatching body data offset" % path)
    materials = [r.string() for _ in range(n_materials)]
    offset = r.pos
    for m in meshes:  # vkr.c:1112-1141
        if m["vertex_buffer_offset"] != offset:
            raise VksError("%s: mismatching data offset for mesh %s" % (path, m["name"]))
        n = m["n_tris"]
        m["qverts"] = data[offset:offset + 24 * n].view(np.uint64)
        offset += 24 * n
        m["qnuv"] = data[offset:offset + 24 * n].view(np.uint64)
        offset += 24 * n
        m["material_id_size"] = 1 if (m["n_materials_in_range"] <= 256 or len(m["segment_tris"]) > 1) else 2
        m["material_ids"] = data[offset:offset + m["material_id_size"] * n]
        offset += m["material_id_size"] * n
        if m["flags"] & 1:  # VKR_MESH_FLAGS_INDICES: vertex-sharing indices, not needed for unrolled triangles
            offset += 12 * n
        if offset > len(data):
            raise VksError("%s: truncated mesh data" % path)
    n_transforms = anim["n_static"] + anim["n_animated"] * anim["n_frames"]
    if version >= 4:
        a0 = anim["offset"]
        if a0 + QUANTIZED_TRANSFORM_SIZE * n_transforms > len(data):
            raise VksError("%s: truncated transform table" % path)
        table = data[a0:a0 + QUANTIZED_TRANSFORM_SIZE * n_transforms]
    else:
        table = None
    out.update(meshes=meshes, instances=instances, lod_groups=lod_groups, materials=materials, animation=anim, transform_table=table,
               transforms_inline=transforms_inline)
    return out


def transform_offset(index, n_static, n_animated, frame):
    """vkr_get_transform_offset (vkr.c:197-209)"""
    if index < n_static:
        return index
    return n_static + n_animated * frame + (index - n_static)


def load_vks(path, ignore_textures=False, load_specularity=False):
    """Scene::load_vkrs (librender/scene.cpp:544-1006) for one file into a scenes.Scene (no camera: .vks files have none)."""
    c = read_vks_container(path)
    s = scenes.Scene()
    s.name = os.path.basename(path)
    # meshes -> geometries + parameterized meshes (:596-710)
    for m in c["meshes"]:
        geoms, base = [], 0
        for n in m["segment_tris"]:
            if n > 0:
                geoms.append(scenes.Geometry(np.ascontiguousarray(m["qverts"][3 * base:3 * (base + n)]), m["scale"], m["offset"],
                                             qnormal_uv=np.ascontiguousarray(m["qnuv"][3 * base:3 * (base + n)]), has_normals=True, has_uvs=True))
            base += n
        mesh_id = s.add_mesh(geoms)
        if len(m["segment_tris"]) == 1 and m["n_materials_in_range"] > 1:
            if m["material_id_size"] != 1:
                raise VksError("%s: 16-bit material ids (deprecated in the format) are not supported" % path)
            s.add_pmesh(mesh_id, [m["material_id_base"]], np.ascontiguousarray(m["material_ids"]))
        else:
            offsets = [b for b, n in zip(m["segment_material_base"], m["segment_tris"]) if n > 0]
            s.add_pmesh(mesh_id, offsets)
    # instances: base LoD level only (:733-755), transform of frame 0 (AnimationData::dequantize, :22-41)
    an = c["animation"]
    for inst in c["instances"]:
        mesh = c["meshes"][inst["mesh_id"]]
        group = c["lod_groups"][mesh["lod_group"]]
        if group["mesh_ids"] and group["mesh_ids"][0] != inst["mesh_id"]:
            continue
        if c["transform_table"] is not None:
            o = QUANTIZED_TRANSFORM_SIZE * transform_offset(inst["transform_index"], an["n_static"], an["n_animated"], 0)
            rec = c["transform_table"][o:o + QUANTIZED_TRANSFORM_SIZE]
            tr = rec[:12].view(np.float32)
            m43 = scenes.vks_instance_transform(tr, rec[12:16].view(np.float32)[0], rec[16:24].view(np.uint16), flip=True)
        else:
            m43 = scenes.vks_flip(c["transforms_inline"][inst["transform_index"]])
        s.add_instance(inst["mesh_id"], m43)
    # materials (:818-1003): three textures each
    tdir = texture_dir(path)
    s.materials = []
    for i, name in enumerate(c["materials"]):
        m = T.BaseMaterial()
        ext_path = os.path.join(tdir, name + "_Ex.txt")
        extended = open(ext_path).read() if os.path.exists(ext_path) else name

        def image(kind, default_texel, color_space, forced_bc=None):
            t = None if ignore_textures else read_vkt(os.path.join(tdir, "%s_%s.vkt" % (name, kind)))
            if t is None:
                return s.add_texture(default_texel, color_space), False
            bc, has_alpha = VK_FORMATS.get(t["format"], (0, False))
            if forced_bc is not None:
                bc = forced_bc
            s.textures.append(("vkt", color_space, t, bc))
            return len(s.textures) - 1, has_alpha
        t_color, has_alpha = image("BaseColor", (255, 255, 255, 255), T.COLOR_SPACE_SRGB)
        if not has_alpha:
            m.flags |= T.BASE_MATERIAL_NOALPHA
        m.base_color = (T.texture_handle(t_color), m.base_color[1], m.base_color[2])
        t_normal, _ = image("Normal", (127, 127, 127, 255), T.COLOR_SPACE_LINEAR, forced_bc=5)
        m.normal_map = t_normal
        t_spec, _ = image("Specular", (255, 127, 0, 255), T.COLOR_SPACE_LINEAR, forced_bc=1)
        m.roughness = T.texture_handle(t_spec, 1)
        m.metallic = T.texture_handle(t_spec, 2)
        if load_specularity:
            m.specular = T.texture_handle(t_spec, 0)
        # parameter files (vkr_load_material, vkr.c:521-592)
        em = _param_file(os.path.join(tdir, name + "_EmissionIntensity.txt"), 4)
        intensity, color = 0.0, (0.0, 0.0, 0.0)
        if len(em) == 1:
            intensity = em[0]
            bcol = _param_file(os.path.join(tdir, name + "_BaseColor.txt"), 3)
            if len(bcol) not in (0, 3):
                raise VksError("three colour components expected for the emission base colour of " + name)
            if bcol:
                color = tuple(bcol)
        elif len(em) == 4:
            intensity, color = em[0], tuple(em[1:])
        elif em:
            raise VksError("one or four components expected for the emission of " + name)
        if intensity > 0:
            if color != (0.0, 0.0, 0.0):
                m.base_color = color
            m.emission_intensity = intensity
        tr = _param_file(os.path.join(tdir, name + "_SpecularTransmission.txt"), 4)
        vals = [0.0, 1.5, 0.0, 0.0]
        vals[:len(tr)] = tr
        m.specular_transmission = vals[0]
        if m.specular_transmission and not any(k in extended for k in ("twosided", "doublesided", "TwoSided", "DoubleSided")):
            m.flags |= T.BASE_MATERIAL_ONESIDED
        m.ior = vals[1]
        s.materials.append(m)
    return s


# ---------------------------------------------------------------------------------------------------------------------------------
# writing (file version 4)
# ---------------------------------------------------------------------------------------------------------------------------------
def _string(sv):
    b = sv.encode("utf-8")
    return struct.pack("<Q", len(b)) + b + b"\0"


def quantize_transform(translation, scaling, quaternion):
    """One 24-byte record of the transform table from a unit quaternion (x, y, z, w) in [-1, 1]: codes = round((q + 1) * 65535 / 2)."""
    q = np.clip(np.round((np.asarray(quaternion, np.float64) + 1.0) * (0xffff / 2.0)), 0, 0xffff).astype(np.uint16)
    return np.asarray(translation, np.float32).tobytes() + np.float32(scaling).tobytes() + q.tobytes()


def write_vkt(path, levels, vk_format):
    """levels: [(h, w, c) uint8 images]; vk_format: VkFormat code (37 = RGBA8, 131.. BCn) -- block formats are encoded with scenes.encode_bc"""
    bc = VK_FORMATS[vk_format][0]
    blobs = []
    for l in levels:
        if bc:
            blobs.append(scenes.encode_bc(l, bc).tobytes())
        else:
            full = np.zeros(l.shape[:2] + (4,), np.uint8)
            full[..., 3] = 255
            full[..., :l.shape[2]] = l
            blobs.append(full.tobytes())
    header = 4 * 6 + 8 + len(levels) * 24
    out = struct.pack("<iiiiii", VKT_MAGIC, 1, len(levels), levels[0].shape[1], levels[0].shape[0], vk_format)
    out += struct.pack("<Q", sum(len(b) for b in blobs))
    off = header
    for l, b in zip(levels, blobs):
        out += struct.pack("<iiQq", l.shape[1], l.shape[0], len(b), off)
        off += len(b)
    with open(path, "wb") as f:
        f.write(out + b"".join(blobs))


def write_vks(path, meshes, instances, materials, transforms, lod_groups=None):
    """meshes: [dict(name, scale, offset, qverts (3n u64), qnuv (3n u64), segments=[(n_tris, material_base)], material_ids (n u8) | None,
    material_id_base, n_materials_in_range, lod_group)]; instances: [(name, mesh_id, transform_index)] (one group each); materials:
    [names]; transforms: [24-byte records] (all static); lod_groups: [[(mesh_id, detail_reduction)]] for groups 1.. (group 0 is empty)."""
    lod_groups = [[]] + list(lod_groups or [])
    fixed = 4 + 4 + 8 * 3 + 8 * 5 + (8 + 8 + 8 + 8) + 4 + 4 + 8 * 4
    mesh_headers, pos = [], fixed
    for m in meshes:
        seg = m["segments"]
        size = 24 + 8 * 3 + 8 * 2 + 4 * 2 + 8 + 8 * 4 + 12 * len(seg) + len(_string(m["name"]))
        pos += size
        mesh_headers.append((pos, m))
    group_blobs = []
    for name, mesh_id, tidx in instances:
        head = 4 + 4 + 8 * 3 + len(_string(name))
        pos += head
        data_off = pos
        pos += 4
        group_blobs.append((pos, data_off, name, mesh_id, tidx))
    lod_offset = pos
    lod_blob = b""
    for g in lod_groups:
        lod_blob += struct.pack("<Q", len(g)) + b"".join(struct.pack("<q", mid) for mid, _ in g) + b"".join(struct.pack("<f", d) for _, d in g)
    pos += len(lod_blob)
    data_offset = pos
    mat_blob = b"".join(_string(n) for n in materials)
    pos += len(mat_blob)
    bodies = []
    for m in meshes:
        n = sum(t for t, _ in m["segments"])
        m["_vb"] = pos
        ids = m.get("material_ids")
        id_size = 1 if (m.get("n_materials_in_range", 1) <= 256 or len(m["segments"]) > 1) else 2
        ids_b = (np.zeros(n, np.uint8) if ids is None else np.asarray(ids, np.uint8)).tobytes() if id_size == 1 else np.asarray(ids, np.uint16).tobytes()
        body = np.ascontiguousarray(m["qverts"], np.uint64).tobytes() + np.ascontiguousarray(m["qnuv"], np.uint64).tobytes() + ids_b
        assert len(body) == 48 * n + id_size * n
        bodies.append(body)
        pos += len(body)
    anim_offset = pos
    out = struct.pack("<iiQQQ", VKR_MAGIC, 4, 0, fixed, data_offset)
    out += struct.pack("<QQQQQ", len(meshes), len(instances), len(materials), sum(sum(t for t, _ in m["segments"]) for m in meshes), len(instances))
    out += struct.pack("<QqQq", len(lod_groups), lod_offset, 0, 0) + struct.pack("<ff", 0.0, 0.0)
    out += struct.pack("<QQQq", 1, len(transforms), 0, anim_offset)
    assert len(out) == fixed
    for end, m in mesh_headers:
        seg = m["segments"]
        out += struct.pack("<ffffff", *m["scale"], *m["offset"]) + struct.pack("<QQQ", 0, end, m["_vb"])
        out += struct.pack("<QQII", len(seg), sum(t for t, _ in seg), m.get("material_id_base", 0), m.get("n_materials_in_range", 1))
        out += struct.pack("<q", m.get("lod_group", 0)) + struct.pack("<4Q", 0, 0, 0, 0)
        out += b"".join(struct.pack("<Q", t) for t, _ in seg) + b"".join(struct.pack("<i", b) for _, b in seg) + _string(m["name"])
        assert len(out) == end
    for end, data_off, name, mesh_id, tidx in group_blobs:
        out += struct.pack("<IiQQQ", 0, mesh_id, end, data_off, 1) + _string(name)
        assert len(out) == data_off
        out += struct.pack("<I", tidx)
        assert len(out) == end
    out += lod_blob + mat_blob + b"".join(bodies) + b"".join(transforms)
    with open(path, "wb") as f:
        f.write(out)
