/* oracle/ref_shim/ref_vkr.c -- TEST INFRASTRUCTURE.
 * Flat accessors over the reference's own .vks / .vkt reader (ext/libvkr/src/vkr.c, compiled where it lies into oracle/_ref): what
 * vkr_open_scene / vkr_open_texture make of a file, and vkr_quantize_transform / vkr_dequantize_transform on one record.
 * tests/test_vks.py holds realtimepathtracingresearchframework_b200/vks.py to these in both directions (files written by vks.py are
 * read by vkr.c; both readers must report the same tables). */
#include <stdint.h>
#include <string.h>
#include <vkr.h>

/* counts: version, numMeshes, numInstances, numMaterials, numTriangles, numLodGroups, numFrames, numStaticTransforms,
 * numAnimatedTransforms, animationOffset, headerSize, dataOffset; returns the VkrResult */
int ref_vkr_scene_counts(const char *path, int64_t *out) {
    VkrScene s;
    const int r = vkr_open_scene(path, &s, 0);
    if (r != VKR_SUCCESS) return r;
    out[0] = s.version; out[1] = (int64_t)s.numMeshes; out[2] = (int64_t)s.numInstances; out[3] = (int64_t)s.numMaterials;
    out[4] = (int64_t)s.numTriangles; out[5] = (int64_t)s.numLodGroups; out[6] = (int64_t)s.numFrames; out[7] = (int64_t)s.numStaticTransforms;
    out[8] = (int64_t)s.numAnimatedTransforms; out[9] = s.animationOffset; out[10] = s.headerSize; out[11] = s.dataOffset;
    vkr_close_scene(&s);
    return 0;
}
/* mesh i: ints = numSegments, numTriangles, materialIdBufferBase, numMaterialsInRange, lodGroup, vertexBufferOffset, normalUvBufferOffset,
 * materialIdBufferOffset, materialIdSize, flags; floats = scale3, offset3; segs = per segment (numTriangles, materialBaseOffset), up to
 * max_segs; name copied into name[128] */
int ref_vkr_mesh(const char *path, int64_t i, int64_t *ints, float *floats, int64_t *segs, int64_t max_segs, char *name) {
    VkrScene s;
    const int r = vkr_open_scene(path, &s, 0);
    if (r != VKR_SUCCESS) return r;
    if (i < 0 || (uint64_t)i >= s.numMeshes) { vkr_close_scene(&s); return -100; }
    const VkrMesh *m = s.meshes + i;
    ints[0] = (int64_t)m->numSegments; ints[1] = (int64_t)m->numTriangles; ints[2] = m->materialIdBufferBase; ints[3] = m->numMaterialsInRange;
    ints[4] = m->lodGroup; ints[5] = m->vertexBufferOffset; ints[6] = m->normalUvBufferOffset; ints[7] = m->materialIdBufferOffset;
    ints[8] = (int64_t)m->materialIdSize; ints[9] = m->flags;
    for (int k = 0; k < 3; ++k) { floats[k] = m->vertexScale[k]; floats[3 + k] = m->vertexOffset[k]; }
    for (uint64_t j = 0; j < m->numSegments && (int64_t)j < max_segs; ++j) { segs[2 * j] = (int64_t)m->segmentNumTriangles[j]; segs[2 * j + 1] = m->segmentMaterialBaseOffsets[j]; }
    strncpy(name, m->name, 127); name[127] = 0;
    vkr_close_scene(&s);
    return 0;
}
/* instance i: meshId, transformIndex, flags; material i: its name and the parameters vkr_load_material found */
int ref_vkr_instance(const char *path, int64_t i, int64_t *out) {
    VkrScene s;
    const int r = vkr_open_scene(path, &s, 0);
    if (r != VKR_SUCCESS) return r;
    if (i < 0 || (uint64_t)i >= s.numInstances) { vkr_close_scene(&s); return -100; }
    out[0] = s.instances[i].meshId; out[1] = s.instances[i].transformIndex; out[2] = s.instances[i].flags;
    vkr_close_scene(&s);
    return 0;
}
/* floats: emissionIntensity, emitterBaseColor3, specularTransmission, iorEta, iorK, translucency; tex: per texture (base colour, normal,
 * specular) present?, width, height, format, numMipLevels, dataSize, dataOffset */
int ref_vkr_material(const char *path, int64_t i, char *name, float *floats, int64_t *tex) {
    VkrScene s;
    const int r = vkr_open_scene(path, &s, 0);
    if (r != VKR_SUCCESS) return r;
    if (i < 0 || (uint64_t)i >= s.numMaterials) { vkr_close_scene(&s); return -100; }
    const VkrMaterial *m = s.materials + i;
    strncpy(name, m->name, 127); name[127] = 0;
    floats[0] = m->emissionIntensity; floats[1] = m->emitterBaseColor[0]; floats[2] = m->emitterBaseColor[1]; floats[3] = m->emitterBaseColor[2];
    floats[4] = m->specularTransmission; floats[5] = m->iorEta; floats[6] = m->iorK; floats[7] = m->translucency;
    const VkrTexture *t[3] = {&m->texBaseColor, &m->texNormal, &m->texSpecularRoughnessMetalness};
    for (int k = 0; k < 3; ++k) {
        tex[7 * k] = t[k]->filename != 0;
        tex[7 * k + 1] = t[k]->width; tex[7 * k + 2] = t[k]->height; tex[7 * k + 3] = t[k]->format; tex[7 * k + 4] = t[k]->numMipLevels;
        tex[7 * k + 5] = (int64_t)t[k]->dataSize; tex[7 * k + 6] = t[k]->dataOffset;
    }
    vkr_close_scene(&s);
    return 0;
}
int64_t ref_vkr_transform_offset(uint32_t index, uint64_t n_static, uint64_t n_animated, uint64_t frame) {
    return (int64_t)vkr_get_transform_offset(index, n_static, n_animated, frame);
}
void ref_vkr_dequantize_transform(const unsigned char *quantized, float *matrix12) { vkr_dequantize_transform((float(*)[3])matrix12, quantized); }
void ref_vkr_quantize_transform(const float *matrix12, unsigned char *quantized) { vkr_quantize_transform(quantized, (const float(*)[3])matrix12); }
