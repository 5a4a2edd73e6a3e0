"""numpy restatement of the LDR display chain (vulkan/process_samples.comp:138-200) the GPU test of readback_framebuffer(uint8*)
compares the CUDA kernel with.  Itself pinned to the reference-executed tonemap() / linear_to_srgb() by
tests/test_oracle_golden.py::test_display_chain_restatement_matches_the_reference."""
import numpy as np


def tonemap(mode, rgb):
    """rendering/postprocess/tonemapping_utils.glsl:9-32; mode < 0 = early tone mapping off"""
    x = np.asarray(rgb, np.float64)
    if mode == 1:  # NEUTRAL_TONE_MAPPING
        level = np.maximum(x.max(-1), 1.0)[..., None]
        return x * (0.1 * np.log2(level) * (1.0 - 0.8) + 1.0 * 0.8) / level
    if mode == 2:  # FAST_TONE_MAPPING
        return x / (1.0 + x)
    return x


def linear_to_srgb(x):
    """rendering/util.glsl:25-28"""
    x = np.asarray(x, np.float64)
    return np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(np.maximum(np.abs(x), 1e-30), 1 / 2.4) - 0.055)


def to_srgb8(rgb, alpha):
    """RGBA8 store of the framebuffer: clamp, scale, round half up"""
    sr = linear_to_srgb(np.maximum(np.asarray(rgb, np.float64), 0.0))
    out = np.zeros(np.shape(rgb)[:-1] + (4,), np.float64)
    out[..., :3] = np.clip(sr, 0, 1) * 255 + 0.5
    out[..., 3] = np.clip(alpha, 0, 1) * 255 + 0.5
    return np.floor(out).astype(np.int32)
