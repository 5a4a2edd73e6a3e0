"""Synthetic frames for the temporal passes (shared by tests/test_temporal_passes.py and oracle/gen_golden.py --post-only)."""
import numpy as np


def synthetic_frame(rng, w, h, motion_scale=0.02, depth=5.0):
    """accumulator sample, history (alpha = 1 - weight), normal / depth images and a smooth + noisy motion field"""
    cur = rng.uniform(0.0, 2.0, (h, w, 4)).astype(np.float32)
    cur[..., 3] = (rng.uniform(size=(h, w)) > 0.2).astype(np.float32)         # alpha of a sample: the primary ray hit something
    cur[rng.uniform(size=(h, w)) > 0.97, 3] = 2.0                              # "non-accumulation object types" (alpha > 1)
    hist = rng.uniform(0.0, 2.0, (h, w, 4)).astype(np.float32)
    hist[..., 3] = rng.choice(np.array([0.0, 0.5, 0.75, 0.875, 1.0], np.float32), size=(h, w))
    def nd_image():
        n = rng.normal(size=(h, w, 3))
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        n = 0.3 * n + np.array([0.0, 0.0, 1.0])                                 # mostly facing the camera, some spread
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        d = depth * (1.0 + 0.02 * rng.normal(size=(h, w, 1)))
        d[rng.uniform(size=(h, w, 1)) > 0.9] *= 3.0                             # depth edges
        return np.concatenate([n, d], -1).astype(np.float16)
    yy, xx = np.mgrid[0:h, 0:w]
    mj = np.zeros((h, w, 4), np.float32)
    mj[..., 0] = motion_scale * np.sin(xx / 7.0) + 0.3 * motion_scale * rng.normal(size=(h, w))
    mj[..., 1] = motion_scale * np.cos(yy / 5.0) + 0.3 * motion_scale * rng.normal(size=(h, w))
    mj[rng.uniform(size=(h, w)) > 0.98, :2] = 3.0                               # reprojects outside the frame
    return cur, hist, nd_image(), nd_image(), mj.astype(np.float16)
