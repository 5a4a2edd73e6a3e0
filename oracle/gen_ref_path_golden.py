#!/usr/bin/env python
"""Writes tests/golden/ref_path_images.npz: small frames rendered by oracle/_ref's WHOLE-PATH driver (ref_shim/ref_path.cpp: the
megakernel's path sample composed from the reference-executed pieces, intersections from the oracle).  Needs oracle/_ref
(i.e. /root/reference at build time).  tests/test_ref_path.py holds the oracle -- and through it the CUDA path -- to these."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
from realtimepathtracingresearchframework_b200 import load_sky_fit, types as T  # noqa: E402
import ref_path_util as U  # noqa: E402

if __name__ == "__main__":
    out = {}
    for name, (make, sky, (w, h), spp) in U.CASES.items():
        s = make()
        o = po.OracleScene(s)
        img = U.ref_path_render(o, s, w, h, s.camera, load_sky_fit(T.SceneConfig(**sky)), spp)
        out[name] = img
        print(name, img.shape, float(img[..., :3].mean()))
    path = os.path.join(ROOT, "tests", "golden", "ref_path_images.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)
