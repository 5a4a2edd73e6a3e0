"""Headless rendering of a scene file on the CUDA backend -- the Python-side stand-in for `rptr --backend cuda --validation`
(main.cpp / libapp/app_state.cpp:464-498) when the scene is a `.vks` file (librender/scene.cpp:61-62 dispatches on the extension):

    python -m realtimepathtracingresearchframework_b200.render yard.vks --img 1280 720 --validation out/yard --validation-spp 64 \
        --eye 0 2 14 --target 0 0 0 --fovy 50 [--batch-spp 8] [--sun 0.35 0.8 0.45] [--transmission]

writes out/yard_0064.pfm (RGB, bottom-up: util/write_image.cpp:34-66) and prints the frame statistics.  Procedural scenes:
`cornell`, `random:N`, `instanced:N:K`."""
import argparse
import json
import sys

from . import RenderCuda, load_sky_fit, scenes, types as T, vks, write_pfm


def load_scene(spec):
    if spec.endswith(".vks") or spec.endswith(".vkrs"):
        return vks.load_vks(spec)
    if spec == "cornell":
        return scenes.cornell_box()
    if spec.startswith("random:"):
        return scenes.random_triangles(int(spec.split(":")[1]))
    if spec.startswith("instanced:"):
        _, n, k = spec.split(":")
        return scenes.instanced_scene(int(n), int(k))
    raise SystemExit("unknown scene %r" % spec)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("scene")
    ap.add_argument("--img", type=int, nargs=2, default=(1280, 720), metavar=("W", "H"))
    ap.add_argument("--validation", default="validation", help="output prefix: <prefix>_<%%04d spp>.pfm")
    ap.add_argument("--validation-spp", type=int, default=16)
    ap.add_argument("--batch-spp", type=int, default=0, help="samples per frame (default: all in one frame)")
    ap.add_argument("--eye", type=float, nargs=3, default=None)
    ap.add_argument("--target", type=float, nargs=3, default=(0.0, 0.0, 0.0))
    ap.add_argument("--fovy", type=float, default=65.0)
    ap.add_argument("--sun", type=float, nargs=3, default=None, help="sun direction of the sky model")
    ap.add_argument("--transmission", action="store_true", help="the GLTF_SUPPORT_TRANSMISSION build")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    s = load_scene(a.scene)
    cam = scenes.look_at_camera(tuple(a.eye), tuple(a.target), fovy=a.fovy) if a.eye else getattr(s, "camera", None)
    if cam is None:
        raise SystemExit("the scene file has no camera: pass --eye X Y Z [--target X Y Z] [--fovy F]")
    r = RenderCuda(device=a.device)
    w, h = a.img
    r.initialize(w, h)
    if a.transmission:
        r.set_option("transmission", 1)
    r.set_scene(s)
    r.update_config(T.SceneConfig(**(dict(sun_dir=tuple(a.sun)) if a.sun else {})))
    st = r.render_spp(cam, a.validation_spp, batch_spp=a.batch_spp or a.validation_spp)
    out = "%s_%04d" % (a.validation, a.validation_spp)
    write_pfm(out, r.framebuffer())
    c = r.counters()
    print(json.dumps(dict(scene=a.scene, triangles=s.total_tris(), width=w, height=h, spp=st.spp, render_time_ms=st.render_time,
                          pfm=out + ".pfm", bvh_nodes=c["bvh_nodes"], bvh_build_ms=c["bvh_build_ms"])))
    return 0


if __name__ == "__main__":
    sys.exit(main())
