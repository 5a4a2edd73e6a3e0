"""host/rptr_cuda_cli: the headless C++ driver over the C ABI that replays `rptr --backend cuda --validation <prefix>
--validation-spp N --pfm --profiling <name>` (libapp/app_state.cpp:464-481, libapp/benchmark_info.cpp:69-124)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from realtimepathtracingresearchframework_b200 import build, scenes

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cli():
    return build.build_cli()


def fnv(chunks):
    h = 1469598103934665603
    for b in chunks:
        for x in b:
            h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def scene_hash(s):
    ch = []
    for g in s.geometries:
        ch += [g.qverts.tobytes(), np.asarray(g.scaling, np.float32).tobytes(), np.asarray(g.offset, np.float32).tobytes()]
    ch += [bytes(m) for m in s.materials]
    ch.append(bytes(s.camera))
    return "%016x" % fnv(ch)


def test_cli_scenes_equal_the_python_generators(cli):
    """The driver generates BASELINE's scenes itself (C++); they must be the scenes of scenes.py byte for byte."""
    for name, s in (("cornell", scenes.cornell_box()), ("random:20000", scenes.random_triangles(20000))):
        out = subprocess.run([cli, "--scene", name, "--scene-hash"], capture_output=True, text=True, check=True).stdout.strip()
        assert out == scene_hash(s), name


def test_cli_usage_and_loud_failure_without_a_device(cli, tmp_path):
    assert "usage: rptr_cuda_cli" in subprocess.run([cli, "--help"], capture_output=True, text=True, check=True).stdout
    r = subprocess.run([cli, "--scene", "cornell"], capture_output=True, text=True)
    assert r.returncode != 0 and "--validation" in r.stderr
    r = subprocess.run([cli, "--backend", "vulkan", "--validation", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode != 0
    import torch
    if not torch.cuda.is_available():  # no CPU fallback: the driver says so and writes nothing
        r = subprocess.run([cli, "--validation", str(tmp_path / "x"), "--img", "64", "36"], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
        assert not list(tmp_path.iterdir())


@pytest.mark.gpu
def test_cli_validation_run_matches_the_python_harness(cli, tmp_path, oracle):
    from realtimepathtracingresearchframework_b200 import RenderCuda, load_sky_fit, read_pfm, types as T
    for name, s, (w, h), spp, batch, sky in (("cornell", scenes.cornell_box(), (256, 144), 5, 2, "default"),
                                           ("random:20000", scenes.random_triangles(20000), (200, 120), 4, 4, "slanted")):
        prefix = str(tmp_path / name.replace(":", "_"))
        r = subprocess.run([cli, "--backend", "cuda", "--disable-ui", "--scene", name, "--img", str(w), str(h), "--validation", prefix, "--validation-spp", str(spp),
                            "--batch-spp", str(batch), "--pfm", "--profiling", prefix + "_bench", "--sky", sky], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = read_pfm("%s_%04d.pfm" % (prefix, spp))
        cfg = T.SceneConfig(sun_dir=(0.35, 0.8, 0.45)) if sky == "slanted" else T.SceneConfig()
        b = RenderCuda(device=0)
        b.initialize(w, h)
        b.set_scene(s)
        b.update_config(cfg)
        b.render_spp(s.camera, spp, batch_spp=batch)
        want = b.framebuffer()
        assert np.array_equal(got.view(np.uint32), want[..., :3].view(np.uint32))
        ref, _ = oracle.OracleScene(s).render(w, h, s.camera, load_sky_fit(cfg), spp=spp, batch_spp=batch)
        assert np.array_equal(got.view(np.uint32), ref[..., :3].view(np.uint32))
        rows = open(prefix + "_bench.csv").read().strip().splitlines()
        assert rows[0] == "frames_total,keyframe,frames_accumulated,render_time_ms,app_time_ms"
        frames = -(-spp // batch)
        assert len(rows) == 1 + frames
        last = rows[-1].split(",")
        assert int(last[0]) == frames and int(last[2]) == spp and float(last[3]) > 0


def vks_scene_hash(s):
    """VksScene::hash (host/vks_loader.cpp) of a scene loaded by vks.load_vks"""
    d = s.desc()
    ch = []
    for g in s.geometries:
        ch += [g.qverts.tobytes(), g.qnormal_uv.tobytes(), np.asarray(g.scaling, np.float32).tobytes(), np.asarray(g.offset, np.float32).tobytes()]
    for pm in s.pmeshes:
        ch.append(np.asarray(pm["material_offsets"], np.int32).tobytes())
        if pm["tri_material_ids"] is not None:
            ch.append(pm["tri_material_ids"].tobytes())
    for pm_id, tr in s.instances:
        ch += [np.int32(pm_id).tobytes(), (np.asarray(tr, np.float32) + np.float32(0.0)).tobytes()]   # -0 and +0 are the same transform
    ch += [bytes(m) for m in s.materials]
    for i in range(d.n_textures):
        t = d.textures[i]
        ch.append(np.array([t.width, t.height, t.channels, t.color_space, t.bc_format, t.mip_levels], np.int32).tobytes())
        entry = s.textures[i]
        ch.append(np.ascontiguousarray(entry[2]["blob"]).tobytes() if isinstance(entry[0], str) else np.ascontiguousarray(entry[0], np.uint8).tobytes())
    return "%016x" % fnv(ch)


def test_cli_reads_vks_files_like_the_python_loader(cli, tmp_path):
    """host/vks_loader.cpp (the C++ twin of vks.py): every table it builds from a .vks file + texture directory -- geometry streams,
    material offsets / per-triangle ids, base-LoD instances with dequantised transforms, materials with texture handles and parameter
    files, texture headers and payloads -- hashes to the same value as the Python loader's scene."""
    import vks_util
    from realtimepathtracingresearchframework_b200 import vks
    path, _ = vks_util.write_test_scene(str(tmp_path))
    out = subprocess.run([cli, "--scene", path, "--scene-hash"], capture_output=True, text=True, check=True).stdout.strip()
    assert out == vks_scene_hash(vks.load_vks(path))
    r = subprocess.run([cli, "--scene", path, "--validation", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode != 0 and "--eye" in r.stderr            # the file has no camera
    bad = tmp_path / "bad.vks"
    bad.write_bytes(open(path, "rb").read()[:300])
    r = subprocess.run([cli, "--scene", str(bad), "--scene-hash"], capture_output=True, text=True)
    assert r.returncode != 0 and ("truncated" in r.stderr or "mismatching" in r.stderr)


@pytest.mark.gpu
def test_cli_renders_a_vks_file(cli, tmp_path, oracle):
    import vks_util
    from realtimepathtracingresearchframework_b200 import load_sky_fit, read_pfm, types as T, vks
    path, _ = vks_util.write_test_scene(str(tmp_path))
    prefix = str(tmp_path / "yard")
    r = subprocess.run([cli, "--scene", path, "--img", "160", "90", "--validation", prefix, "--validation-spp", "3", "--eye", "0", "2", "14", "--fovy", "50",
                        "--transmission"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s = vks.load_vks(path)
    cam = scenes.look_at_camera((0, 2, 14), (0, 0, 0), fovy=50.0)
    ref, _ = oracle.OracleScene(s).render(160, 90, cam, load_sky_fit(T.SceneConfig()), spp=3, transmission=1)
    assert np.array_equal(read_pfm(prefix + "_0003.pfm"), ref[..., :3])
