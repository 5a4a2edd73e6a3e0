#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: Msamples/s of the path-tracing inner loop at 1920x1080 on the
synthetic 1M-triangle scene (BASELINE.json configs[1]: 64 spp, diffuse + GGX, sun + sky NEE).

  python bench.py --gpus 1 --steps K --warmup W              our arm (CUDA wavefront behind the C ABI)
  torchrun ... bench.py --gpus N ...                          one rank per GPU, frame sharded by interleaved row bands,
                                                              one NCCL reduce of the HDR accumulator per readback
  python bench.py --impl reference ...                        the reference's CPU path = the restated megakernel
                                                              (oracle/, all host threads) on a bounded sample

A "step" is one frame of `spp` samples per pixel (begin_frame / draw_frame / end_frame with batch_spp = spp).
`value` is device-timed with the scene resident in HBM; `e2e` goes through the public API with host buffers:
parameters in, the RGBA32F framebuffer read back to pinned host memory (and reduced over NCCL for N > 1) every step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080

# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL: "NCCL version ..." from rank 0), so fd 1 is
# pointed at stderr for everything but our own line.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--scene", default="c2", choices=["c2", "c4"], help="c2 = headline workload (configs[1]); c4 = configs[3], informational")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--wave-paths", type=int, default=0)
    ap.add_argument("--option", action="append", default=[], help="backend option name=value (rptr_cuda_set_option)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_traffic(args, overlap):
    """DRAM bytes per closest-hit launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the launches of a
    frame) from the committed ncu capture of this same command, or None when the workload is not the captured one."""
    p = os.path.join(ROOT, "profiles", "r01_trace_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    w = t["workload"]
    same = (args.scene == w["scene"] and args.tris == w["tris"] and args.spp == w["spp"] and args.width == w["width"] and
            args.height == w["height"] and not args.wave_paths and not args.option and args.gpus == 1)
    if not same:
        return None, None
    if overlap:  # per launch over the closest-hit AND shadow launches, like `achieved`
        n = t["closest"]["launches"] + t["shadow"]["launches"]
        tot = t["closest"]["avg_dram_bytes_per_launch"] * t["closest"]["launches"] + t["shadow"]["avg_dram_bytes_per_launch"] * t["shadow"]["launches"]
        return tot / n, "profiles/r01_trace_traffic.json (closest + shadow launches)"
    return t["closest"]["avg_dram_bytes_per_launch"], "profiles/r01_trace_traffic.json"


def make_scene(args):
    from realtimepathtracingresearchframework_b200 import scenes
    if args.scene == "c4":  # BASELINE configs[3]: 100 k-triangle mesh x 100 instances, full BSDF set + area-light NEE
        return scenes.instanced_scene(100_000, 100)
    return scenes.random_triangles(args.tris)


def workload_name(args):
    if args.scene == "c4":
        return "synthetic 10M-triangle instanced scene (100k mesh x 100), %dx%d, %d spp, GGX+transmission+emissive tri-light NEE (BASELINE configs[3])" % (
            args.width, args.height, args.spp)
    return "synthetic %d random-triangle scene, %dx%d, %d spp, diffuse+GGX, sun+sky NEE (BASELINE configs[1])" % (
        args.tris, args.width, args.height, args.spp)


# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline(args, scene, seconds):
    """The restated reference megakernel (oracle/) on this box's host cores, on a bounded sample of the same workload."""
    from oracle import pyoracle as po
    from realtimepathtracingresearchframework_b200 import load_sky_fit
    o = po.OracleScene(scene)
    sp = load_sky_fit()
    cores = po.lib().oracle_num_threads()
    # probe: a 1920 x 32 band in the middle of the frame, 1 spp
    y0 = args.height // 2 - 16
    t0 = time.perf_counter()
    o.render(args.width, args.height, scene.camera, sp, spp=1, region=(0, y0, args.width, y0 + 32))
    probe = time.perf_counter() - t0
    rate = args.width * 32 / probe
    # bounded sample: full-width bands spread over the frame so sky and geometry rows are both represented; once the
    # whole frame fits the budget, more samples per pixel instead
    rows = int(max(32, min(args.height, seconds * rate / args.width)))
    rows -= rows % 8
    spp = 1
    if rows >= args.height - 8:
        rows = args.height - args.height % 8
        spp = int(max(1, min(args.spp, seconds * rate / (args.width * args.height))))
    step = args.height / (rows / 8)
    samples, t = 0, 0.0
    img = np.zeros((args.height, args.width, 4), np.float32)
    t0 = time.perf_counter()
    for b in range(rows // 8):
        ys = int(b * step)
        o.render(args.width, args.height, scene.camera, sp, spp=spp, region=(0, ys, args.width, min(ys + 8, args.height)), out=img)
        samples += args.width * (min(ys + 8, args.height) - ys) * spp
    t = time.perf_counter() - t0
    return {"value": samples / t / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": "%d rows (bands of 8 spread over the frame) x %d px x %d spp of the same scene/camera, %.1f s; "
                      "restated reference megakernel, OpenMP" % (rows, args.width, spp, t)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = make_scene(args)
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args, scene, max(2.0, args.cpu_seconds / max(1, args.steps)))
        if i >= args.warmup:
            vals.append(base["value"])
        if i == 0 and args.warmup > 1:
            pass
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args)}, "cpu_baseline": base,
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from realtimepathtracingresearchframework_b200 import RenderConfiguration, RenderCuda, load_sky_fit, types as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene = make_scene(args)
    r = RenderCuda(device=local)
    r.initialize(args.width, args.height)
    r.set_option("stage_timing", 1)
    if args.scene == "c4":
        r.set_option("transmission", 1)
        r.set_option("bvh_builder", 1)
    if args.wave_paths:
        r.set_option("wave_paths", args.wave_paths)
    for kv in args.option:
        k, v = kv.split("=")
        r.set_option(k, int(v))
    if world > 1:
        r.set_option("tile_world", world)
        r.set_option("tile_rank", rank)
        r.set_option("tile_rows", 8)
    t0 = time.perf_counter()
    r.set_scene(scene)
    scene_s = time.perf_counter() - t0
    r.update_config(T.SceneConfig())
    r.params.batch_spp = args.spp
    cam = scene.camera

    stream = torch.cuda.ExternalStream(r.stream_handle(), device=torch.device("cuda", local))
    n_px = args.width * args.height
    # framebuffer as a torch view of the library's accumulator (NCCL reduce) + pinned host buffer for the readback
    fb_ptr = r.framebuffer_device_ptr()

    class _Arr:
        __cuda_array_interface__ = {"shape": (n_px * 4,), "typestr": "<f4", "data": (fb_ptr, False), "version": 3}
    fb = torch.as_tensor(_Arr(), device=torch.device("cuda", local))
    reduced = torch.empty_like(fb) if world > 1 else None
    host = torch.empty(n_px * 4, dtype=torch.float32, pin_memory=True)

    def step(readback):
        cfg = RenderConfiguration(cam, reset_accumulation=True)
        r.begin_frame(None, cfg)
        r.draw_frame()
        r.end_frame()
        if readback:
            with torch.cuda.stream(stream):
                src = fb
                if world > 1:  # disjoint row bands: sum == gather, exact in fp32
                    reduced.copy_(fb)
                    dist.reduce(reduced, dst=0, op=dist.ReduceOp.SUM)
                    src = reduced
                if rank == 0:
                    host.copy_(src, non_blocking=True)
            stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step(True)
    # ---- device-timed region: K steps, inputs resident ----
    barrier()
    r.reset_counters()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step(False)
    ev1.record(stream)
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    cnt = r.counters()
    clk = clocks.stop() if rank == 0 else None
    # ---- end-to-end region: public API, host buffers, readback (and reduce) inside ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)

    total_samples = float(n_px) * args.spp * args.steps  # all ranks together render every pixel once per step
    value = total_samples / (dev_ms * 1e-3) / 1e6
    e2e = total_samples / (e2e_ms * 1e-3) / 1e6
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (closest-hit trace): algorithmic bytes / event-timed duration ----
    peak, peak_src = peaks()
    node_b, tri_b = cnt["node_bytes"], cnt["tri_bytes"]
    closest_bytes = cnt["closest_rays"] * (32 + 32) + cnt["closest_nodes"] * node_b + cnt["closest_tris"] * tri_b
    shadow_bytes = cnt["shadow_rays"] * (32 + 4) + cnt["shadow_nodes"] * node_b + cnt["shadow_tris"] * tri_b
    # With option overlap_shadow (default) every shadow launch shares the GPU with the next closest-hit launch, so the two
    # instantiations of the trace kernel are timed together (ms_trace = union of their intervals) and the roofline is
    # stated for both: algorithmic bytes of closest-hit + shadow rays over that time.
    overlap = bool(cnt.get("trace_overlap", 0))
    trace_bytes = closest_bytes + (shadow_bytes if overlap else 0)
    trace_rays = cnt["closest_rays"] + (cnt["shadow_rays"] if overlap else 0)
    ach = trace_bytes / (cnt["ms_trace"] * 1e-3) / 1e9 if cnt["ms_trace"] > 0 else None
    stage_ms = {k: cnt[k] for k in ("ms_trace", "ms_shade", "ms_shadow", "ms_other")}
    spl = max(1, cnt["samples"])
    traffic, traffic_src = measured_traffic(args, overlap)
    kernel_name = "k_trace_persistent (closest-hit + any-hit launches, overlapped on two streams)" if overlap else "k_trace_persistent (closest hit)"
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "per_launch": {"launches": cnt["trace_launches"], "avg_ms": cnt["ms_trace"] / max(1, cnt["trace_launches"]),
                               "avg_algorithmic_bytes": trace_bytes / max(1, cnt["trace_launches"])},
                "per_ray": {"closest": {"nodes": cnt["closest_nodes"] / max(1, cnt["closest_rays"]), "tris": cnt["closest_tris"] / max(1, cnt["closest_rays"]),
                                        "bytes": closest_bytes / max(1, cnt["closest_rays"])},
                            "shadow": {"nodes": cnt["shadow_nodes"] / max(1, cnt["shadow_rays"]), "tris": cnt["shadow_tris"] / max(1, cnt["shadow_rays"]),
                                       "bytes": shadow_bytes / max(1, cnt["shadow_rays"])},
                            "bytes": trace_bytes / max(1, trace_rays), "node_bytes": node_b, "tri_bytes": tri_b},
                "per_sample": {"closest_rays": cnt["closest_rays"] / spl, "shadow_rays": cnt["shadow_rays"] / spl, "vertices": cnt["shaded_vertices"] / spl},
                "stage_ms_rank0": stage_ms, "mrays_per_s": (cnt["closest_rays"] + cnt["shadow_rays"]) / (dev_ms * 1e-3) / 1e6 * world}
    line = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "parallelism": "screen tiles x%d (interleaved 8-row bands), scene replicated" % world,
                       "l2": "no explicit flush: per-wave path state (>1 GB) and scene+BVH (%.0f MB) exceed the 126 MB L2" % (args.tris * 110e-6),
                       "scene_setup_s": scene_s},
            "e2e": {"value": e2e, "unit": "Msamples/s", "h2d_bytes_per_step": C.sizeof(T.RenderCameraParams) + C.sizeof(T.RenderParams) + C.sizeof(T.LightSamplingConfig),
                    "d2h_bytes_per_step": n_px * 16, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(cnt["launches"]), "clocks": clk, "roofline": roofline}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args, scene, args.cpu_seconds)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
