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


def measured_counters(args, overlap):
    """Hardware counters of the trace kernel per launch -- DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), warp and
    thread instructions (smsp__inst_executed.sum, smsp__thread_inst_executed.sum) -- from the committed ncu capture of this same
    command (tools/ncu_trace_metrics.py), or None when the workload is not a captured one.  Hardware counters need the profiler,
    and nothing timed under a profiler is a bench value: the counters are deterministic for a fixed scene / seed, the time is
    this run's."""
    for name in ("r02_trace_metrics_%s.json" % args.scene, "r01_trace_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        with open(p) as f:
            t = json.load(f)
        w = t["workload"]
        same = (args.scene == w["scene"] and args.tris == w["tris"] and args.spp == w["spp"] and args.width == w["width"] and
                args.height == w["height"] and not args.wave_paths and not args.option and args.gpus == 1)
        if not same:
            continue
        kinds = ["closest", "shadow"] if overlap else ["closest"]  # per launch over the same launches as `achieved`
        n = sum(t[k]["launches"] for k in kinds)

        def avg(key):
            if any(key not in t[k] for k in kinds):
                return None
            return sum(t[k][key] * t[k]["launches"] for k in kinds) / n
        return {"dram_bytes": avg("avg_dram_bytes_per_launch"), "warp_inst": avg("avg_warp_inst_per_launch"),
                "thread_inst": avg("avg_thread_inst_per_launch"), "source": "profiles/%s (ncu, %s launches)" % (name, " + ".join(kinds))}
    return None


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
def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed on.  Passed to the oracle explicitly, so a launcher
    that exports OMP_NUM_THREADS=1 (torch.distributed.run does) cannot throttle the baseline."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class CpuArm:
    """The restated reference megakernel (oracle/) on this box's host cores, on a bounded sample of the same workload:
    the same scene, camera, frame size and samples per pixel as the GPU arm, on a subset of the frame's rows -- full-width
    bands of 16 rows spread evenly over the frame so that sky and geometry rows are both represented.  Every pixel sample is
    an independent path with its own seed, so Msamples/s of the subset is the rate of the whole frame."""

    BAND = 16

    def __init__(self, args, scene):
        from oracle import pyoracle as po
        from realtimepathtracingresearchframework_b200 import load_sky_fit
        self.po, self.args, self.scene = po, args, scene
        self.o = po.OracleScene(scene)
        self.sp = load_sky_fit()
        self.threads = host_threads()
        self.img = np.zeros((args.height, args.width, 4), np.float32)
        self.rate = None  # samples / s, refined by every run

    def _render(self, ys, rows, spp):
        a = self.args
        t0 = time.perf_counter()
        self.o.render(a.width, a.height, self.scene.camera, self.sp, spp=spp, region=(0, ys, a.width, min(ys + rows, a.height)), out=self.img,
                      n_threads=self.threads)
        return time.perf_counter() - t0

    def run(self, seconds):
        a = self.args
        if self.rate is None:  # probe: one band in the middle of the frame, 1 spp
            dt = self._render(a.height // 2 - self.BAND // 2, self.BAND, 1)
            self.rate = a.width * self.BAND / dt
        per_band = a.width * self.BAND * a.spp
        bands = int(max(1, min(a.height // self.BAND, seconds * self.rate / per_band)))
        step = a.height / bands
        samples, t = 0, 0.0
        for b in range(bands):
            ys = min(int(b * step + 0.5 * (step - self.BAND)), a.height - self.BAND)
            ys = max(ys, 0)
            t += self._render(ys, self.BAND, a.spp)
            samples += a.width * (min(ys + self.BAND, a.height) - ys) * a.spp
        self.rate = samples / t
        used = int(self.po.lib().oracle_last_threads())
        return {"value": samples / t / 1e6, "unit": "Msamples/s", "cores": used, "kind": "port", "seconds": t,
                "sample": "%d of %d rows (%d bands of %d rows spread over the frame) x %d px x %d spp = the GPU arm's scene, camera, "
                          "frame size and spp on a subset of its pixels (samples are independent paths: the rate of the subset is the "
                          "rate of the frame), %.1f s on %d OpenMP threads; restated reference megakernel (the reference has no CPU path)"
                          % (bands * self.BAND, a.height, bands, self.BAND, a.width, a.spp, t, used)}


def cpu_baseline(args, scene, seconds):
    return CpuArm(args, scene).run(seconds)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = make_scene(args)
    arm = CpuArm(args, scene)
    vals, secs = [], []
    base = None
    budget = max(2.0, args.cpu_seconds * 4 / max(1, args.steps + args.warmup))  # whole run ~ a minute of CPU work
    for i in range(args.warmup + args.steps):
        base = arm.run(budget)
        if i >= args.warmup:
            vals.append(base["value"])
            secs.append(base["seconds"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
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
    if args.wave_paths:
        r.set_option("wave_paths", args.wave_paths)
    for kv in args.option:
        k, v = kv.split("=")
        r.set_option(k, int(v))
    if world > 1:
        # the product's own communicator (NCCL inside librptr_cuda.so): rank 0's id travels over torch.distributed, which is
        # plumbing here (barriers, the max over ranks of the timings)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(RenderCuda.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        r.comm_init_rank(world, rank, bytes(uid.cpu().numpy().tobytes()))  # sets tile_world / tile_rank
        r.set_option("tile_rows", 8)
    t0 = time.perf_counter()
    r.set_scene(scene)
    scene_s = time.perf_counter() - t0
    r.update_config(T.SceneConfig())
    r.params.batch_spp = args.spp
    cam = scene.camera

    stream = torch.cuda.ExternalStream(r.stream_handle(), device=torch.device("cuda", local))
    n_px = args.width * args.height
    host = torch.empty(n_px * 4, dtype=torch.float32, pin_memory=True)  # pinned host buffer the framebuffer is read back into
    host_np = host.numpy()

    def step(readback):
        cfg = RenderConfiguration(cam, reset_accumulation=True)
        r.begin_frame(None, cfg)
        r.draw_frame()
        r.end_frame()
        if readback:
            if world > 1:  # one NCCL reduce of the HDR accumulator to rank 0 (disjoint row bands: sum == gather, exact in fp32)
                r.reduce_framebuffer(0)
            if rank == 0:
                if r.readback_framebuffer(host_np) != host_np.size:  # device -> pinned host, synchronises the stream
                    raise RuntimeError("readback failed: " + r.last_error())
            else:
                r.flush_pipeline()

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

    # ---- hash frame: the very first frame after set_scene (frame_id = 0, frame_offset = 0) is the image the parity tests hold
    # against the oracle (tests/test_gpu_parity.py::test_c2_bench_configuration_full_frame); its SHA-256 goes into the line and
    # must be the same for every N (the sharded frame after the reduce is bit-identical to the 1-GPU frame)
    import hashlib
    step(True)
    fb_sha = hashlib.sha256(host_np.tobytes()).hexdigest() if rank == 0 else None
    for _ in range(max(0, args.warmup - 1)):
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
    hw = measured_counters(args, overlap)
    kernel_name = "k_trace_persistent (closest-hit + any-hit launches, overlapped on two streams)" if overlap else "k_trace_persistent (closest hit)"
    avg_s = cnt["ms_trace"] * 1e-3 / max(1, cnt["trace_launches"])
    sm_hz = ((clk or {}).get("sm_mhz") or (clk or {}).get("sm_max_mhz") or 1965.0) * 1e6
    issue_slots = avg_s * cnt["num_sms"] * 4 * sm_hz  # warp-instruction issue slots per launch: SMs x 4 schedulers x cycles
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": hw["dram_bytes"] if hw else None, "traffic_source": hw["source"] if hw else None, "peak_source": peak_src,
                # what `achieved` is: ALGORITHMIC bytes (ray + hit records + 64 B per node visit + 48 B per triangle test) over
                # the live launch time.  Scene + BVH fit the 126 MB L2, so most of those bytes never reach HBM: dram_frac is the
                # measured DRAM rate (ncu dram__bytes per launch over the live launch time) against the same peak, and the limiter
                # the profile shows is instruction issue: issue_frac = warp instructions per launch / issue slots of the launch,
                # lane_adjusted_issue_frac = thread instructions / (32 x issue slots).
                "achieved_is": "algorithmic bytes / live launch time (served mostly from L2, see dram_frac)",
                "dram_frac": (hw["dram_bytes"] / avg_s / 1e9 / peak) if hw and hw["dram_bytes"] and avg_s > 0 else None,
                "dram_gbs": (hw["dram_bytes"] / avg_s / 1e9) if hw and hw["dram_bytes"] and avg_s > 0 else None,
                "issue_frac": (hw["warp_inst"] / issue_slots) if hw and hw["warp_inst"] and issue_slots > 0 else None,
                "lane_adjusted_issue_frac": (hw["thread_inst"] / (32.0 * issue_slots)) if hw and hw["thread_inst"] and issue_slots > 0 else None,
                "binding": "L1 data-pipe wavefronts (one per scattered 32-byte sector: 70-84 % of peak) and the ALU pipe (64-68 %), ncu: "
                           "profiles/r02_full_trace_*_final.txt; the scene is L2-resident, HBM traffic is the ray / hit records",
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
            "config": {"workload": workload_name(args), "parallelism": "screen tiles x%d (interleaved 8-row bands), scene replicated, one NCCL reduce of the accumulator per readback inside librptr_cuda.so" % world,
                       "l2": "no explicit flush: per-wave path state (>1 GB) and scene+BVH (%.0f MB) exceed the 126 MB L2" % (args.tris * 110e-6),
                       "scene_setup_s": scene_s},
            "e2e": {"value": e2e, "unit": "Msamples/s", "h2d_bytes_per_step": C.sizeof(T.RenderCameraParams) + C.sizeof(T.RenderParams) + C.sizeof(T.LightSamplingConfig),
                    "d2h_bytes_per_step": n_px * 16, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(cnt["launches"]), "clocks": clk, "roofline": roofline,
            "framebuffer_sha256": fb_sha, "framebuffer_sha256_of": "first frame after set_scene (frame_id 0, frame_offset 0), RGBA32F as read back on rank 0"}
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
