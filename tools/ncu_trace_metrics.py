#!/usr/bin/env python
"""Turns the per-launch ncu CSV of the trace kernel into the JSON bench.py reads for roofline.traffic / dram_frac /
issue_frac (it cannot measure those itself: hardware counters need the profiler, and a number printed under a profiler is
never a bench value -- so the counters come from one committed capture of the same command and the time from the live run).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,\
smsp__thread_inst_executed.sum,sm__cycles_elapsed.max --clock-control none -k regex:k_trace_persistent -c 40 --csv \
      --log-file gpurun_out/trace_metrics.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline [--scene c4 --spp 16]
  python tools/ncu_trace_metrics.py gpurun_out/trace_metrics.csv --scene c2 --tris 1000000 --spp 64 > profiles/r02_trace_metrics_c2.json

bench.py --steps 1 --warmup 0 renders three frames (hash frame, device-timed step, end-to-end step); counters per launch are
averaged over all captured launches of each instantiation (closest: <0, *>, shadow: <1, *>).
"""
import argparse
import csv
import json
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--scene", default="c2")
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--command", default="")
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.csv)) if len(r) > 10]
    hdr = rows[0]
    ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    out = {"source": a.command or "ncu per-launch metrics of k_trace_persistent under python bench.py --steps 1 --warmup 0 --no-cpu-baseline",
           "workload": {"scene": a.scene, "tris": a.tris, "spp": a.spp, "width": a.width, "height": a.height}}
    for name, tag in (("closest", "k_trace_persistent<0"), ("shadow", "k_trace_persistent<1")):
        ls = [d for _, d in sorted(launches.items()) if tag in d["kernel"].replace("(bool)", "").replace(" ", "")]
        if not ls:
            continue
        n = len(ls)
        dram = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in ls]
        out[name] = {
            "launches": n,
            "avg_dram_bytes_per_launch": sum(dram) / n,
            "avg_ms_under_ncu": sum(d.get("gpu__time_duration.sum", 0.0) for d in ls) / n * 1e-6,
            "avg_warp_inst_per_launch": sum(d.get("smsp__inst_executed.sum", 0.0) for d in ls) / n,
            "avg_thread_inst_per_launch": sum(d.get("smsp__thread_inst_executed.sum", 0.0) for d in ls) / n,
            "avg_cycles_under_ncu": sum(d.get("sm__cycles_elapsed.max", 0.0) for d in ls) / n,
            "per_launch_dram_mb": [round(x / 1e6, 1) for x in dram],
            "per_launch_ms": [round(d.get("gpu__time_duration.sum", 0.0) * 1e-6, 3) for d in ls],
        }
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
