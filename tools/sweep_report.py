"""Prints one line per gpurun_out/sweep_*.json (value, stage times, nodes / ray) -- the tail of run_sweep.sh."""
import glob
import json
import os

for f in sorted(glob.glob("gpurun_out/sweep_*.json"), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print("%-28s FAILED (%s)" % (os.path.basename(f), e))
        continue
    r = d["roofline"]
    st = r["stage_ms_rank0"]
    print("%-28s %8.1f Msamples/s  trace %7.2f shade %6.2f other %5.2f ms  nodes/ray %.1f / %.1f" % (
        os.path.basename(f)[6:-5], d["value"], st["ms_trace"] / d["steps"], st["ms_shade"] / d["steps"], st["ms_other"] / d["steps"],
        r["per_ray"]["closest"]["nodes"], r["per_ray"]["shadow"]["nodes"]))
