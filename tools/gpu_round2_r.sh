#!/bin/bash
# Round 2, visit R: ray reordering between shade and trace (rptr_reorder.cuh) -- key modes for the bounce / shadow queues, C2 and C4
mkdir -p gpurun_out
run() { # name, bench args...
  local name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 1 "$@" > gpurun_out/ro_$name.json 2> gpurun_out/ro_$name.err
}
run c2_base
run c2_s3 --option reorder_shadow=3
run c2_s1 --option reorder_shadow=1
run c2_s2 --option reorder_shadow=2
run c2_b1 --option reorder_bounce=1
run c2_b2 --option reorder_bounce=2
run c2_b4 --option reorder_bounce=4
run c2_s3b2 --option reorder_shadow=3 --option reorder_bounce=2
run c2_s3b1 --option reorder_shadow=3 --option reorder_bounce=1
run c4_base --scene c4 --spp 16
run c4_s2b2 --scene c4 --spp 16 --option reorder_shadow=2 --option reorder_bounce=2
run c4_s1b1 --scene c4 --spp 16 --option reorder_shadow=1 --option reorder_bounce=1
run c4_s4b4 --scene c4 --spp 16 --option reorder_shadow=4 --option reorder_bounce=4
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ro_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
        print("%-12s %.1f Msamples/s e2e %.1f"%(f[14:-5], d["value"], d["e2e"]["value"]), r["stage_ms_rank0"], d["framebuffer_sha256"][:12])
    except Exception as e: print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
