#!/bin/bash
# Round 2, visit S: 64-byte seven-child nodes against the 96-byte eight-child nodes (variants/librptr_cuda_node96.so = the previous build), full GPU suite
mkdir -p gpurun_out
run() { # name, lib, bench args...
  local name=$1 lib=$2; shift; shift
  RPTR_CUDA_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2 "$@" > gpurun_out/n64_$name.json 2> gpurun_out/n64_$name.err
}
NEW=realtimepathtracingresearchframework_b200/librptr_cuda.so
OLD=variants/librptr_cuda_node96.so
run c2_new $NEW
run c2_old $OLD
run c4_new $NEW --scene c4 --spp 16
run c4_old $OLD --scene c4 --spp 16
run c2_new_host $NEW --option bvh_builder=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n64_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
        print("%-12s %.1f Msamples/s e2e %.1f"%(f[15:-5], d["value"], d["e2e"]["value"]), r["stage_ms_rank0"], d["framebuffer_sha256"][:12], "per ray", r.get("per_ray"), "setup", d["config"].get("scene_setup_s"))
    except Exception as e: print(f, "failed", e); print(open(f[:-5]+".err").read()[-800:])
PY
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/t_s.log 2>&1; tail -14 gpurun_out/t_s.log
