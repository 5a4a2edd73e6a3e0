#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lbvh or c4_instanced or smooth or textured or c4_full_size_prop" 2>&1 | tail -4
for b in 1; do timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --option bvh_builder=$b > gpurun_out/bld_c2_$b.json 2>gpurun_out/bld_c2_$b.err; done
for b in 1; do timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --scene c4 --spp 16 --option bvh_builder=$b > gpurun_out/bld_c4_$b.json 2>gpurun_out/bld_c4_$b.err; done
python - <<'PY'
import json
for f in ("bld_c2_1","bld_c4_1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "%.1f Msamples/s"%d["value"], "setup %.2fs"%d["config"]["scene_setup_s"], "nodes/ray %.1f / %.1f"%(r["per_ray"]["closest"]["nodes"], r["per_ray"]["shadow"]["nodes"]), "tris %.2f"%r["per_ray"]["closest"]["tris"])
    except Exception as e: print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-600:])
PY
python - <<'PY'
import time
from realtimepathtracingresearchframework_b200 import RenderCuda, scenes
s = scenes.random_triangles(1000000)
r = RenderCuda(device=0); r.initialize(64,64); r.set_option("bvh_builder",1)
for i in range(3):
    t=time.time(); r.set_scene(s); dt=time.time()-t
    print("set_scene 1M tris, device builder: %.3f s total, bvh_build_ms %.1f, nodes %d"%(dt, r.counters()["bvh_build_ms"], r.counters()["bvh_nodes"]))
PY
