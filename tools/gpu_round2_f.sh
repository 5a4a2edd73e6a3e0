#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli.py tests/test_ref_path.py -m gpu -x -q 2>&1 | tail -3
for b in 0 1; do timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --option bvh_builder=$b > gpurun_out/bld_c2_$b.json 2>/dev/null; done
for b in 0 1; do timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --scene c4 --spp 16 --option bvh_builder=$b > gpurun_out/bld_c4_$b.json 2>/dev/null; done
python - <<'PY'
import json
for f in ("bld_c2_0","bld_c2_1","bld_c4_0","bld_c4_1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "%.1f Msamples/s"%d["value"], "setup %.2fs"%d["config"]["scene_setup_s"], "nodes/ray %.1f / %.1f"%(r["per_ray"]["closest"]["nodes"], r["per_ray"]["shadow"]["nodes"]), "tris %.2f"%r["per_ray"]["closest"]["tris"])
    except Exception as e: print(f, "failed", e)
PY
