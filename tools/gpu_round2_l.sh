#!/bin/bash
# Round 2, visit L: full GPU suite after the f2 / f3 work, C2 + C4 benches (no regression: scenes without image textures run the same kernels)
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/t_l.log 2>&1; tail -14 gpurun_out/t_l.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_l.json 2> gpurun_out/bench_c2_l.err
timeout 900 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4_l.json 2> gpurun_out/bench_c4_l.err
python - <<'PY'
import json
for f in ("bench_c2_l","bench_c4_l"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "%.1f Msamples/s e2e %.1f"%(d["value"], d["e2e"]["value"]), r["stage_ms_rank0"], d["framebuffer_sha256"][:12], "setup", d["config"]["scene_setup_s"])
    except Exception as e: print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
