#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "not full_frame and not c3_ and not c4_full" ) > gpurun_out/t_d.log 2>&1; tail -6 gpurun_out/t_d.log
bash run_sweep.sh
for cw in 1 2; do timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --option concurrent_waves=$cw > gpurun_out/cw_$cw.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/cw_$cw.json").read().strip().splitlines()[-1]); print("concurrent_waves $cw: %.1f Msamples/s e2e %.1f  %s" % (d["value"], d["e2e"]["value"], d["framebuffer_sha256"][:12]))
PY
done
for cw in 1 2; do timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --spp 8 --option concurrent_waves=$cw > gpurun_out/cw8_$cw.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/cw8_$cw.json").read().strip().splitlines()[-1]); print("8 spp concurrent_waves $cw: %.1f Msamples/s e2e %.1f" % (d["value"], d["e2e"]["value"]))
PY
done
