#!/bin/bash
# strong scaling of the bench frame: N = 8, 4 (N = 1, 2 were measured in earlier visits)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1])
    print("N=$n", "%.1f Msamples/s"%d["value"], "e2e %.1f"%d["e2e"]["value"], "ms/step %.2f"%d["ms_per_step"], d["roofline"]["stage_ms_rank0"], d["framebuffer_sha256"][:12], d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
