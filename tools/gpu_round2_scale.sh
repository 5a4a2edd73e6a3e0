#!/bin/bash
# strong scaling of the bench frame: bash tools/gpu_round2_scale.sh N [N ...]  (on a box with at least max(N) GPUs)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1])
    print("N=$n", "%.1f Msamples/s"%d["value"], "e2e %.1f"%d["e2e"]["value"], "ms/step %.2f"%d["ms_per_step"], d["roofline"]["stage_ms_rank0"], d["framebuffer_sha256"][:12])
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
