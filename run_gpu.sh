N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; python -c "
import json; j=json.load(open('gpurun_out/bench_${N}gpu.json')); print('scale', j['n_gpus'], j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['stage_ms_rank0'])"; tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-200
