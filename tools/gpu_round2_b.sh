#!/bin/bash
# Round 2, visit B (2 GPUs): new boundary tests (ray queries, trace_rays kernels, options, in-library NCCL reduce) + 2-GPU bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "ray_queries or trace_rays or options or nccl or two_devices or smooth or error_behaviour or ray_query" ) > gpurun_out/t_b.log 2>&1; tail -15 gpurun_out/t_b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1200 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.json
