#!/bin/bash
# One GPU-box visit: parity tests, headline bench, ncu launch list, ncu --set full of the trace and shade kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1500 gpurun_out/bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -f -o gpurun_out/trace_full \
    python bench.py --steps 1 --warmup 0 --spp 4 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 1 -f -o gpurun_out/shade_full \
    python bench.py --steps 1 --warmup 0 --spp 4 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out
