#!/bin/bash
# Round 2, visit C: the eight-wide BVH + group-stack traversal kernel: parity tests, bench C2 / C4, full ncu of the first launches
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/t_c.log 2>&1; tail -12 gpurun_out/t_c.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_bvh8.json 2> gpurun_out/bench_c2_bvh8.err; tail -c 1800 gpurun_out/bench_c2_bvh8.json; tail -3 gpurun_out/bench_c2_bvh8.err
timeout 900 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4_bvh8.json 2> gpurun_out/bench_c4_bvh8.err; tail -c 1500 gpurun_out/bench_c4_bvh8.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 3 -f -o gpurun_out/trace_full_c2_bvh8 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/b_ncu_c2f.log 2>&1
