#!/bin/bash
# Round 2, visit O: ncu --set full of the final kernels: trace (closest, shadow) on C2 and C4, shade on C2 and C4
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 4 -f -o gpurun_out/full_trace_c2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/o1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 4 -f -o gpurun_out/full_trace_c4 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --scene c4 --spp 8 > gpurun_out/o2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 2 -f -o gpurun_out/full_shade_c2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/o3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 2 -f -o gpurun_out/full_shade_c4 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --scene c4 --spp 8 > gpurun_out/o4.log 2>&1
ls -la gpurun_out/full_*.ncu-rep
