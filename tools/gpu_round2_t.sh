#!/bin/bash
# Round 2, visit T: source-level ncu capture of the trace kernel (where do the stalls sit?), then the full GPU suite
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo SMOKE FAILED; tail -5 gpurun_out/smoke.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 3 -f -o gpurun_out/src_trace_c2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/t1.log 2>&1
ls -la gpurun_out/src_trace_c2.ncu-rep
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 ) > gpurun_out/t_t.log 2>&1; tail -9 gpurun_out/t_t.log
