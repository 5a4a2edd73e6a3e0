#!/bin/bash
# One GPU-box visit that reproduces the evidence under profiles/ (run as: gpurun -- 'bash tools/gpu_round.sh'):
# parity tests, smoke, headline bench (+ reference arm), ncu launch list, ncu --set full of the trace and shade kernels,
# per-launch DRAM traffic of the trace kernel.  Summaries: python profiles/ncu_keys.py <(ncu -i X.ncu-rep --page raw --csv).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1500 gpurun_out/bench_full.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -f -o gpurun_out/trace_full \
    python bench.py --steps 1 --warmup 0 --spp 4 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 1 -f -o gpurun_out/shade_full \
    python bench.py --steps 1 --warmup 0 --spp 4 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace_persistent -c 40 --csv \
    --log-file gpurun_out/trace_dram.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu4.log 2>&1
ls -la gpurun_out
