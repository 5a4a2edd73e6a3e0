#!/bin/bash
# Round 2, visit A: parity tests (incl. the new full-frame ones), smoke, headline bench + reference arm, C4 bench,
# per-launch hardware counters of the trace kernel on C2 and C4, ncu --set full of the first closest-hit launches on C4.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/t.log 2>&1; tail -15 gpurun_out/t.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 2500 gpurun_out/bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; tail -c 900 gpurun_out/bench_ref.json
timeout 900 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 1500 gpurun_out/bench_c4.json
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_elapsed.max
timeout 900 ncu --metrics $M --clock-control none -k regex:k_trace_persistent -c 60 --csv --log-file gpurun_out/trace_metrics_c2.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu_c2.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:k_trace_persistent -c 60 --csv --log-file gpurun_out/trace_metrics_c4.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --scene c4 --spp 16 > gpurun_out/b_ncu_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 3 -f -o gpurun_out/trace_full_c4 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --scene c4 --spp 16 > gpurun_out/b_ncu_c4f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 3 -f -o gpurun_out/trace_full_c2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/b_ncu_c2f.log 2>&1
ls -la gpurun_out | tail -15
