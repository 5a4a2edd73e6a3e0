#!/bin/bash
# Round 2, visit E: all parity tests (textures, reference-path images), racecheck on the tiny tour, benches, per-launch counters
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/t_e.log 2>&1; tail -12 gpurun_out/t_e.log
timeout 900 python bench.py > gpurun_out/bench_c2_e.json 2> gpurun_out/bench_c2_e.err; python tools/sweep_report.py >/dev/null 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2_e.json",):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "%.1f Msamples/s e2e %.1f"%(d["value"], d["e2e"]["value"]), r["stage_ms_rank0"], "cpu", d.get("cpu_baseline",{}).get("value"), d["framebuffer_sha256"][:12])
PY
timeout 900 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4_e.json 2> gpurun_out/bench_c4_e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c4_e.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("c4 %.1f Msamples/s e2e %.1f"%(d["value"], d["e2e"]["value"]), r["stage_ms_rank0"], "setup", d["config"]["scene_setup_s"])
PY
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_elapsed.max
timeout 900 ncu --metrics $M --clock-control none -k regex:k_trace_persistent -c 60 --csv --log-file gpurun_out/trace_metrics_c2_e.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu_c2.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:k_trace_persistent -c 60 --csv --log-file gpurun_out/trace_metrics_c4_e.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --scene c4 --spp 16 > gpurun_out/b_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 3 -f -o gpurun_out/trace_full_c2_e \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/b_ncu_c2f.log 2>&1
( time RPTR_CUDA_LIB=variants/librptr_cuda_t128.so timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitize.py --tiny ) > gpurun_out/racecheck_tiny.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/racecheck_tiny.log
( time RPTR_CUDA_LIB=variants/librptr_cuda_t128.so timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py --tiny ) > gpurun_out/memcheck_tiny.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_tiny.log
