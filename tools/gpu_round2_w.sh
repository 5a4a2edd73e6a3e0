#!/bin/bash
# Round 2, visit W (final: tail hand-over with resume): full GPU suite, benches, ncu counters + captures, launch lists, sanitizers on the tiny tour
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo SMOKE FAILED; tail -5 gpurun_out/smoke.log; exit 1; }
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/t_w.log 2>&1; tail -5 gpurun_out/t_w.log
timeout 600 python bench.py > gpurun_out/bench_c2_w.json 2> gpurun_out/bench_c2_w.err
timeout 600 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4_w.json 2> gpurun_out/bench_c4_w.err
timeout 600 python bench.py --spp 8 --no-cpu-baseline > gpurun_out/bench_c2_8spp_w.json 2> gpurun_out/bench_c2_8spp_w.err
timeout 600 python bench.py --spp 1 --no-cpu-baseline > gpurun_out/bench_c2_1spp_w.json 2> gpurun_out/bench_c2_1spp_w.err
python - <<'PY'
import json
for f in ("bench_c2_w","bench_c4_w","bench_c2_8spp_w","bench_c2_1spp_w"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "%.1f Msamples/s e2e %.1f"%(d["value"], d["e2e"]["value"]), r["stage_ms_rank0"], d["framebuffer_sha256"][:12], "cpu", d.get("cpu_baseline",{}).get("value"), "launches", d["gpu_launches"], "frac", r["frac"], r.get("dram_frac"), r.get("issue_frac"))
    except Exception as e: print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_elapsed.max
timeout 900 ncu --metrics $M --clock-control none -k regex:k_trace_persistent -c 60 --csv --log-file gpurun_out/trace_metrics_c2_w.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu_c2.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:k_trace_persistent -c 60 --csv --log-file gpurun_out/trace_metrics_c4_w.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --scene c4 --spp 16 > gpurun_out/b_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 6 -f -o gpurun_out/full_trace_c2_w \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/b_ncu_c2f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 6 -f -o gpurun_out/full_trace_c4_w \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --scene c4 --spp 8 > gpurun_out/b_ncu_c4f.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_w.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_8spp_w.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 8 > gpurun_out/b_l8.log 2>&1
for tool in racecheck memcheck synccheck initcheck; do
  ( time RPTR_CUDA_LIB=variants/librptr_cuda_t128.so timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize.py --tiny ) > gpurun_out/${tool}_tiny.log 2>&1; echo "$tool rc=$?"; grep -i "summary\|contexts ok" gpurun_out/${tool}_tiny.log | tail -2
done
ls -la gpurun_out/*_w.ncu-rep gpurun_out/*_w.csv
