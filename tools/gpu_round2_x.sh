#!/bin/bash
# Round 2, visit X (final build: batched key gathers in the shade stage on top of visit W): full GPU suite, benches, launch list
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { echo SMOKE FAILED; tail -5 gpurun_out/smoke.log; exit 1; }
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=3 ) > gpurun_out/t_x.log 2>&1; tail -4 gpurun_out/t_x.log
timeout 600 python bench.py > gpurun_out/bench_c2_x.json 2> gpurun_out/bench_c2_x.err
timeout 600 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4_x.json 2> gpurun_out/bench_c4_x.err
timeout 600 python bench.py --spp 8 --no-cpu-baseline > gpurun_out/bench_c2_8spp_x.json 2> gpurun_out/bench_c2_8spp_x.err
timeout 600 python bench.py --spp 1 --no-cpu-baseline > gpurun_out/bench_c2_1spp_x.json 2> gpurun_out/bench_c2_1spp_x.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_x.json 2> gpurun_out/bench_ref_x.err
python - <<'PY'
import json
for f in ("bench_c2_x","bench_c4_x","bench_c2_8spp_x","bench_c2_1spp_x","bench_ref_x"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); r=d.get("roofline",{})
        print(f, "%.2f Msamples/s e2e %.2f"%(d["value"], d["e2e"]["value"]), r.get("stage_ms_rank0"), d.get("framebuffer_sha256","")[:12], "cpu", d.get("cpu_baseline",{}).get("value"), "launches", d.get("gpu_launches"), "frac", r.get("frac"), r.get("dram_frac"), r.get("issue_frac"))
    except Exception as e: print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_x.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 2 -f -o gpurun_out/full_shade_c2_x python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 16 > gpurun_out/s1.log 2>&1
ls -la gpurun_out/launches_x.csv gpurun_out/full_shade_c2_x.ncu-rep
