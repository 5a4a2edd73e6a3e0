timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
$B > gpurun_out/sweep_diet.json 2> gpurun_out/sweep_diet.err
for f in diet; do python -c "
import sys, json
j = json.loads(open('gpurun_out/sweep_$f.json').read()); print('$f', round(j['value'],1), round(j['roofline']['frac'],3), j['roofline']['stage_ms_rank0'])"; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 2 -o gpurun_out/prof_trace_r1i -f python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
