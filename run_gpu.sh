timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
$B > gpurun_out/sweep_base.json 2> gpurun_out/sweep_base.err
for f in base; do python -c "
import sys, json
j = json.loads(open('gpurun_out/sweep_$f.json').read()); print('$f', round(j['value'],1), round(j['roofline']['frac'],3), j['roofline']['stage_ms_rank0'])"; done
