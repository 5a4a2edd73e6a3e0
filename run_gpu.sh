timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -2 gpurun_out/t.log
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_$i.json 2> gpurun_out/bench_c2.err; python -c "
import json; j=json.load(open('gpurun_out/bench_c2_$i.json')); r=j['roofline']; print('c2', j['value'], j['e2e']['value'], r['frac'], r['stage_ms_rank0'])"; done
