timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -2 gpurun_out/t.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; j=json.load(open('gpurun_out/bench_c2.json')); r=j['roofline']; print('c2', j['value'], j['e2e']['value'], r['frac'])"
