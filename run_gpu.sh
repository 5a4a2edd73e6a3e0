mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -8 gpurun_out/t.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_after_motion_aov.json 2> gpurun_out/bench_after_motion_aov.err; python -c "
import json; j=json.load(open('gpurun_out/bench_after_motion_aov.json')); print(j['value'], j['e2e']['value'], j['roofline']['frac'])"
