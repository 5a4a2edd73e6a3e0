timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -5 gpurun_out/t.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; j=json.load(open('gpurun_out/bench_c2.json')); print('c2 aov on', j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['stage_ms_rank0'])"
timeout 600 python bench.py --no-cpu-baseline --option aov_buffers=0 > gpurun_out/bench_c2_noaov.json 2> gpurun_out/bench_c2.err; python -c "
import json; j=json.load(open('gpurun_out/bench_c2_noaov.json')); print('c2 aov off', j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['stage_ms_rank0'])"
