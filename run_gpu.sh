timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -4 gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python -c "
import json; j=json.load(open('gpurun_out/bench_full.json')); r=j['roofline']; print('c2', j['value'], j['e2e']['value'], r['frac'], r['traffic'], r['per_launch'], r['stage_ms_rank0'], j['cpu_baseline']['value'])"
timeout 600 python bench.py --no-cpu-baseline --option overlap_shadow=0 > gpurun_out/bench_noovl.json 2> gpurun_out/bench_noovl.err; python -c "
import json; j=json.load(open('gpurun_out/bench_noovl.json')); r=j['roofline']; print('c2 no overlap', j['value'], j['e2e']['value'], r['frac'], r['stage_ms_rank0'])"
timeout 900 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; j=json.load(open('gpurun_out/bench_c4.json')); print('c4', j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['stage_ms_rank0'])"
