timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -2 gpurun_out/t.log
timeout 900 python bench.py --scene c4 --spp 16 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; j=json.load(open('gpurun_out/bench_c4.json')); print('c4', j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['stage_ms_rank0'], j['config']['scene_setup_s'])"
timeout 600 python bench.py --no-cpu-baseline --option bvh_builder=1 > gpurun_out/bench_c2_lbvh.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/bench_c2_lbvh.json')); print('c2 device lbvh', j['value'], j['e2e']['value'], j['roofline']['frac'], j['config']['scene_setup_s'])"
