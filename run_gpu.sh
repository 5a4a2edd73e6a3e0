mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -5 gpurun_out/t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_after_alpha_qmc.json 2> gpurun_out/bench_after_alpha_qmc.err; python -c "
import json; j=json.load(open('gpurun_out/bench_after_alpha_qmc.json')); print(j['value'], j['e2e']['value'], j['roofline']['frac'])"
