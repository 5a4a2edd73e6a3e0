B="timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
$B > gpurun_out/sweep_base2.json 2> gpurun_out/sweep_base2.err
for f in base2; do python -c "
import sys, json
j = json.loads(open('gpurun_out/sweep_$f.json').read()); print('$f', round(j['value'],1), round(j['roofline']['frac'],3), j['roofline']['stage_ms_rank0'])"; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 3 -o gpurun_out/prof_shade_r1a -f python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
