timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -2 gpurun_out/t.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python -c "
import json; j=json.load(open('gpurun_out/bench_full.json')); r=j['roofline']; print('c2', j['value'], j['e2e']['value'], r['frac'], r['traffic'], r['stage_ms_rank0'], j['cpu_baseline']['value'], j['config']['scene_setup_s'], j['gpu_launches'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -f -o gpurun_out/trace_full \
    python bench.py --steps 1 --warmup 0 --spp 4 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace_persistent -c 40 --csv \
   --log-file gpurun_out/trace_dram.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
RPTR_CUDA_LIB=variants/librptr_cuda_implied.so timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_implied.log 2>&1; tail -2 gpurun_out/t_implied.log
for v in base implied base implied; do RPTR_CUDA_LIB=variants/librptr_cuda_$v.so timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/sweep_$v.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/sweep_$v.json')); r=j['roofline']; print('var', '$v', round(j['value'],1), {k: round(x,1) for k,x in r['stage_ms_rank0'].items()})"; done
for v in base lbvh1 lbvh2; do RPTR_CUDA_LIB=variants/librptr_cuda_$v.so timeout 600 python bench.py --scene c4 --spp 16 --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/sweep_c4_$v.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/sweep_c4_$v.json')); r=j['roofline']; print('var c4', '$v', round(j['value'],1), r['per_ray']['closest'], {k: round(x,1) for k,x in r['stage_ms_rank0'].items()}, j['config']['scene_setup_s'])"; done
for v in base lbvh1; do RPTR_CUDA_LIB=variants/librptr_cuda_$v.so timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --option bvh_builder=1 > gpurun_out/sweep_c2lbvh_$v.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/sweep_c2lbvh_$v.json')); r=j['roofline']; print('var c2 lbvh', '$v', round(j['value'],1), r['per_ray']['closest'], j['config']['scene_setup_s'])"; done
