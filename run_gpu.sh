for v in base nopf base nopf; do RPTR_CUDA_LIB=variants/librptr_cuda_$v.so timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/sweep_$v.json 2> gpurun_out/sweep_$v.err || tail -2 gpurun_out/sweep_$v.err; python -c "
import json; j=json.load(open('gpurun_out/sweep_$v.json')); r=j['roofline']; print('var', '$v', round(j['value'],1), r['frac'], r['stage_ms_rank0'])"; done
RPTR_CUDA_LIB=variants/librptr_cuda_base.so timeout 600 python bench.py --scene c4 --spp 16 --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/sweep_c4pf.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/sweep_c4pf.json')); print('var c4 prefetch', j['value'], j['roofline']['stage_ms_rank0'])"
RPTR_CUDA_LIB=variants/librptr_cuda_nopf.so timeout 600 python bench.py --scene c4 --spp 16 --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/sweep_c4nopf.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/sweep_c4nopf.json')); print('var c4 no prefetch', j['value'], j['roofline']['stage_ms_rank0'])"
