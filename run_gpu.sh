timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -15 gpurun_out/t.log
timeout 300 python bench.py --steps 2 --warmup 2 --spp 16 --no-cpu-baseline --option bvh_builder=1 > gpurun_out/lbvh.json 2> gpurun_out/lbvh.err; tail -2 gpurun_out/lbvh.err
