timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -4 gpurun_out/t.log
timeout 300 python bench.py --steps 2 --warmup 2 --spp 16 --no-cpu-baseline > gpurun_out/c2.json 2> gpurun_out/c2.err; tail -2 gpurun_out/c2.err
timeout 300 python bench.py --steps 2 --warmup 2 --spp 16 --no-cpu-baseline --scene c4 > gpurun_out/c4.json 2> gpurun_out/c4.err; tail -2 gpurun_out/c4.err
