timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -2 gpurun_out/t.log
timeout 300 python bench.py --steps 2 --warmup 2 --spp 16 --no-cpu-baseline > gpurun_out/ab_0.json 2> gpurun_out/ab_0.err; tail -2 gpurun_out/ab_0.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 2 -o gpurun_out/prof_trace_r1f -f python bench.py --steps 1 --warmup 0 --spp 4 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
