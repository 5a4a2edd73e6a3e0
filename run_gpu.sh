timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -c 2 -o gpurun_out/prof_trace_r1g -f python bench.py --steps 1 --warmup 0 --spp 4 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 0 --spp 8 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
