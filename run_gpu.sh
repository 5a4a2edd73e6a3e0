timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -5 gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace_persistent -c 40 --csv \
   --log-file gpurun_out/trace_dram.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu4.log 2>&1
tail -2 gpurun_out/trace_dram.csv
