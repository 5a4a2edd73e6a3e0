nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --spp 16 > gpurun_out/mg2.json 2> gpurun_out/mg2.err; tail -5 gpurun_out/mg2.err; cat gpurun_out/mg2.json | cut -c1-900
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 --cpu-seconds 6 > gpurun_out/ref.json 2> gpurun_out/ref.err; tail -2 gpurun_out/ref.err; cat gpurun_out/ref.json | cut -c1-600
