#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2b_bench_8gpu.json 2> gpurun_out/r2b_bench_8gpu.err
python tools/show_bench.py gpurun_out/r2b_bench_8gpu.json
