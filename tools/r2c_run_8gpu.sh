#!/bin/bash
# the default bench under torchrun on 8 GPUs (final tree)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2c_bench_8gpu.json 2> gpurun_out/r2c_bench_8gpu.err
python tools/show_bench.py gpurun_out/r2c_bench_8gpu.json || tail -20 gpurun_out/r2c_bench_8gpu.err
