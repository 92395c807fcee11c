#!/bin/bash
# the default bench under torchrun on 2 GPUs (final tree)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err
python tools/show_bench.py gpurun_out/r2c_bench_2gpu.json || tail -20 gpurun_out/r2c_bench_2gpu.err
