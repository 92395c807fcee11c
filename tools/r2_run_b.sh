#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/dbg_lorenz.py > gpurun_out/dbg_lorenz.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_lorenz.py tests/test_gpu_statistics.py tests/test_gpu_chains.py -m gpu -q -s -x > gpurun_out/r2_gputest4.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
timeout 300 python tools/cold_start.py 256 1024 2000 200 > gpurun_out/r2_cold_start_cap8N.txt 2>&1
