#!/bin/bash
# GPU tests (all), then the default bench, then the cold-start probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2_gputest3.log
python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
python tools/cold_start.py 256 1024 2000 200 > gpurun_out/r2_cold_start_cap8N.txt 2>&1
