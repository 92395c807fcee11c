#!/bin/bash
# full GPU test suite, Lorenz bench line, ncu of the Lorenz queue kernel (source page) + occupancy probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_b_tests.log 2>&1
tail -3 gpurun_out/r2b_b_tests.log
timeout 600 python bench.py --workload lorenz_rw --no-extra --no-cpu-baseline > gpurun_out/r2b_b_bench_lorenz.json 2> gpurun_out/r2b_b_bench_lorenz.err
cut -c1-400 gpurun_out/r2b_b_bench_lorenz.json
bash tools/r2b_lorenz_ncu.sh r2b_lorenz
cat gpurun_out/r2b_lorenz_occupancy.txt
