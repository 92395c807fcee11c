#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gputest5.log 2>&1
L="python bench.py --workload lorenz_rw --steps 1 --warmup 3 --no-extra --no-cpu-baseline --mcmc-steps 8"
timeout 300 $L > gpurun_out/r2_lorenz_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lorenz_chain_queue -s 3 -c 1 -o gpurun_out/r2_lorenz_v2 $L > gpurun_out/r2_lorenz_ncu.log 2>&1
bash tools/ncu_export.sh > /dev/null 2>&1
