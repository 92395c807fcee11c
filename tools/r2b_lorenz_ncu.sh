#!/bin/bash
# ncu --set full (source page) of the Lorenz queue kernel at the bench configuration, exported to CSV.
mkdir -p gpurun_out
TAG=${1:-r2b_lorenz}
L="python bench.py --workload lorenz_rw --steps 1 --warmup 3 --no-extra --no-cpu-baseline --mcmc-steps 8"
timeout 300 $L > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lorenz_chain_queue -s 3 -c 1 -o gpurun_out/${TAG} $L > gpurun_out/${TAG}_ncu.log 2>&1
bash tools/ncu_export.sh > /dev/null 2>&1
timeout 300 python tools/lorenz_occupancy.py > gpurun_out/${TAG}_occupancy.txt 2>&1
