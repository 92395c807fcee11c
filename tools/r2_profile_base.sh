#!/bin/bash
# Round-2 baseline captures on one B200: ncu --set full of the shipped Lorenz and Burgers queue kernels
# (source page included), and the cold-start probe.
set -x
mkdir -p gpurun_out
python tools/cold_start.py 256 1024 5000 200 > gpurun_out/r2_cold_start_base.txt 2>&1
L="python bench.py --workload lorenz_rw --steps 1 --warmup 3 --no-extra --no-cpu-baseline"
B="python bench.py --workload burgers_pcn_256 --steps 1 --warmup 3 --no-extra --no-cpu-baseline --burn-in 400"
$L > gpurun_out/r2_lorenz_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lorenz_chain_queue -s 3 -c 1 -o gpurun_out/r2_lorenz_base $L > gpurun_out/r2_lorenz_ncu.log 2>&1
$B > gpurun_out/r2_burgers_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:burgers_chain_queue -s 4 -c 1 -o gpurun_out/r2_burgers_base $B > gpurun_out/r2_burgers_ncu.log 2>&1
bash tools/ncu_export.sh
