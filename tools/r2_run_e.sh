#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_burgers.py tests/test_gpu_chains.py -m gpu -q -x > gpurun_out/r2_gputest7.log 2>&1
echo "== new" > gpurun_out/r2_team_probe.txt
timeout 600 python tools/team_probe.py >> gpurun_out/r2_team_probe.txt 2>&1
echo "== old (nomono variant, round-1 team code)" >> gpurun_out/r2_team_probe.txt
IPMCMC_LIB=$PWD/gpurun_variants/libipmcmc_nomono.so timeout 600 python tools/team_probe.py >> gpurun_out/r2_team_probe.txt 2>&1
