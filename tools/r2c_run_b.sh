#!/bin/bash
# Round-2 (third session), call b: GPU tests incl. the exact stationary-law control, e2e host-time probe, full sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_gputest.log 2>&1
tail -3 gpurun_out/r2c_gputest.log
timeout 300 python tools/e2e_probe.py > gpurun_out/r2c_e2e_probe.txt 2>&1
head -8 gpurun_out/r2c_e2e_probe.txt
timeout 900 python sweep.py --out gpurun_out/r2c_sweep.json > gpurun_out/r2c_sweep.log 2>&1
tail -3 gpurun_out/r2c_sweep.log
