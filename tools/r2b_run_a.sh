#!/bin/bash
# Lorenz GPU tests with the new FUSED controller, then A/B of the variants against the base library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lorenz.py tests/test_gpu_statistics.py -m gpu -q -x > gpurun_out/r2b_a_tests.log 2>&1
tail -5 gpurun_out/r2b_a_tests.log
timeout 600 python tools/ab_lorenz.py > gpurun_out/r2b_a_ab.txt 2>&1
cat gpurun_out/r2b_a_ab.txt
