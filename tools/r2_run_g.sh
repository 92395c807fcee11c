#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/r2_sanitize_plain.log 2>&1 &&
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/r2_sanitize_memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/r2_sanitize_memcheck.log
timeout 1200 python sweep.py --quick --out gpurun_out/r2_sweep_quick.json > gpurun_out/r2_sweep_quick.log 2>&1
