#!/bin/bash
# final check of the round: smoke(), the GPU test suite, the reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_gputest.log 2>&1
tail -2 gpurun_out/r2c_gputest.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_reference_arm.err
cut -c1-260 gpurun_out/r2c_bench_reference_arm.json
