#!/bin/bash
# Round-2 (third session) verification on one B200: GPU tests, then the default bench.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_gputest.log 2>&1
tail -3 gpurun_out/r2c_gputest.log
echo "tests: $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench: $(( $(date +%s) - T0 )) s"
python tools/show_bench.py gpurun_out/r2c_bench.json
