#!/bin/bash
# Round-2 (third session), call c: compute-sanitizer over a tiny exercise of every kernel family (the tool is closed on
# this pool: it exits at once with a notice), then the default bench (e2e legs with 2 warm + >= 5 timed calls)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  T0=$(date +%s)
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2c_sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(( $(date +%s) - T0 )) s: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2c_sanitizer_$tool.log | tail -1)"
done
timeout 900 python bench.py > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err
python tools/show_bench.py gpurun_out/r2c_bench2.json
