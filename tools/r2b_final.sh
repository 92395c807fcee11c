#!/bin/bash
# Round-2 (second half) final captures on one B200: GPU tests, the default bench, the reference arm, the ncu launch list
# of the bench command, ncu --set full of the two dominant kernels (source page), exported to small CSVs.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_gputest_final.log 2>&1
tail -2 gpurun_out/r2b_gputest_final.log
timeout 1200 python bench.py > gpurun_out/r2b_bench_final.json 2> gpurun_out/r2b_bench_final.err
cut -c1-300 gpurun_out/r2b_bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_bench_reference_arm.err
cut -c1-300 gpurun_out/r2b_bench_reference_arm.json
B="python bench.py --no-cpu-baseline --no-extra --steps 2 --warmup 3 --burn-in 400"
L="python bench.py --workload lorenz_rw --steps 1 --warmup 3 --no-extra --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2b_plain_b.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_bench.csv $B > gpurun_out/r2b_ncu_launches.log 2>&1
timeout 300 $B > gpurun_out/r2b_plain_b2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:burgers_chain_queue -s 4 -c 1 -o gpurun_out/r2b_burgers_final $B > gpurun_out/r2b_burgers_ncu.log 2>&1
timeout 300 $L > gpurun_out/r2b_plain_l.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lorenz_chain_queue -s 3 -c 1 -o gpurun_out/r2b_lorenz_final $L > gpurun_out/r2b_lorenz_ncu.log 2>&1
bash tools/ncu_export.sh > /dev/null 2>&1
ls gpurun_out
