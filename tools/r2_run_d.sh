#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_burgers.py tests/test_gpu_chains.py tests/test_gpu_statistics.py -m gpu -q -x > gpurun_out/r2_gputest6.log 2>&1
bash tools/ab_bench.sh nomono > gpurun_out/r2_ab_mono.txt 2>&1
unset IPMCMC_LIB
echo "main (mono)" >> gpurun_out/r2_ab_mono.txt
for w in burgers_pcn_256 burgers_pcn_1024; do
timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 8 --warmup 3 --workload $w 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('  ','$w',round(d['value']),'frac',round(d['roofline']['frac'],4),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'acc',round(d['acceptance_rate'],4),'peak',round(d['roofline']['peak'],2))" >> gpurun_out/r2_ab_mono.txt
done
