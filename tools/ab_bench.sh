#!/bin/bash
# A/B of kernel variants (tools/build_variants.py) through bench.py itself: tools/ab_bench.sh [-1] tag ...
# (-1: only the 1024 x 256 workload)
run() { timeout 150 python bench.py --no-cpu-baseline --no-extra --steps 8 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('  ',round(d['value']),'frac',round(d['roofline']['frac'],4),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'acc',round(d['acceptance_rate'],4),'peak',round(d['roofline']['peak'],2),'mhz',d['clocks']['sm_mhz'],d['clocks']['reasons'])"; }
ONLY256=0; if [ "$1" = "-1" ]; then ONLY256=1; shift; fi
for t in "$@"; do
  export IPMCMC_LIB=$PWD/gpurun_variants/libipmcmc_$t.so
  echo "$t 256";  run
  [ $ONLY256 = 1 ] || { echo "$t 1024"; run --workload burgers_pcn_1024 --steps 3; }
done
