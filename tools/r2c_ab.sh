#!/bin/bash
# A/B through bench.py: the product library against variants in gpurun_variants/ (tools/build_variants.py).
# usage: tools/r2c_ab.sh tag ...   ("product" = ip_mcmc_b200/libipmcmc.so)
run() { timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 8 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('  ',round(d['value']),'frac',round(d['roofline']['frac'],4),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'acc',round(d['acceptance_rate'],4),'peak',round(d['roofline']['peak'],2),'mhz',d['clocks']['sm_mhz'],d['clocks']['reasons'])"; }
for t in "$@"; do
  if [ "$t" = product ]; then unset IPMCMC_LIB; else export IPMCMC_LIB=$PWD/gpurun_variants/libipmcmc_$t.so; fi
  echo "$t 256";  run
  echo "$t 1024"; run --workload burgers_pcn_1024 --steps 3
done
