run() { timeout 150 python bench.py --no-cpu-baseline --no-extra --steps 8 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('  ',round(d['value']),'frac',round(d['roofline']['frac'],4),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'acc',round(d['acceptance_rate'],4))"; }
echo "v2 256 s50"; run
echo "v2 256 s100"; run --mcmc-steps 100
echo "v2 256 s200"; run --mcmc-steps 200 --steps 4
echo "v2 256 s50 chunk2"; IPMCMC_SCHED_CHUNK=2 run
echo "v2 256 s200 chunk2"; IPMCMC_SCHED_CHUNK=2 run --mcmc-steps 200 --steps 4
echo "v2 256 s200 chunk5"; IPMCMC_SCHED_CHUNK=5 run --mcmc-steps 200 --steps 4
echo "v2 1024"; run --workload burgers_pcn_1024 --steps 3
