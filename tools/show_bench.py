#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line: python tools/show_bench.py <file>"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print("main   %.3f M chain-steps/s, %.1f %% of the DFMA peak (%.1f %% nominal), e2e %.3f M, c-abi %.3f M, acceptance %.3f, clocks %s"
      % (d["value"] / 1e6, 100 * d["roofline"]["frac"], 100 * d["roofline"]["frac_of_nominal"], d["e2e"]["value"] / 1e6,
         d["e2e_c_abi"]["value"] / 1e6, d["acceptance_rate"], d["clocks"]))
if d.get("short_launch"):
    print("short  %.3f M, %.1f %% (%d steps per launch)" % (d["short_launch"]["chain_steps_per_sec"] / 1e6,
          100 * d["short_launch"]["roofline_frac"], d["short_launch"]["mcmc_steps_per_launch"]))
if d.get("cold_start"):
    c = d["cold_start"]
    print("cold   %.3f M, %.1f %%, slowest/fastest %.3f, capped solves %d" % (c["chain_steps_per_sec"] / 1e6, 100 * c["roofline_frac"],
          c["slowest_over_fastest"], c["capped_solves_per_rank"]))
for k, v in (d.get("extra_workloads") or {}).items():
    e = " e2e %.3f M c-abi %.3f M" % (v["e2e"]["value"] / 1e6, v["e2e_c_abi"]["value"] / 1e6) if "e2e" in v else ""
    cpu = " cpu %.1f / C %.1f" % (v["cpu_baseline"]["value"], v["cpu_baseline_c"].get("value", float("nan"))) if "cpu_baseline" in v else ""
    print("%-22s %.3f M, %.1f %%, pool %.3f ms, reduce %.3f ms%s%s" % (k, v["chain_steps_per_sec"] / 1e6, 100 * v["roofline_frac"],
          v["pool_moments_ms"], v["allreduce_ms"], e, cpu))
if "cpu_baseline" in d:
    print("cpu    %.1f chain-steps/s (%s, %d cores); C port %.1f" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"],
          d["cpu_baseline"]["cores"], d.get("cpu_baseline_c", {}).get("value", float("nan"))))
