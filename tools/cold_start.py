#!/usr/bin/env python
"""Cold start of the bench workload from the reference's u_0 = 0 (burgers_mcmc.py:129-134):
python tools/cold_start.py [N] [chains] [total_steps] [steps_per_launch]
Per launch: time, fp64 TFLOP/s, mean/max FV steps per chain, non-finite (capped) solves, acceptance."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M  # noqa: E402
import bench  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
total = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
S = int(sys.argv[4]) if len(sys.argv) > 4 else 200
wl = dict(bench.WORKLOADS["burgers_pcn_256"])
wl["N"], wl["chains"] = N, n
pot, proposer, accepter, u0 = bench.build_problem(M, wl, "fused")
sampler = M.MCMCSampler(proposer, accepter, np.random.default_rng(2))
spec, pot, a = sampler._compile(10 ** 9, 0, 1, None)
peak = M.fp64_peak_tflops(3)
ch = M.ChainBatch(pot.problem(), u0, n_chains=n)
tot_ms, tot_fl = 0.0, 0.0
done = 0
while done < total:
    c0 = ch.counters.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.run(spec, S); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d = (ch.counters - c0).double()
    w = d[:, 2]
    fl = 29.0 * N * w.sum().item()
    tot_ms += ms
    tot_fl += fl
    done += S
    print("steps %5d: %.2f ms  %.2f TFLOP/s (%.3f)  FV/chain mean %.0f max %.0f (max/mean %.2f)  nonfinite %d  acc %.3f"
          % (done, ms, fl / ms / 1e9, fl / ms / 1e9 / peak, w.mean().item(), w.max().item(), (w.max() / w.mean()).item(),
             d[:, 4].sum().item(), d[:, 1].sum().item() / d[:, 0].sum().item()), flush=True)
print("TOTAL %d steps x %d chains: %.1f ms -> %.3f M chain-steps/s, %.2f TFLOP/s = %.3f of DFMA peak %.1f"
      % (total, n, tot_ms, n * total / tot_ms / 1e3, tot_fl / tot_ms / 1e9, tot_fl / tot_ms / 1e9 / peak, peak))
