#!/usr/bin/env python
"""Find straggler chains of the bench workload: python tools/straggler_probe.py [chain_offset] [n_chains]
Prints, per launch of 50 steps, the launch time, the mean/max per-chain FV steps and the heaviest chains."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M  # noqa: E402
import bench  # noqa: E402

off = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
wl = dict(bench.WORKLOADS["burgers_pcn_256"])
pot, proposer, accepter, u0 = bench.build_problem(M, wl, "fused")
sampler = M.MCMCSampler(proposer, accepter, np.random.default_rng(2))
spec, pot, a = sampler._compile(10 ** 9, 0, 1, None)
start = bench.TRUTH - bench.PRIOR_MEAN if len(sys.argv) > 3 and sys.argv[3] == "posterior" else u0
ch = M.ChainBatch(pot.problem(), start, n_chains=n, chain_offset=off)
ch.run(spec, int(sys.argv[4]) if len(sys.argv) > 4 else 1500)
for _ in range(3):
    ch.run(spec, 50)
for it in range(10):
    c0 = ch.counters.clone()
    slog = torch.zeros((n, 50, 4), dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.run(spec, 50, steplog=slog); e1.record(); torch.cuda.synchronize()
    d = (ch.counters - c0)
    w = d[:, 2].double()
    top = torch.argsort(w, descending=True)[:3].tolist()
    print("launch %d: %.2f ms, FV steps/chain mean %.0f max %.0f; nonfinite %d; acc %.3f; u mean %s std %s" % (
        it, e0.elapsed_time(e1), w.mean().item(), w.max().item(), d[:, 4].sum().item(), d[:, 1].sum().item() / d[:, 0].sum().item(),
        np.round(ch.u.mean(0).cpu().numpy(), 3).tolist(), np.round(ch.u.std(0).cpu().numpy(), 3).tolist()))
    for c in top[:1]:
        s = slog[c].cpu().numpy()
        print("    chain %d (global %d): work %d, u = %s, per-step FV steps max %d (n>2000: %d), nonfinite %d"
              % (c, off + c, w[c].item(), np.round(ch.u[c].cpu().numpy(), 4).tolist(), s[:, 3].max(), (s[:, 3] > 2000).sum(), d[c, 4].item()))
