#!/usr/bin/env python
"""sweep.py -- BASELINE.json configs[4]: chains x grid x pCN-beta sweep of the Burgers inverse
problem on one B200, reporting acceptance rate, chain-steps/s, ESS/s and the fp64 roofline
fraction for every point.

    python sweep.py [--quick] [--out profiles/sweep_rNN.json]

Every point runs the reference's problem (burgers_mcmc.py:22-83) at N cells with B chains started
at u_0 = 0, an untimed burn-in, then timed launches whose recorded states feed the ESS estimator
(stats.ess: n / (1 + 2 sum rho_k), Geyer cut-off, rho from MCMCSampler.autocorr; min over
parameters, mean over traced chains).  Device timing with CUDA events.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M  # noqa: E402
from ip_mcmc_b200.engine import ChainBatch, F64  # noqa: E402

TRUTH = np.array([0.025, -0.025, -0.02])
PRIOR_MEAN = np.array([1.5, 0.25, -0.5])


def run_point(N, B, beta, numerics, budget_fv_steps, trace_chains=32):
    f = M.BurgersFVM(N=N, numerics=numerics)
    y = f.at_parameters(TRUTH)
    prior = M.GaussianDistribution(PRIOR_MEAN, 0.25 ** 2 * np.identity(3))
    pot = M.EvolutionPotential(f, y, M.GaussianDistribution(np.zeros(5), 0.05 ** 2 * np.identity(5)))
    sampler = M.MCMCSampler(M.ConstSteppCNProposer(beta, prior), M.CountedAccepter(M.pCNAccepter(pot)),
                            np.random.default_rng(2))
    spec, pot, _ = sampler._compile(10 ** 9, 0, 1, None)
    chains = ChainBatch(pot.problem(), np.zeros(3), n_chains=B)
    # size the run so that every point costs about the same number of cell updates
    per_step = 1.2 * N * N * B                       # ~ FV steps (1.2 N) x cells per chain-step
    S = int(max(8, min(400, budget_fv_steps / per_step)))
    burn = int(max(50, min(2000, 4 * S)))
    chains.run(spec, burn)
    trace = torch.empty((B, S, 3), dtype=F64, device="cuda")
    c0 = chains.counters.sum(0)
    kept = []
    K = 3
    ms = 0.0
    for _ in range(K):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chains.run(spec, S, trace=trace)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
        kept.append(trace[:trace_chains].clone())
    dc = (chains.counters.sum(0) - c0).double().cpu().numpy()
    tr = torch.cat(kept, dim=1).cpu().numpy()
    ess_tot, _ = M.stats.ess_multichain(tr)
    ess_chain = ess_tot / tr.shape[0]
    flops = 29.0 * N * dc[2]
    return dict(N=N, chains=B, beta=beta, mcmc_steps_timed=K * S, burn_in=burn,
                chain_steps_per_sec=B * K * S / (ms * 1e-3), acceptance=dc[1] / dc[0],
                mean_fv_steps_per_solve=dc[2] / dc[3], ess_per_chain=ess_chain,
                ess_per_sec=ess_chain * B / (ms * 1e-3), tflops=flops / (ms * 1e-3) / 1e12, ms=ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--numerics", default="fused")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    peak = M.fp64_peak_tflops(5)
    if args.quick:
        grid = [(N, B, 0.25) for N in (64, 256, 1024) for B in (1024, 16384)]
        betas = [(256, 4096, b) for b in (0.05, 0.5)]
        budget = 2e11
    else:
        grid = [(N, B, 0.25) for N in (64, 128, 256, 512, 1024, 2048, 4096) for B in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20)
                if N * B <= (1 << 26)]
        betas = [(256, 4096, b) for b in (0.01, 0.05, 0.15, 0.5)]
        budget = 1.5e12
    points = []
    t0 = time.time()
    for (N, B, beta) in grid + betas:
        r = run_point(N, B, beta, args.numerics, budget)
        r["roofline_frac"] = r["tflops"] / peak
        points.append(r)
        print("N %5d chains %8d beta %.2f: %10.0f chain-steps/s  acc %.3f  ESS/s %9.0f  %.2f TFLOP/s (%.0f%% of fp64 peak)  [%.1fs]"
              % (N, B, beta, r["chain_steps_per_sec"], r["acceptance"], r["ess_per_sec"], r["tflops"],
                 100 * r["roofline_frac"], time.time() - t0), flush=True)
    out = dict(fp64_peak_tflops=peak, numerics=args.numerics, gpu=torch.cuda.get_device_name(0), points=points,
               note="N <= 1024: one warp per chain; N = 2048 / 4096: 2 / 4 warps per chain (burgers_team.cuh)")
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
