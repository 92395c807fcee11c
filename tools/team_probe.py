#!/usr/bin/env python
"""Team solver (2048 / 4096 cells, 2 / 4 warps per chain): forward solves of posterior-region parameters.
python tools/team_probe.py [n_chains]   ->  ms, FV steps per solve, fp64 TFLOP/s (29 FLOP per cell-step)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
TRUTH, PM = np.array([0.025, -0.025, -0.02]), np.array([1.5, 0.25, -0.5])
peak = M.fp64_peak_tflops(3)
rng = np.random.default_rng(0)
for N in (2048, 4096):
    for num in ("fused", "exact"):
        f = M.BurgersFVM(N=N, numerics=num)
        for label, u in (("posterior (right state < 0)", 0.02 * rng.standard_normal((n, 3)) + (TRUTH - PM)),
                         ("prior mean (positive)", 0.05 * rng.standard_normal((n, 3)))):
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = f.batch(u); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = min(ts)
            nfv = r["work"][:, 0].double().sum().item()
            tf = 29.0 * N * nfv / (t * 1e-3) / 1e12
            print("N %d %s %-28s: %8.2f ms, %5.0f FV steps/solve, %.2f TFLOP/s = %.3f of %.1f" % (N, num, label, t, nfv / n, tf, tf / peak, peak), flush=True)
