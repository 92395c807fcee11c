#!/usr/bin/env python
"""Lorenz forward kernel (one T = 20 solve per chain, identical parameters) at 1, 1.39, 2 warps per SM
sub-partition: cycles per RK45 attempt per warp and fp64 pipe utilisation.  python tools/lorenz_occupancy.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M
g = np.load(os.path.join(ROOT, "tests", "golden", "lorenz_problem_K6_J4.npz"))
for numerics in ("fused", "exact"):
    f = M.Lorenz96Moments(6, 4, 20.0, 1.0, g["prior_means"], g["IC"], numerics=numerics)
    for warps in (148, 296, 592, 820, 1184, 2368):
        n = warps * 5
        u = np.tile(g["u0"], (n, 1))
        ic = np.tile(g["IC"], (n, 1))
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = f.batch(u, ic); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = min(ts)
        att = r["work"].double().sum(1).mean().item()
        cyc = t * 1e-3 * 1.965e9 / att
        tf = 3444.0 * att * n / (t * 1e-3) / 1e12
        print("%s warps %5d (%.2f/SMSP): %.2f ms, %d attempts -> %.0f cycles per attempt per warp slot, %.2f TFLOP/s"
              % (numerics, warps, warps / 592.0, t, att, cyc, tf), flush=True)
