#!/usr/bin/env python
"""EXACT Burgers forward solves (identical parameters per batch) with and without the positive-monotone shortcut:
ms per batch and fp64 TFLOP/s (29 FLOP per cell-step).  python tools/exact_mono_probe.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M
PM = np.array([1.5, 0.25, -0.5])
cases = {"positive shock (2.525 | 0.225)": np.array([1.525, 0.225, -0.52]) - PM,
         "sign-changing shock (2.525 | -0.275)": np.array([1.525, -0.275, -0.52]) - PM}
for N in (256, 1024):
    for name, u1 in cases.items():
        for n in (1024, 8192):
            u = np.tile(u1, (n, 1))
            for sc in (True, False):
                f = M.BurgersFVM(N=N, numerics="exact", monotone_shortcut=sc)
                ts = []
                for _ in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); r = f.batch(u); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                t = min(ts)
                nfv = r["work"][:, 0].double().mean().item()
                print("N=%4d %-38s chains %5d shortcut %-5s %.2f ms, %d FV steps, %.2f TFLOP/s"
                      % (N, name, n, sc, t, nfv, 29.0 * N * nfv * n / (t * 1e-3) / 1e12), flush=True)
