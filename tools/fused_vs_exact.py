#!/usr/bin/env python
"""Largest relative difference between the FUSED and the EXACT numerics of the Burgers forward model
(G, Phi, end state) over random prior draws on several grids (run on the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ip_mcmc_b200 as M

TRUTH = np.array([0.025, -0.025, -0.02])
worst = dict(G=0.0, phi=0.0, state=0.0)
for N in (64, 128, 256, 512, 1024):
    rng = np.random.default_rng(N)
    u = 0.25 * rng.standard_normal((256, 3))
    u[:, 2] = np.clip(u[:, 2], -0.4, 1.4)          # keep the jump inside the domain (no capped blow-up solves)
    res = {}
    for num in ("exact", "fused"):
        f = M.BurgersFVM(N=N, numerics=num)
        y = f.at_parameters(TRUTH)
        pot = M.EvolutionPotential(f, y, M.GaussianDistribution(np.zeros(5), 0.05 ** 2 * np.identity(5)))
        r = pot.problem().forward(u, want_state=True)
        res[num] = {k: r[k].cpu().numpy() for k in ("G", "phi", "state", "work")}
    same_steps = np.array_equal(res["exact"]["work"][:, 0], res["fused"]["work"][:, 0])
    out = {}
    for k in ("G", "phi", "state"):
        a, b = res["exact"][k], res["fused"][k]
        ok = np.isfinite(a) & np.isfinite(b)
        rel = np.abs(a - b)[ok] / np.maximum(np.abs(a)[ok], 1e-3)
        out[k] = rel.max()
        worst[k] = max(worst[k], out[k])
    print("N %5d: same FV step counts %s; max rel diff G %.2e  Phi %.2e  state %.2e" % (N, same_steps, out["G"], out["phi"], out["state"]))
print("worst:", {k: "%.2e" % v for k, v in worst.items()})
