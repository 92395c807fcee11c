#!/usr/bin/env python
"""Evidence for the monotone-state CFL shortcut (burgers.cuh, time_loop_mono): G, Phi, the end state and the FV
step count of FUSED solves must be BIT-IDENTICAL with and without it (the end cells carry max|u| at every step).
    python tools/mono_check.py dump out.npz         (run once per library; IPMCMC_LIB selects the variant)
    python tools/mono_check.py compare a.npz b.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "dump":
    import ip_mcmc_b200 as M
    out = {}
    rng = np.random.default_rng(5)
    for N, n in ((64, 4096), (100, 2048), (256, 8192), (1024, 2048), (2048, 512)):
        f = M.BurgersFVM(N=N, numerics="fused")
        u = 0.25 * rng.standard_normal((n, 3))               # prior draws: shocks, rarefactions, sign changes,
        u[: n // 4, 2] = rng.uniform(-0.55, 1.55, n // 4)    # jumps at / outside both boundaries (blow-ups capped)
        r = f.batch(u, want_state=True)
        for k in ("G", "phi", "state", "work"):
            out["N%d_%s" % (N, k)] = r[k].cpu().numpy()
    np.savez(sys.argv[2], **out)
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    worst = 0
    for k in a.files:
        same = np.array_equal(a[k], b[k], equal_nan=True)
        n_diff = int(np.sum(~((a[k] == b[k]) | (np.isnan(a[k].astype(float)) & np.isnan(b[k].astype(float))))))
        print("%-14s %-14s bit-identical: %s (%d differing entries of %d)" % (k, a[k].shape, same, n_diff, a[k].size))
        worst += n_diff
    print("TOTAL differing entries:", worst)
