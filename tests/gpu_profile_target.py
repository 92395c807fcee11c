"""Small profiling target (not a pytest file): python tests/gpu_profile_target.py <what> [N] [chains]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "forward"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
numerics = sys.argv[4] if len(sys.argv) > 4 else "fused"
TRUTH, PM = np.array([0.025, -0.025, -0.02]), np.array([1.5, 0.25, -0.5])
if what == "forward":
    f = M.BurgersFVM(N=N, numerics=numerics)
    u = 0.02 * np.random.default_rng(0).standard_normal((n, 3)) + (TRUTH - PM)
    pr = f._problem()
    for _ in range(4):
        r = pr.forward(u)
    torch.cuda.synchronize()
    print("ok", r["work"][:, 0].double().mean().item())
elif what == "lorenz":
    g = np.load(os.path.join(ROOT, "tests", "golden", "lorenz_problem_K6_J4.npz"))
    f = M.Lorenz96Moments(6, 4, 20.0, 1.0, g["prior_means"], g["IC"])
    u = np.tile(g["u0"], (n, 1)) + 0.01 * np.random.default_rng(0).standard_normal((n, 3))
    ic = np.tile(g["IC"], (n, 1))
    for _ in range(3):
        r = f.batch(u, ic)
    torch.cuda.synchronize()
    print("ok", r["work"].double().mean(0).tolist())
elif what == "chainstats":
    # per-chain work imbalance of one launch of the bench workload
    import bench
    wl = dict(bench.WORKLOADS["burgers_pcn_256"])
    wl["N"], wl["chains"] = N, n

    class A:
        numerics = "fused"
    pot, proposer, accepter, u0 = bench.build_problem(M, wl, numerics)
    sampler = M.MCMCSampler(proposer, accepter, np.random.default_rng(2))
    spec, pot, a = sampler._compile(10 ** 9, 0, 1, None)
    ch = M.ChainBatch(pot.problem(), u0, n_chains=n)
    for burn in (0, 200, 1500, 4000, 10000, 20000):
        if burn:
            ch.run(spec, burn - ch.step)
        c0 = ch.counters.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.run(spec, 50); e1.record(); torch.cuda.synchronize()
        d = (ch.counters - c0).double()
        w = d[:, 2]
        print("after %5d steps: launch %.2f ms, per-chain FV steps in 50 MCMC steps: mean %.0f max %.0f min %.0f  max/mean %.3f  acc %.3f  TFLOP/s %.2f"
              % (burn, e0.elapsed_time(e1), w.mean().item(), w.max().item(), w.min().item(), (w.max() / w.mean()).item(),
                 d[:, 1].sum().item() / d[:, 0].sum().item(), 29.0 * N * w.sum().item() / (e0.elapsed_time(e1) * 1e-3) / 1e12))
elif what == "occupancy":
    # identical work per chain; vary chains per SM sub-partition through the CTA shape
    f = M.BurgersFVM(N=N, numerics=numerics)
    pr = f._problem()
    for wpc, nch in ((4, 592), (8, 1184), (4, 1184), (4, 1776), (8, 2368), (4, 2368), (4, 4736), (1, 1024), (7, 1036), (4, 8192)):
        os.environ["IPMCMC_FWD_WPC"] = str(wpc)
        u = np.tile(TRUTH - PM, (nch, 1))
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = pr.forward(u); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = min(ts)
        nfv = r["work"][:, 0].double().mean().item()
        print("wpc %d chains %5d (%.2f warps/SMSP): %.3f ms  -> %.0f cycles per FV step per warp-slot, TFLOP/s %.2f"
              % (wpc, nch, nch / 592.0, t, t * 1e-3 * 1.965e9 / nfv, 29.0 * N * nfv * nch / (t * 1e-3) / 1e12))
elif what == "chainflat":
    # chain kernel with IDENTICAL work on every chain (zero injected noise, never accept):
    # separates per-step cost from load imbalance
    from ip_mcmc_b200 import _lib
    f = M.BurgersFVM(N=N, numerics=numerics)
    y = f.at_parameters(TRUTH)
    pot = M.EvolutionPotential(f, y, M.GaussianDistribution(np.zeros(5), 0.05 ** 2 * np.identity(5)))
    S = 20
    for nch in (592, 1024, 1184, 1776, 2048, 2368, 4736, 8192):
        spec = M.SamplerSpec(3, _lib.PROPOSE_PCN, _lib.ACCEPT_PCN, coef_u=1.0, coef_w=0.0)
        ch = M.ChainBatch(pot.problem(), TRUTH - PM, n_chains=nch)
        w = torch.zeros((nch, S, 3), dtype=torch.float64, device="cuda")
        U = torch.ones((nch, S), dtype=torch.float64, device="cuda")
        ts = []
        for _ in range(4):
            c0 = ch.counters[:, 2].sum().item()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ch.run(spec, S, inject_w=w, inject_u=U); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            nfv = (ch.counters[:, 2].sum().item() - c0) / nch
        t = min(ts)
        print("chains %5d (%.2f warps/SMSP): %.3f ms, %d FV steps/chain -> %.0f cycles per FV step per warp-slot, TFLOP/s %.2f"
              % (nch, nch / 592.0, t, nfv, t * 1e-3 * 1.965e9 / nfv, 29.0 * N * nfv * nch / (t * 1e-3) / 1e12), flush=True)
