#!/usr/bin/env python
"""A/B measurement of kernel variants built by tools/build_variants.py (run on the GPU box):
    python tools/ab_variants.py [tag ...]        (default: every gpurun_variants/libipmcmc_*.so)
For each variant (own process, IPMCMC_LIB): flat-work chain launches (identical work per chain,
1024 and 8192 chains) and the bench workload (1024 chains x 256 cells pCN after 1500 burn-in steps),
plus a checksum of the chain states -- all variants are exact reformulations and must agree bit for bit."""
import glob
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(N=256):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import ip_mcmc_b200 as M
    from ip_mcmc_b200 import _lib
    import bench
    TRUTH, PM = bench.TRUTH, bench.PRIOR_MEAN
    out = {}
    f = M.BurgersFVM(N=N, numerics="fused")
    y = f.at_parameters(TRUTH)
    pot = M.EvolutionPotential(f, y, M.GaussianDistribution(np.zeros(5), 0.05 ** 2 * np.identity(5)))
    S = 20
    for nch in (1024, 8192):
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
        out["flat%d_tflops" % nch] = round(29.0 * N * nfv * nch / (t * 1e-3) / 1e12, 3)
        out["flat%d_cyc" % nch] = round(t * 1e-3 * 1.965e9 / nfv, 1)
    wl = dict(bench.WORKLOADS["burgers_pcn_256"])
    pot, proposer, accepter, u0 = bench.build_problem(M, wl, "fused")
    sampler = M.MCMCSampler(proposer, accepter, np.random.default_rng(2))
    spec, pot, a = sampler._compile(10 ** 9, 0, 1, None)
    ch = M.ChainBatch(pot.problem(), u0, n_chains=wl["chains"])
    ch.run(spec, 1500)
    for _ in range(3):
        ch.run(spec, 50)
    torch.cuda.synchronize()
    c0 = ch.counters[:, 2].sum().item()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.run(spec, 50); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    work = ch.counters[:, 2].sum().item() - c0
    out["bench_tflops"] = round(29.0 * N * work / (sum(ts) * 1e-3) / 1e12, 3)
    out["bench_ms"] = round(sum(ts) / len(ts), 3)
    out["checksum"] = hashlib.sha1(ch.u.cpu().numpy().tobytes() + ch.counters.cpu().numpy().tobytes()).hexdigest()[:12]
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("AB_CHILD"):
        child()
        sys.exit(0)
    tags = sys.argv[1:]
    libs = sorted(glob.glob(os.path.join(ROOT, "gpurun_variants", "libipmcmc_*.so")))
    for lib in libs:
        tag = os.path.basename(lib)[len("libipmcmc_"):-3]
        if tags and tag not in tags:
            continue
        env = dict(os.environ, IPMCMC_LIB=lib, AB_CHILD="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        res = [l for l in r.stdout.split("\n") if l.startswith("RESULT ")]
        print("%-8s %s" % (tag, res[0][7:] if res else "FAILED: " + r.stderr[-600:]), flush=True)
