#!/usr/bin/env python
"""A/B of Lorenz kernel variants built by tools/build_lorenz_variants.py (run on the GPU box):
    python tools/ab_lorenz.py [tag ...]      (default: the product library, then every gpurun_variants/libipmcmc_*.so)
Per variant (own process, IPMCMC_LIB): identical forward solves at 1 / 1.39 / 2 / 4 warps per SM sub-partition (cycles per
RK45 attempt) and the bench workload (4096 chains, RW, T = 20, 2 solves per step) in chain-steps/s."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import ip_mcmc_b200 as M
    g = np.load(os.path.join(ROOT, "tests", "golden", "lorenz_problem_K6_J4.npz"))
    out = {}
    f = M.Lorenz96Moments(6, 4, 20.0, 1.0, g["prior_means"], g["IC"], numerics="fused")
    for warps in (592, 820, 1184, 2368):
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
        out["w%d_cyc" % warps] = round(t * 1e-3 * 1.965e9 / att)
        out["w%d_tflops" % warps] = round(3444.0 * att * n / (t * 1e-3) / 1e12, 2)
        out["attempts"] = att
        out["G0"] = float(r["G"][0, 0].item())
    import bench
    wl = dict(bench.WORKLOADS["lorenz_rw"])
    pot, proposer, accepter, u0 = bench.build_problem(M, wl, "fused")
    sampler = M.MCMCSampler(proposer, accepter, np.random.default_rng(2))
    spec, pot, a = sampler._compile(10 ** 9, 0, 1, None)
    ch = M.ChainBatch(pot.problem(), u0, n_chains=wl["chains"])
    S = 8
    for _ in range(3):
        ch.run(spec, S)
    torch.cuda.synchronize()
    ts = []
    c0 = ch.counters[:, 2:4].sum().item()
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.run(spec, S); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    att = ch.counters[:, 2:4].sum().item() - c0
    tt = sum(ts) * 1e-3
    out["chain_steps_per_s"] = round(4 * S * wl["chains"] / tt)
    out["bench_tflops"] = round(3444.0 * att / tt / 1e12, 2)
    out["accept"] = round(ch.counters[:, 1].sum().item() / ch.counters[:, 0].sum().item(), 4)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
        sys.exit(0)
    tags = sys.argv[1:]
    libs = [("product", "")] if not tags else []
    for p in sorted(glob.glob(os.path.join(ROOT, "gpurun_variants", "libipmcmc_*.so"))):
        tag = os.path.basename(p)[len("libipmcmc_"):-3]
        if not tags or tag in tags:
            libs.append((tag, p))
    for tag, p in libs:
        env = dict(os.environ)
        if p:
            env["IPMCMC_LIB"] = p
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        line = r.stdout.strip().split("\n")[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED " + r.stderr[-800:]
        print("%-10s %s" % (tag, line), flush=True)
