#!/usr/bin/env python
"""Where does MCMCSampler.run spend host time beside the kernel?  Lorenz RW 4096 chains x 128 steps per call, the bench's
e2e leg, through `run` (torch staging) and `run_host` (C entry point), with a cProfile of `run`.
usage: python tools/e2e_probe.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ip_mcmc_b200 as M  # noqa: E402

wl = dict(bench.WORKLOADS["lorenz_rw"])
pot, proposer, accepter, u0 = bench.build_problem(M, wl, "fused")
s = M.MCMCSampler(proposer, accepter, np.random.default_rng(1))
B, S = wl["chains"], wl["mcmc_steps"]
u_host = np.broadcast_to(u0, (B, 3)).copy()
out_host = torch.empty((B, S, 3), dtype=torch.float64).pin_memory()
out_np = np.empty((B, S, 3))


def t(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


run = lambda: s.run(u_host, S, 0, 1, n_chains=B, out=out_host)
run_dev = lambda: s.run(u_host, S, 0, 1, n_chains=B, return_device=True)
host = lambda: s.run_host(u_host, S, 0, 1, n_chains=B, out=out_np)
print("run      %.2f ms per call" % t(run))
print("run(dev) %.2f ms per call" % t(run_dev))
print("run_host %.2f ms per call" % t(host))
print("run      %.2f ms per call" % t(run))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
run_dev()
e1.record()
torch.cuda.synchronize()
print("device time of one run(dev) call: %.2f ms" % e0.elapsed_time(e1))
pr = cProfile.Profile()
pr.enable()
run()
run()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
