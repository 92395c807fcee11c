"""Exploratory GPU script (not a pytest file): prints device-vs-golden/oracle differences.
Run on the GPU box:  python tests/gpu_explore.py"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ip_mcmc_b200 as M  # noqa: E402
from ip_mcmc_b200 import _lib  # noqa: E402
from oracle import burgers_np as B, lorenz_np as L, mcmc_np as O, philox_np as P  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def gold(n):
    return np.load(os.path.join(GOLD, n))


def section(f):
    print("=" * 20, f.__name__, flush=True)
    t0 = time.time()
    try:
        f()
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()
    print(f"[{f.__name__}] {time.time() - t0:.2f}s", flush=True)


def s_rng():
    lib = _lib.load()
    out = torch.empty((3, 50, 4), dtype=torch.float64, device="cuda")
    _lib.check(lib.ipmcmc_rng_probe(2 + (5 << 32), 7, 1000, 3, 50, 3, out.data_ptr(), None))
    o = out.cpu().numpy()
    ref = np.empty_like(o)
    for c in range(3):
        z, U = P.chain_noise(2 + (5 << 32), 7 + c, 1000, 50, 3)
        ref[c, :, :3] = z
        ref[c, :, 3] = U
    print("uniform exact:", np.array_equal(o[..., 3], ref[..., 3]), "normal max abs diff:", np.abs(o[..., :3] - ref[..., :3]).max())


def s_burgers_forward():
    for numerics in ("exact", "fused"):
        for N in (32, 64, 100, 128, 200, 256, 1024):
            g = gold(f"burgers_forward_N{N}.npz")
            f = M.BurgersFVM(N=N, numerics=numerics)
            noise = M.GaussianDistribution(np.zeros(5), 0.05 ** 2 * np.identity(5))
            pot = M.EvolutionPotential(f, g["y"], noise)
            r = pot.problem().forward(g["u"], want_state=True)
            G, phi, st, w = r["G"].cpu().numpy(), r["phi"].cpu().numpy(), r["state"].cpu().numpy(), r["work"].cpu().numpy()
            print(numerics, N, "state eq", np.array_equal(st, g["end_state"]), "G eq", np.array_equal(G, g["G"]),
                  "phi eq", np.array_equal(phi, g["phi"]), "nfv", w[:, 0].tolist(), g["n_fv"].tolist(),
                  "max rel G", np.max(np.abs(G - g["G"]) / np.abs(g["G"])), "max rel phi", np.max(np.abs(phi - g["phi"]) / np.abs(g["phi"])))


def s_burgers_chain():
    for name, prop, acc in (("chain_burgers_pcn_N64.npz", "pcn", "pcn"), ("chain_burgers_rw_N64.npz", "rw", "rw"),
                            ("chain_burgers_pcn_N128.npz", "pcn", "pcn")):
        g = gold(name)
        N = int(g["N"])
        f = M.BurgersFVM(N=N)
        y = B.BurgersProblem(N).G_params(np.array([0.025, -0.025, -0.02]))
        noise = M.GaussianDistribution(np.zeros(5), 0.05 ** 2 * np.identity(5))
        prior = M.GaussianDistribution(np.array([1.5, .25, -.5]), 0.25 ** 2 * np.identity(3))
        pot = M.EvolutionPotential(f, y, noise)
        n = len(g["normals"])
        if prop == "pcn":
            beta = float(g["beta"])
            spec = M.SamplerSpec(3, _lib.PROPOSE_PCN, _lib.ACCEPT_PCN, coef_u=np.sqrt(1 - beta ** 2), coef_w=beta,
                                 record_start=0, record_interval=1)
        else:
            delta = float(g["delta"])
            spec = M.SamplerSpec(3, _lib.PROPOSE_RW, _lib.ACCEPT_RW, coef_u=1.0, coef_w=np.sqrt(2 * delta),
                                 prior_chol=prior.L, record_start=0, record_interval=1)
        ch = M.ChainBatch(pot.problem(), g["u0"], n_chains=2)
        w = torch.tensor(np.stack([g["normals"]] * 2), device="cuda")
        U = torch.tensor(np.stack([g["uniforms"]] * 2), device="cuda")
        trace = torch.empty((2, n, 3), dtype=torch.float64, device="cuda")
        slog = torch.empty((2, n, 4), dtype=torch.float64, device="cuda")
        ch.run(spec, n, trace=trace, steplog=slog, inject_w=w, inject_u=U)
        tr = trace.cpu().numpy()
        sl = slog.cpu().numpy()
        print(name, "states eq", np.array_equal(tr[0], g["samples"]), np.array_equal(tr[1], g["samples"]),
              "phi_v eq", np.array_equal(sl[0, :, 0], g["phi_v"]), "max rel phi_v", np.max(np.abs(sl[0, :, 0] - g["phi_v"]) / np.abs(g["phi_v"])),
              "accepts", int(sl[0, :, 2].sum()), int(g["accepts"]), "counters", ch.counters.cpu().numpy()[0].tolist())


def s_lorenz():
    lib = _lib.load()
    g = gold("lorenz_rhs.npz")
    for i in range(int(g["n_cases"])):
        K, J = int(g[f"case{i}_K"]), int(g[f"case{i}_J"])
        th = torch.tensor([[float(g[f"case{i}_F"]), float(g[f"case{i}_h"]), float(g[f"case{i}_c"]), float(g[f"case{i}_b"])]] * 7, device="cuda", dtype=torch.float64)
        st = torch.tensor(np.stack([g[f"case{i}_state"]] * 7), device="cuda")
        out = torch.empty_like(st)
        _lib.check(lib.ipmcmc_lorenz_rhs(K, J, 0, 7, th.data_ptr(), st.data_ptr(), out.data_ptr(), None))
        o = out.cpu().numpy()
        print("rhs", K, J, "eq", all(np.array_equal(o[k], g[f"case{i}_rhs"]) for k in range(7)), np.abs(o[0] - g[f"case{i}_rhs"]).max())
    p = gold("lorenz_problem_K6_J4.npz")
    # single attempt vs oracle
    K, J = 6, 4
    theta = np.array([10.1, 9.9, 1.0, 9.9])
    fun = lambda t, s: L.lorenz_rhs(s, K, J, *theta)
    y0 = p["IC"]
    for h in (1e-3, 1e-2, 5e-2):
        yn, fn, err, _ = L.rk45_attempt(fun, 0.0, y0, fun(0, y0), h)
        out = torch.empty((1, 61), dtype=torch.float64, device="cuda")
        _lib.check(lib.ipmcmc_lorenz_rk45_attempt(K, J, 0, 1, torch.tensor([theta], device="cuda").data_ptr(), torch.tensor([y0], device="cuda").data_ptr(),
                                                  torch.tensor([h], device="cuda", dtype=torch.float64).data_ptr(), 1e-3, 1e-6, out.data_ptr(), None))
        o = out.cpu().numpy()[0]
        print("attempt h", h, "ynew rel", np.max(np.abs(o[:30] - yn) / np.abs(yn)), "fnew rel", np.max(np.abs(o[30:60] - fn) / np.abs(fn)), "err", o[60], err, abs(o[60] - err) / err)
    # solves vs golden
    gs = gold("lorenz_solves.npz")
    for i in range(int(gs["n_cases"])):
        T = float(gs[f"case{i}_T"])
        u = gs[f"case{i}_u"]
        f = M.Lorenz96Moments(6, 4, T, 1, p["prior_means"], p["IC"])
        r = f.batch(u.reshape(1, 3), p["IC"].reshape(1, -1))
        G = r["G"].cpu().numpy()[0]
        w = r["work"].cpu().numpy()[0]
        print("solve T", T, "acc/rej", w.tolist(), "ref n_t", int(gs[f"case{i}_n_t"]), "nfev", int(gs[f"case{i}_nfev"]),
              "G max rel", np.max(np.abs(G - gs[f"case{i}_G"]) / (np.abs(gs[f"case{i}_G"]) + 1e-300)),
              "G max |d|/sigma", np.max(np.abs(G - gs[f"case{i}_G"]) / np.sqrt(p["var"])),
              "IC_end max abs", np.abs(r["state"].cpu().numpy()[0] - gs[f"case{i}_IC_end"]).max())
    # timing of a T=20 batch
    f = M.Lorenz96Moments(6, 4, 20.0, 1, p["prior_means"], p["IC"])
    noise = M.GaussianDistribution(np.zeros(30), 0.25 * np.diag(p["var"]))
    pot = M.EvolutionPotential(f, p["y"], noise)
    n = 4096
    u = np.tile(np.array([-1.9, 1.9, 0.9]), (n, 1))
    ic = np.tile(p["IC"], (n, 1))
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = pot.batch(u, ic)
        torch.cuda.synchronize(); dt = time.time() - t0
    w = r["work"].cpu().numpy()
    print("lorenz 4096 solves T=20:", dt, "s; attempts mean", w.sum(1).mean(), "acc", w[:, 0].mean(), "rej", w[:, 1].mean(), "phi mean", r["phi"].mean().item(), "std", r["phi"].std().item())


def s_timing():
    print("fp64 peak TFLOP/s:", M.fp64_peak_tflops())
    for numerics in ("exact", "fused"):
        for (N, n) in ((256, 1024), (256, 8192), (1024, 1024), (1024, 8192), (128, 8192), (64, 8192)):
            f = M.BurgersFVM(N=N, numerics=numerics)
            u = 0.1 * np.random.default_rng(0).standard_normal((n, 3)) + (np.array([0.025, -0.025, -0.02]) - np.array([1.5, .25, -.5]))
            pr = f._problem()
            for rep in range(2):
                torch.cuda.synchronize(); t0 = time.time()
                r = pr.forward(u)
                torch.cuda.synchronize(); dt = time.time() - t0
            nfv = r["work"][:, 0].sum().item()
            print(numerics, "N", N, "chains", n, f"{dt * 1e3:.2f} ms", "solves/s", n / dt, "mean nfv", nfv / n, "TFLOP/s(29/cell-step)", 29.0 * N * nfv / dt / 1e12)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ["rng", "burgers_forward", "burgers_chain", "lorenz", "timing"]
    for w in which:
        section(globals()["s_" + w])
