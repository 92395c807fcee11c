"""Statistical parity (BASELINE.json north_star, correctness leg 2): free-running device chains against
CPU chains of the SAME configuration the benchmark measures -- Burgers pCN beta = 0.25 on 256 cells
(configs[2]) and Lorenz-96 RW delta = 0.125, T = 20 (configs[1]), both with the FUSED numerics -- must
agree in acceptance rate, posterior mean AND variance, and integrated autocorrelation time (hence ESS)
within stated Monte Carlo error.

The CPU chains come from oracle/oracle_c.c (the plain-C restatement of the reference path, pinned bit-
identical to the reference's recorded chains for Burgers and to rounding per RK attempt for Lorenz,
tests/test_oracle_c.py); they run live here in a few seconds.  The device draws Philox noise, the CPU
chains numpy PCG64 noise: the chains are independent realisations of the same Markov kernel.

Every comparison is a two-sample z-test on PER-CHAIN summary statistics (chain acceptance rate, chain
mean, chain variance, chain tau_int over the same window on both sides), whose standard error is
estimated from the spread across chains -- which accounts for the autocorrelation inside a chain without
modelling it.  Tolerance: |z| < 4 (seeds are fixed, so the outcome is deterministic; 4 sigma leaves room
for 3 x 4 simultaneous comparisons), tau_int within 25 %."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import burgers_np as B
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu
Z_MAX = 4.0


@pytest.fixture(scope="module")
def G():
    import gpu_common
    return gpu_common


def _z(a, b):
    """two-sample z statistic of the means of per-chain statistics a [n_a, ...] and b [n_b, ...]"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    se = np.sqrt(a.var(0, ddof=1) / a.shape[0] + b.var(0, ddof=1) / b.shape[0])
    return (a.mean(0) - b.mean(0)) / se


def _chain_stats(states, accepted, burn, stats):
    """states [n_chains, n, d], accepted [n_chains, n] -> per-chain acceptance, mean, variance, tau_int (min-ESS parameter)"""
    x = states[:, burn:]
    acc = accepted[:, burn:].mean(1)
    tau = np.array([[stats.integrated_autocorr_time(c[:, j]) for j in range(x.shape[2])] for c in x])
    return dict(acc=acc, mean=x.mean(1), var=x.var(1, ddof=1), tau=tau)


def _compare(dev, cpu, what):
    z_acc = _z(dev["acc"], cpu["acc"])
    z_mean = _z(dev["mean"], cpu["mean"])
    z_var = _z(dev["var"], cpu["var"])
    tau_d, tau_c = dev["tau"].mean(0), cpu["tau"].mean(0)
    msg = "%s: acceptance dev %.4f cpu %.4f (z %.2f); mean z %s; var z %s; tau_int dev %s cpu %s" % (
        what, dev["acc"].mean(), cpu["acc"].mean(), z_acc, np.round(z_mean, 2), np.round(z_var, 2),
        np.round(tau_d, 1), np.round(tau_c, 1))
    print(msg)
    assert abs(z_acc) < Z_MAX, msg
    assert np.all(np.abs(z_mean) < Z_MAX), msg
    assert np.all(np.abs(z_var) < Z_MAX), msg
    assert np.all(np.abs(tau_d - tau_c) < 0.25 * tau_c), msg
    return msg


def test_burgers_bench_config_statistics_vs_cpu_chains(G):
    """configs[2]: N = 256, pCN beta = 0.25, FUSED numerics, posterior-region start (bench.py), 1024 device
    chains vs 32 CPU chains x 2500 steps, the first 500 discarded."""
    import ip_mcmc_b200 as M
    N, beta, n_steps, burn = 256, 0.25, 2500, 500
    y = B.BurgersProblem(N).G_params(G.TRUTH)          # the reference's noise-free data G(u*), both sides
    f, pot, prior, _ = G.burgers_setup(N, "fused", y=y)
    start = G.TRUTH - G.PRIOR_MEAN
    s = M.MCMCSampler(M.ConstSteppCNProposer(beta, prior), M.CountedAccepter(M.pCNAccepter(pot)), np.random.default_rng(2))
    dev_states = s.run(start, n_steps, 0, 1, n_chains=1024)
    acc_dev = np.any(np.diff(np.concatenate([np.broadcast_to(start, (1024, 1, 3)), dev_states], axis=1), axis=1) != 0, axis=2)
    assert acc_dev.sum() == s.last_run["counters"]["accepts"]
    assert s.last_run["counters"]["nonfinite"] == 0
    np.testing.assert_allclose(f.at_parameters(G.TRUTH), y, rtol=1e-10)      # FUSED G(u*) vs the reference's
    Pc = CO.BurgersC(N, y=y, noise_cov=G.NOISE_COV)
    rng = np.random.default_rng(77)
    n_cpu = 32
    r = Pc.run_chains(start, 0.25 * rng.standard_normal((n_cpu, n_steps, 3)), rng.random((n_cpu, n_steps)), CO.PCN, CO.PCN, beta)
    dev = _chain_stats(dev_states, acc_dev, burn, M.stats)
    cpu = _chain_stats(r["u"], r["accepted"], burn, M.stats)
    _compare(dev, cpu, "burgers pCN 256 cells")
    # the on-device Welford moments are the moments of the recorded states
    flat = dev_states.reshape(-1, 3)
    np.testing.assert_allclose(s.last_run["pooled_mean"], flat.mean(0), rtol=1e-10)
    np.testing.assert_allclose(s.last_run["pooled_var"], flat.var(0, ddof=1), rtol=1e-9)
    # stationary acceptance of this problem is 0.05-0.08 (bench: 0.077 incl. the start transient)
    assert 0.04 < dev["acc"].mean() < 0.09


def test_lorenz_bench_config_statistics_vs_cpu_chains(G):
    """configs[1]: K = 6, J = 4, T = 20, RW delta = 0.125, r = 0.5, FUSED numerics, the reference's start
    u_0 = (-1.9, 1.9, 0.9) (lorenz_mcmc.py:139); 640 device chains vs 24 CPU chains x 700 steps, the first
    200 discarded; 2 solves per step with the carried initial condition on both sides."""
    import ip_mcmc_b200 as M
    p = golden("lorenz_problem_K6_J4.npz")
    T, delta, n_steps, burn = 20.0, 0.125, 700, 200
    prior_cov = np.diag([10., 1, 10])
    noise_cov = 0.5 ** 2 * np.diag(p["var"])
    f = M.Lorenz96Moments(6, 4, T, 1.0, p["prior_means"], p["IC"], numerics="fused")
    pot = M.EvolutionPotential(f, p["y"], M.GaussianDistribution(np.zeros(30), noise_cov))
    prior = M.GaussianDistribution(np.zeros(3), prior_cov)
    s = M.MCMCSampler(M.ConstStepStandardRWProposer(delta, prior), M.CountedAccepter(M.StandardRWAccepter(pot, prior)),
                      np.random.default_rng(1))
    n_dev = 640
    dev_states = s.run(p["u0"], n_steps, 0, 1, n_chains=n_dev)
    acc_dev = np.any(np.diff(np.concatenate([np.broadcast_to(p["u0"], (n_dev, 1, 3)), dev_states], axis=1), axis=1) != 0, axis=2)
    assert acc_dev.sum() == s.last_run["counters"]["accepts"]
    c = s.last_run["counters"]
    attempts_dev = (c["work_a"] + c["work_b"]) / (2.0 * n_dev * n_steps)
    Lc = CO.LorenzC(6, 4, T, 1.0, p["prior_means"], y=p["y"], noise_cov=noise_cov)
    rng = np.random.default_rng(78)
    n_cpu = 24
    z = rng.standard_normal((n_cpu, n_steps, 3)) * np.sqrt(np.diag(prior_cov))
    r = Lc.run_chains(p["u0"], p["IC"], z, rng.random((n_cpu, n_steps)), CO.RW, CO.RW, delta, prior_cov=prior_cov)
    attempts_cpu = r["work"].sum() / (2.0 * n_cpu * n_steps)
    dev = _chain_stats(dev_states, acc_dev, burn, M.stats)
    cpu = _chain_stats(r["u"], r["accepted"], burn, M.stats)
    _compare(dev, cpu, "lorenz RW T=20")
    # same integrator work: RK45 attempts per solve within 2 % (controller parity at the statistical level)
    assert abs(attempts_dev - attempts_cpu) < 0.02 * attempts_cpu, (attempts_dev, attempts_cpu)


@pytest.mark.parametrize("kind", ["pcn", "rw"])
def test_flat_likelihood_exact_stationary_law(G, kind):
    """Sampler-correctness control with a KNOWN answer ('exact check'): with a noise covariance of
    1e12 the misfit is flat to 1e-10, so the stationary law of the chain is the Gaussian its proposer / accepter
    pair targets, N(0, C) -- for pCN because every proposal is accepted and v = sqrt(1-beta^2) u + beta w is
    an AR(1) process with that law (proposer.py:78-81), for RW because exp(I(u) - I(v)), I(w) = 1/2 |L w|^2
    (accepter.py:99-106), is the Metropolis ratio of N(0, (L^T L)^-1) = N(0, C) when C = I.  The end states of
    4096 independent chains are 4096 independent draws: mean and variance must agree with 0 and 1 within
    4 standard errors (sd/sqrt(n) and sqrt(2/(n-1))), the pooled device moments with the recorded states."""
    import ip_mcmc_b200 as M
    n_chains, n_steps = 4096, 300
    # the step cap is lifted: the blow-up solves of the reference's interior-only CFL (0.2 % of the steps under
    # this wide prior) finish with a finite, equally flat misfit instead of being rejected as capped
    f = M.BurgersFVM(N=32, numerics="fused", max_fv_steps=10 ** 7)
    prior = M.GaussianDistribution(G.PRIOR_MEAN, np.identity(3))
    pot = M.EvolutionPotential(f, f.at_parameters(G.TRUTH), M.GaussianDistribution(np.zeros(5), 1e12 * np.identity(5)))
    if kind == "pcn":
        s = M.MCMCSampler(M.ConstSteppCNProposer(0.5, prior), M.CountedAccepter(M.pCNAccepter(pot)), np.random.default_rng(5))
    else:
        s = M.MCMCSampler(M.ConstStepStandardRWProposer(0.5, prior), M.CountedAccepter(M.StandardRWAccepter(pot, prior)),
                          np.random.default_rng(6))
    states = s.run(np.zeros(3), n_steps, 0, 1, n_chains=n_chains)
    c = s.last_run["counters"]
    assert c["nonfinite"] < 1e-4 * c["calls"], c
    end = states[:, -1]
    z_mean = end.mean(0) * np.sqrt(n_chains) / end.std(0, ddof=1)
    z_var = (end.var(0, ddof=1) - 1.0) / np.sqrt(2.0 / (n_chains - 1))
    msg = "%s: acceptance %.4f, end-state mean z %s, variance z %s" % (kind, c["accepts"] / c["calls"], np.round(z_mean, 2), np.round(z_var, 2))
    print(msg)
    assert np.all(np.abs(z_mean) < Z_MAX) and np.all(np.abs(z_var) < Z_MAX), msg
    if kind == "pcn":
        assert c["accepts"] >= 0.999 * (c["calls"] - c["nonfinite"]), msg
        # AR(1) with rho = sqrt(1 - beta^2): lag-1 autocorrelation of the recorded chains
        x = states[:, 100:]
        rho = np.mean(x[:, 1:] * x[:, :-1]) / np.mean(x * x)
        assert abs(rho - np.sqrt(0.75)) < 5e-3, rho
    else:
        assert 0.2 < c["accepts"] / c["calls"] < 0.6, msg   # RW Metropolis on N(0, I_3), proposal sd 1
    flat = states.reshape(-1, 3)
    np.testing.assert_allclose(s.last_run["pooled_mean"], flat.mean(0), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(s.last_run["pooled_var"], flat.var(0, ddof=1), rtol=1e-9)
