"""CPU tests of the host-side mirror of the reference interface and of the C-ABI library.
The values restate the reference's own unit tests (ip_mcmc/ip_mcmc/*_test.py); no compute call
into the CUDA library is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

import ip_mcmc_b200 as M
from ip_mcmc_b200 import _lib
from conftest import ROOT, golden


class MockRNG(np.random.Generator):
    """Same seam as the reference's test_utilities.MockRNG (test_utilities.py:11-26)."""

    def __init__(self, result):
        super().__init__(np.random.PCG64(0))
        self.result = result

    def multivariate_normal(self, mean, cov, *a, **k):
        return self.result * np.ones_like(mean)

    def random(self, *a, **k):
        return self.result if 0 <= self.result <= 1 else 0.5


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ipmcmc.h")).read()
    declared = set(re.findall(r"\b(ipmcmc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()                       # built by __graft_entry__.build()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ipmcmc_abi_version() == 2


def test_ctypes_struct_layout_matches_header_sizes():
    # field order and 8-byte alignment as declared in include/ipmcmc.h
    assert ctypes.sizeof(_lib.PotentialDesc) == 48
    assert ctypes.sizeof(_lib.BurgersDesc) == 16 + 24 + 32 + 16 + 48
    assert ctypes.sizeof(_lib.LorenzDesc) == 16 + 32 + 8 + 48
    assert ctypes.sizeof(_lib.SamplerDesc) == 32 + 16 + 16 + 40 + 40
    assert ctypes.sizeof(_lib.ChainBuffers) == 18 * 8
    assert ctypes.sizeof(_lib.HostIO) == 10 * 8


def test_argument_validation_without_gpu():
    """Error convention: negative return code + message, surfaced as ValueError/EngineError."""
    lib = _lib.load()
    d = _lib.BurgersDesc()
    d.n_cells = 1
    h = ctypes.c_void_p()
    rc = lib.ipmcmc_burgers_create(ctypes.byref(d), ctypes.byref(h))
    assert rc == -1 and b"n_cells" in lib.ipmcmc_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    d.n_cells = 5000
    assert lib.ipmcmc_burgers_create(ctypes.byref(d), ctypes.byref(h)) == -3
    d.n_cells, d.n_kl_modes, d.n_params = 256, 254, 257           # beyond IPMCMC_MAX_DIM_WIDE
    assert lib.ipmcmc_burgers_create(ctypes.byref(d), ctypes.byref(h)) == -3 and b"n_kl_modes" in lib.ipmcmc_last_error()
    d.n_cells, d.n_kl_modes, d.n_params = 2048, 40, 43             # the wide path has no team solver
    assert lib.ipmcmc_burgers_create(ctypes.byref(d), ctypes.byref(h)) == -3 and b"wide path" in lib.ipmcmc_last_error()
    ld = _lib.LorenzDesc()
    ld.K, ld.J = 40, 4
    assert lib.ipmcmc_lorenz_create(ctypes.byref(ld), ctypes.byref(h)) == -3


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    f = M.BurgersFVM(N=32)
    with pytest.raises(_lib.EngineError):
        f(np.zeros(3))


def test_gaussian_distribution_reference_kats():
    # distribution_test.py:7-24
    g = M.GaussianDistribution(0, 2)
    assert g.k == 1
    assert np.allclose(g.apply_covariance(3), 6) and np.allclose(g.apply_sqrt_covariance(3), 3 * np.sqrt(2))
    assert np.allclose(g.apply_precision(3), 1.5) and np.allclose(g.apply_sqrt_precision(3), 3 / np.sqrt(2))
    g = M.GaussianDistribution(np.zeros(3), np.diag([1., 2, 3]))
    x = np.ones(3)
    assert np.allclose(g.apply_covariance(x), [1, 2, 3]) and np.allclose(g.apply_sqrt_covariance(x), np.sqrt([1, 2, 3]))
    assert np.allclose(g.apply_precision(x), [1, 1 / 2, 1 / 3])
    assert np.allclose(g.apply_sqrt_precision(x), 1 / np.sqrt([1, 2, 3]))
    with pytest.raises(AssertionError):
        M.GaussianDistribution(np.zeros(3), np.identity(2))


def test_gaussian_tables_match_reference_fixture():
    k = golden("operator_kats.npz")
    g = M.GaussianDistribution(np.zeros(3), k["cov"])
    assert np.allclose([g.logpdf(x) for x in k["xs"]], k["logpdf"], rtol=1e-14)
    assert np.allclose([g.apply_sqrt_covariance(x) for x in k["xs"]], k["sqrtcov"], rtol=1e-15)
    assert np.allclose(g.sample_factor() @ k["z"], k["w"], rtol=1e-14)
    assert np.array_equal(g.sample(np.random.default_rng(11)), k["w"])
    LP, logdet, rank = g.whitener()
    manual = [-0.5 * (rank * np.log(2 * np.pi) + logdet + np.sum(np.square(x @ LP))) for x in k["xs"]]
    assert np.allclose(manual, k["logpdf"], rtol=1e-14)


def test_potential_tables_diag_and_dense():
    from ip_mcmc_b200.engine import potential_tables
    p = golden("lorenz_problem_K6_J4.npz")
    tab = potential_tables(p["y"], M.GaussianDistribution(np.zeros(30), 0.25 * np.diag(p["var"])))
    assert not tab["dense"] and sorted(tab["perm"]) == list(range(30))
    # the permuted whitening reproduces scipy's ordering: r_i = dev[perm_i]*scale_i
    dev = np.random.default_rng(0).standard_normal(30)
    LP, _, _ = M.GaussianDistribution(np.zeros(30), 0.25 * np.diag(p["var"])).whitener()
    assert np.array_equal(dev[tab["perm"]] * tab["scale"], dev @ LP)
    A = np.random.default_rng(1).standard_normal((4, 4))
    tab = potential_tables(np.zeros(4), M.GaussianDistribution(np.zeros(4), A @ A.T + np.identity(4)))
    assert tab["dense"] and tab["LP"].shape == (4, 4)


def test_proposer_reference_kats():
    # proposer_test.py:8-58
    prior = M.GaussianDistribution(0, 1)
    p = M.ConstStepStandardRWProposer(2, prior)
    assert np.isclose(p.prefactor, 2)
    assert np.allclose(p(np.array([1.0]), MockRNG(3)), 1 + 2 * 3)
    prior3 = M.GaussianDistribution(np.array([5., 5, 5]), np.identity(3))
    p = M.ConstStepStandardRWProposer(0.125, prior3)
    assert np.allclose(p(np.array([1., 2, 3]), MockRNG(1)), np.array([1, 2, 3]) + 0.5)
    assert np.all(p.w.mean == 0)                                   # prior mean ignored
    p = M.ConstSteppCNProposer(0.25, prior)
    assert np.isclose(p.contraction, np.sqrt(1 - 0.0625))
    assert np.allclose(p(np.array([2.0]), MockRNG(3)), np.sqrt(1 - 0.0625) * 2 + 0.75)
    cov = np.array([[2., 1], [1, 2]])
    p = M.ConstSteppCNProposer(0.5, M.GaussianDistribution(np.zeros(2), cov))
    assert np.allclose(p(np.array([1., 2]), MockRNG(2)), np.sqrt(0.75) * np.array([1, 2]) + 1)
    with pytest.raises(AssertionError):
        M.ConstSteppCNProposer(1.5, prior)
    assert M.pCNProposer is M.ConstSteppCNProposer


def test_varstep_proposers_preincrement_and_device_tables():
    # proposer.py:54-55, 111-112: the first proposal uses schedule(1)
    prior = M.GaussianDistribution(0, 1)
    seen = []
    p = M.VarStepStandardRWProposer(lambda i: seen.append(i) or 0.5 * i, prior)
    v = p(np.array([0.0]), MockRNG(1))
    assert seen == [1] and np.allclose(v, np.sqrt(2) * np.sqrt(0.5))
    spec = p.device_spec(3)
    assert seen == [1, 2, 3, 4] and np.allclose(spec["schedule"][:, 1], np.sqrt(2) * np.sqrt(0.5 * np.array([2, 3, 4])))
    assert np.all(spec["schedule"][:, 0] == 1.0) and p.i == 4
    q = M.VarSteppCNProposer(lambda i: 0.1 * i, prior)
    s = q.device_spec(2)["schedule"]
    assert np.allclose(s, [[np.sqrt(1 - 0.01), 0.1], [np.sqrt(1 - 0.04), 0.2]])
    assert np.allclose(q(np.array([1.0]), MockRNG(2)), np.sqrt(1 - 0.09) + 0.3 * 2)


class _Pot:
    def __init__(self, f):
        self.f = f

    def __call__(self, u):
        return self.f(u)


def test_accepter_reference_kats():
    # accepter_test.py:20-42 with lambda potentials
    prior = M.GaussianDistribution(0, 2)
    a = M.StandardRWAccepter(_Pot(lambda u: float(u[0])), prior)
    assert np.isclose(a._I(np.array([1.0])), 2) and np.isclose(a._I(np.array([5.0])), 30)
    assert np.isclose(a.accept_probability(np.array([5.0]), np.array([1.0])), np.exp(28))
    b = M.pCNAccepter(_Pot(lambda u: float(u[0])))
    assert np.isclose(b.accept_probability(np.array([np.log(2)]), np.array([0.0])), 2)       # un-clipped
    c = M.pCNAccepter(_Pot(lambda u: 0.0))
    assert c(np.zeros(1), np.zeros(1), MockRNG(0.999)) and not c(np.zeros(1), np.zeros(1), MockRNG(1.0))
    nanp = M.pCNAccepter(_Pot(lambda u: np.nan))
    assert not nanp(np.zeros(1), np.zeros(1), MockRNG(0.0))


def test_counted_and_constrain_accepters():
    # accepter.py:13-55
    inner = M.pCNAccepter(_Pot(lambda u: 0.0))
    c = M.CountedAccepter(inner)
    with pytest.raises(ValueError):
        c.ratio()
    assert c(np.zeros(1), np.zeros(1), MockRNG(0.3)) and c.calls == 1 and c.accepts == 1
    c.reset()
    assert c.calls == 0

    class CountingRNG(MockRNG):
        n = 0

        def random(self, *a, **k):
            CountingRNG.n += 1
            return 0.3

    box = M.BoxConstraint([-1.0], [1.0], shift=[-0.5])
    ca = M.ConstrainAccepter(inner, box)
    rng = CountingRNG(0.3)
    assert ca(np.zeros(1), np.array([1.2]), rng) and CountingRNG.n == 1       # 1.2 - 0.5 inside
    assert not ca(np.zeros(1), np.array([1.6]), rng) and CountingRNG.n == 1   # rejected, no U drawn
    from ip_mcmc_b200.accepter import device_spec
    spec = device_spec(M.CountedAccepter(ca))
    assert spec["constraint"] is box and spec["kind"] == _lib.ACCEPT_PCN and spec["outer_counted"]
    with pytest.raises(TypeError):
        device_spec(M.ConstrainAccepter(inner, lambda v: True))
    # nesting decides what a counter sees (accepter.py:20-27, 52-55): inside the ConstrainAccepter it never
    # sees a constraint-rejected proposal; the device counters are credited accordingly
    from ip_mcmc_b200.accepter import credit_counters
    c_in, c_out = M.CountedAccepter(inner), None
    c_out = M.CountedAccepter(M.ConstrainAccepter(c_in, box))
    rng = MockRNG(0.3)
    for v in (1.2, 1.6, 0.9, 1.7, -0.9):                     # 1.6, 1.7 and -0.9 violate -1 < v - 0.5 < 1
        c_out(np.zeros(1), np.array([v]), rng)
    assert (c_out.calls, c_out.accepts, c_in.calls, c_in.accepts) == (5, 2, 2, 2)
    spec = device_spec(c_out)
    assert [(c is c_out, inside) for c, inside in spec["counted"]][0] == (True, False)
    assert [(c is c_in, inside) for c, inside in spec["counted"]][1] == (True, True)
    c_out.reset(), c_in.reset()
    credit_counters(spec, calls=5, accepts=2, constraint_rejects=3)
    assert (c_out.calls, c_out.accepts, c_in.calls, c_in.accepts) == (5, 2, 2, 2)


def test_burgers_grid_tables_match_reference_fixture():
    for N in (100, 128, 200, 256):
        g = golden(f"burgers_forward_N{N}.npz")
        f = M.BurgersFVM(N=N)
        assert np.array_equal(f.x, g["x"]) and f.dx == g["dx"] and f.dx_meas == g["dx_meas"]
        assert np.array_equal(f.left_limits, g["left"]) and np.array_equal(f.right_limits, g["right"])


def test_stats_autocorr_and_ess():
    assert np.array_equal(M.MCMCSampler.autocorr(np.ones(5)), np.ones(5))           # sampler.py:49-53
    assert np.allclose(M.MCMCSampler.autocorr(np.array([1., -1, 1, -1])), [1, -0.75, 0.5, -0.25])
    rng = np.random.default_rng(0)
    x = rng.standard_normal(5000)
    assert np.allclose(M.stats.autocorr(x)[:50], np.correlate(x - x.mean(), x - x.mean(), "full")[-5000:][:50] / np.sum((x - x.mean()) ** 2))
    assert 0.8 < M.stats.ess(x[:, None])[0] / 5000 <= 1.0
    ar = np.zeros(20000)
    for i in range(1, 20000):
        ar[i] = 0.9 * ar[i - 1] + rng.standard_normal()
    tau = M.stats.integrated_autocorr_time(ar)
    assert 14 < tau < 25                                                            # (1+rho)/(1-rho) = 19
    n, mean, m2 = M.stats.merge_moments([(len(a), a.mean(0), ((a - a.mean(0)) ** 2).sum(0)) for a in np.split(x[:4998].reshape(-1, 1), 3)])
    assert n == 4998 and np.allclose(mean, x[:4998].mean()) and np.allclose(m2 / (n - 1), x[:4998].var(ddof=1))


def test_tau0_matches_reference_fixture():
    """stats.uncorrelated_sample_spacing == the reference's utilities.uncorrelated_sample_spacing
    (utilities.py:169-186) on series generated by oracle/make_golden.py:gen_tau0, incl. the chain that is
    too short (the reference then returns len(x), the number of variables) and helpers.autocorrelation."""
    g = golden("tau0_reference.npz")
    for i in range(int(g["n_cases"])):
        x = g[f"case{i}_x"]
        assert M.stats.uncorrelated_sample_spacing(x) == int(g[f"case{i}_tau0"])
        if x.shape[1] >= 20:
            assert np.array_equal(M.stats.windowed_autocorrelation(x, 20), g[f"case{i}_ac20"])
    assert M.stats.uncorrelated_sample_spacing(g["case5_x"]) == 3


def test_chain_segments_match_the_reference_split():
    """studies.chain_segments == the halving loop of burgers_wasserstein_chain.py:261-266."""
    from ip_mcmc_b200 import studies
    for n in (5000, 4999, 37, 16):
        samples, want = np.arange(n), []
        for _ in range(4):
            l = int(len(samples) / 2)
            want.append(samples[l + 1:])
            samples = samples[:l]
        got = [np.arange(a, b) for a, b in studies.chain_segments(n)]
        assert all(np.array_equal(g, w) for g, w in zip(got, want))


def test_studies_histogram_cache_and_schedule(tmp_path):
    from ip_mcmc_b200 import studies
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5000, 3)) * 0.3
    x[0] = [0.5, 0.5, 0.5]                      # upper edge belongs to the last bin
    iv = [(-0.5, 0.5)] * 3
    h = studies.histogram3d(x, iv, bins=20).numpy()
    ref, _ = np.histogramdd(x, bins=20, range=iv, density=True)
    np.testing.assert_allclose(h, ref, rtol=1e-12)
    calls = []
    f = lambda a, b: calls.append(1) or np.arange(a * b, dtype=float).reshape(a, b)
    r1 = studies.load_or_compute("chain", f, (3, 2), data_dir=str(tmp_path))
    r2 = studies.load_or_compute("chain", f, (3, 2), data_dir=str(tmp_path))
    assert len(calls) == 1 and np.array_equal(r1, r2) and (tmp_path / "chain.npy").exists()
    g = golden("chain_burgers_varstep_rw_N64.npz")          # the reference's PWLinear(0.1, 0.001, 50)
    s = studies.PWLinear(0.1, 0.001, 50)
    assert np.array_equal([s(i) for i in range(1, 101)], g["schedule"]) and repr(s) == "pwl_0.1_0.001_50"
