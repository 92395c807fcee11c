"""Pins oracle/mcmc_np.py against the reference's unit tests (proposer_test.py, accepter_test.py,
sampler_test.py, distribution_test.py -- values restated here, the reference tree is not read)
and against chains recorded from the live reference with a noise tape (tests/golden/chain_*.npz)."""
import numpy as np
import pytest
import scipy.linalg

from conftest import golden
from oracle import burgers_np as B
from oracle import mcmc_np as M

PRIOR_COV = 0.25 ** 2 * np.identity(3)
NOISE_COV = 0.05 ** 2 * np.identity(5)


def test_proposer_kats():
    # proposer_test.py:8-33 (RW: prefactor sqrt(2 delta)); :36-58 (pCN: sqrt(1-beta^2) u + beta w)
    delta = 2
    assert np.isclose(M.propose(M.RW, np.array([0.]), np.array([1.]), delta)[0], np.sqrt(2 * delta))
    v = M.propose(M.RW, np.array([1., 2, 3]), np.array([.5, .5, .5]), 0.125)
    assert np.allclose(v, np.array([1, 2, 3]) + 0.5 * 0.5)
    beta = 0.25
    v = M.propose(M.PCN, np.array([1., 2]), np.array([3., 3]), beta)
    assert np.allclose(v, np.sqrt(1 - beta ** 2) * np.array([1, 2]) + beta * 3)


def test_rw_regulariser_uses_L_not_inverse():
    # accepter_test.py:20-33: prior covariance 2 -> _I(1) = Phi + .5*||sqrt(2)*1||^2 = 1+1, _I(5) = 5+25
    L = np.array([[np.sqrt(2.)]])
    assert np.isclose(1 + M.prior_regulariser(np.array([1.]), L), 2)
    assert np.isclose(5 + M.prior_regulariser(np.array([5.]), L), 30)


def test_accept_rule_strict_and_unclipped():
    # accepter_test.py:9-17, :36-42
    assert M.accept_probability(M.PCN, 1.0, 1.0 - np.log(2)) == pytest.approx(2.0)   # un-clipped
    pot = lambda u: 0.0
    out = M.run_chain(pot, np.zeros(1), np.ones((1, 1)), np.array([1.0 - 1e-16]), M.PCN, M.PCN, 0.5)
    assert out["accepted"][0]            # a = 1 > U
    nanpot = lambda u: np.nan
    out = M.run_chain(nanpot, np.zeros(1), np.ones((1, 1)), np.array([0.0]), M.PCN, M.PCN, 0.5)
    assert not out["accepted"][0]        # NaN > U is False


def test_sampler_step_count():
    # sampler_test.py:14-16: n=100, burn_in=20, interval=10 -> 1010 ... restated: 100,20,10 -> 280? no:
    # calls = max(0, burn_in - si) + n*si
    assert M.sampler_total_steps(27, 20, 10) == 280
    assert M.sampler_total_steps(5, 0, 1) == 5
    st = np.arange(280)[:, None] * np.ones((1, 2))
    s = M.samples_from_states(st, 27, 20, 10)
    assert s.shape == (27, 2) and s[0, 0] == 19 and s[-1, 0] == 279


def test_gaussian_kats():
    g = golden("operator_kats.npz")
    LP, logdet, rank = M.psd_whitener(g["cov"])
    for x, lp in zip(g["xs"], g["logpdf"]):
        assert M.gaussian_logpdf(x, LP, logdet, rank) == pytest.approx(lp, rel=1e-14)
    L = np.tril(scipy.linalg.cho_factor(g["cov"], lower=True)[0])
    assert np.allclose(g["sqrtcov"], g["xs"] @ L.T, rtol=1e-15)
    assert np.allclose(M.mvn_factor(g["cov"]) @ g["z"], g["w"], rtol=1e-14)
    LP, logdet, rank = M.psd_whitener(NOISE_COV)
    for x, lp in zip(g["devs"], g["logpdf_diag"]):
        assert M.gaussian_logpdf(x, LP, logdet, rank) == lp
    # distribution_test.py:7-24: covariance diag(1,2,3): sqrt-cov is diag(1, sqrt2, sqrt3)
    L = np.tril(scipy.linalg.cho_factor(np.diag([1., 2, 3]), lower=True)[0])
    assert np.allclose(L @ np.ones(3), np.sqrt([1, 2, 3]))


def _replay(name, proposer, accepter, **kw):
    g = golden(name)
    N = int(g["N"])
    P = B.BurgersProblem(N)
    y = P.G_params(np.array([0.025, -0.025, -0.02]))
    pot = M.Potential(P, y, NOISE_COV)
    out = M.run_chain(pot, g["u0"], g["normals"], g["uniforms"], proposer, accepter,
                      prior_cov=PRIOR_COV, **kw)
    return g, out


def test_replay_reference_pcn_chain():
    for name in ("chain_burgers_pcn_N64.npz", "chain_burgers_pcn_N128.npz"):
        g, out = _replay(name, M.PCN, M.PCN, step=float(golden(name)["beta"]))
        assert np.array_equal(out["u"], g["samples"])
        assert np.array_equal(out["v"], g["v"])
        assert np.array_equal(out["phi_v"], g["phi_v"])
        assert np.array_equal(out["phi_u"], g["phi_u"])
        assert out["accepts"] == int(g["accepts"]) and out["calls"] == int(g["calls"])


def test_replay_reference_rw_chain():
    g, out = _replay("chain_burgers_rw_N64.npz", M.RW, M.RW, step=float(golden("chain_burgers_rw_N64.npz")["delta"]))
    assert np.array_equal(out["u"], g["samples"])
    assert np.array_equal(out["phi_v"], g["phi_v"])
    assert out["accepts"] == int(g["accepts"])


def test_replay_reference_varstep_chain():
    g = golden("chain_burgers_varstep_rw_N64.npz")
    g, out = _replay("chain_burgers_varstep_rw_N64.npz", M.RW, M.RW, step=g["schedule"], varstep=True)
    assert np.array_equal(out["u"], g["samples"])
    assert out["accepts"] == int(g["accepts"])


def test_replay_reference_constrained_chain():
    g = golden("chain_burgers_constrained_rw_N64.npz")
    lo, hi = float(g["lo"]), float(g["hi"])
    ok = lambda v: lo < v[2] + (-0.5) < hi
    g, out = _replay("chain_burgers_constrained_rw_N64.npz", M.RW, M.RW, step=float(g["delta"]), constraint=ok)
    assert np.array_equal(out["u"], g["samples"])
    assert out["uniforms_used"] == len(g["uniforms"]) < len(g["normals"])
    assert out["accepts"] == int(g["accepts"]) and out["calls"] == int(g["calls"])


def test_autocorr_definition():
    # sampler.py:43-54: lag-0 normalised, constant series -> ones
    assert np.array_equal(M.autocorr(np.ones(5)), np.ones(5))
    x = np.array([1., -1, 1, -1])
    assert np.allclose(M.autocorr(x), [1, -0.75, 0.5, -0.25])
