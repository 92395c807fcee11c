"""Pins oracle/lorenz_np.py against the reference's in-file Lorenz-96 known-answer tests
(report/scripts/lorenz.py:114-171, values restated), against fixtures from the live reference
(tests/golden/lorenz_*.npz) and against scipy's RK45 (third-party dependency of the reference)."""
import numpy as np
import pytest

from conftest import golden
from oracle import lorenz_np as L
from oracle import mcmc_np as M


def rhs(K, J, F, h, c, b, state):
    return L.lorenz_rhs(np.array(state, dtype=float), K, J, F, h, c, b)


def test_reference_rhs_kats():
    # lorenz.py:115-147
    assert np.allclose(rhs(3, 1, 2, 1, 1, 1, [0, 0, 0, 0, 0, 0]), [2, 2, 2, 0, 0, 0])
    assert np.allclose(rhs(4, 1, 0, 0, 0, 0, [1, 2, 3, 4, 0, 0, 0, 0]), [-5, -3, 3, -7, 0, 0, 0, 0])
    assert np.allclose(rhs(1, 4, 0, 0, 1, 2, [0, 1, 2, 3, 4]), [0, 3, -20, 5, -2])
    assert np.allclose(rhs(1, 1, 0, 2, -1, 0, [1, 0]), [-1, -2])
    assert np.allclose(rhs(1, 2, 0, 2, -1, 0, [0, 1, 2]), [3, 1, 2])
    assert np.allclose(rhs(2, 2, 1, 1, 1, 1, [2, 3, 4, 5, 6, 7]), [-2.5, -10.5, 2, -8, 2.5, -11.5])


def test_rhs_bit_identical_to_reference():
    g = golden("lorenz_rhs.npz")
    for i in range(int(g["n_cases"])):
        K, J = int(g[f"case{i}_K"]), int(g[f"case{i}_J"])
        r = L.lorenz_rhs(g[f"case{i}_state"], K, J, float(g[f"case{i}_F"]), float(g[f"case{i}_h"]),
                         float(g[f"case{i}_c"]), float(g[f"case{i}_b"]))
        assert np.array_equal(r, g[f"case{i}_rhs"])


def test_tableau_matches_scipy():
    from scipy.integrate._ivp.rk import RK45, SAFETY, MIN_FACTOR, MAX_FACTOR
    assert np.array_equal(RK45.A, L.RK45_A) and np.array_equal(RK45.B, L.RK45_B)
    assert np.array_equal(RK45.C, L.RK45_C) and np.array_equal(RK45.E, L.RK45_E)
    assert (SAFETY, MIN_FACTOR, MAX_FACTOR) == (L.SAFETY, L.MIN_FACTOR, L.MAX_FACTOR)
    assert RK45.error_estimator_order == 4


def test_solve_bit_identical_to_scipy():
    from scipy.integrate import solve_ivp
    p = golden("lorenz_problem_K6_J4.npz")
    fun = lambda t, s: L.lorenz_rhs(s, 6, 4, 10.1, 9.9, 1, 9.9)
    r = solve_ivp(fun=fun, t_span=(0, 3.0), y0=p["IC"], method="RK45")
    sol = L.rk45_solve(fun, p["IC"], 3.0)
    assert np.array_equal(r.t, sol["t"]) and np.array_equal(r.y, sol["y"])
    assert r.nfev == sol["nfev"]


def test_solves_match_reference_fixture():
    g = golden("lorenz_solves.npz")
    p = golden("lorenz_problem_K6_J4.npz")
    for i in range(int(g["n_cases"])):
        T = float(g[f"case{i}_T"])
        if T > 5:
            continue    # kept for the GPU statistical tests; 1 s each on CPU
        op = L.LorenzProblem(6, 4, T, 1, p["prior_means"], p["IC"])
        G = op(g[f"case{i}_u"])
        assert np.array_equal(G, g[f"case{i}_G"])
        assert np.array_equal(op.IC, g[f"case{i}_IC_end"])
        assert op.last["t"].size == int(g[f"case{i}_n_t"]) and op.last["nfev"] == int(g[f"case{i}_nfev"])
        assert np.array_equal(op.last["t"][:8], g[f"case{i}_t_head"][:op.last["t"].size])


def test_problem_constants():
    p = golden("lorenz_problem_K6_J4.npz")
    assert p["y"].shape == (30,) and p["var"].shape == (30,) and p["IC"].shape == (30,)
    assert int(p["n_t"]) == 23821            # T_r = 500 run, saved states
    assert np.all(p["var"] > 0)


def test_replay_reference_lorenz_chain_T2():
    """Stateful G: Phi(u) then Phi(v) every step, IC carried."""
    g = golden("chain_lorenz_pcn_T2.npz")
    p = golden("lorenz_problem_K6_J4.npz")
    op = L.LorenzProblem(6, 4, float(g["T"]), 1, p["prior_means"], p["IC"])
    pot = M.Potential(op, p["y"], 0.5 ** 2 * np.diag(p["var"]))
    out = M.run_chain(pot, g["u0"], g["normals"], g["uniforms"], M.PCN, M.PCN, float(g["beta"]),
                      recompute_phi_u=True)
    assert np.array_equal(out["u"], g["samples"])
    assert np.allclose(out["phi_v"], g["phi_v"], rtol=1e-12)
    assert out["accepts"] == int(g["accepts"])
