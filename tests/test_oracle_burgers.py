"""Pins oracle/burgers_np.py against (i) the reference's in-file known-answer tests
(report/scripts/burgers/rusanov.py:112-170) and (ii) fixtures produced by the live reference
(tests/golden/burgers_forward_N*.npz, oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import golden
from oracle import burgers_np as B
from oracle import mcmc_np as M


def test_grid_kat():
    # rusanov.py:113-122
    x, dx = B.rusanov_grid((-1, 1), 10)
    assert np.isclose(dx, 0.2) and x.shape[0] == 12
    assert np.isclose(x[1], -0.9) and np.isclose(x[-2], 0.9)


def test_cfl_kat():
    # rusanov.py:124-135
    assert np.isclose(B.cfl_dt(np.array([1., 1, 1, 1, 1, 1]), 0.25), 0.125)
    assert np.isclose(B.cfl_dt(np.array([1., 1, -2, 1, 1, 1]), 0.25), 0.0625)


def test_bc_kat():
    # rusanov.py:137-142
    u = np.array([1., 2, 1])
    B.apply_bc(u)
    assert u[0] == 2 and u[2] == 2


def test_flux_kat():
    # rusanov.py:144-153
    f = lambda a, b: B.rusanov_flux(np.float64(a), np.float64(b))
    assert np.isclose(f(1, 1), .5)
    assert np.isclose(f(0, 1), -.25)
    assert np.isclose(f(0, -1), .75)
    assert np.isclose(f(4, 5), 7.75)


def test_rate_of_change_kat():
    # rusanov.py:155-164
    dudt = B.rate_of_change(np.array([1., 1, -1, 2, 2]), 0.3 / 3)
    assert np.allclose(dudt[1:-1], [-10, 32.5, -37.5])
    assert dudt[0] == 0 and dudt[-1] == 0


@pytest.mark.parametrize("N", [32, 64, 100, 128, 200, 256])
def test_forward_bit_identical_to_reference(N):
    g = golden(f"burgers_forward_N{N}.npz")
    P = B.BurgersProblem(N)
    assert np.array_equal(P.x, g["x"]) and P.dx == g["dx"]
    assert np.array_equal(P.left, g["left"]) and np.array_equal(P.right, g["right"])
    assert P.dx_meas == g["dx_meas"]
    pot = M.Potential(P, g["y"], 0.05 ** 2 * np.identity(5))
    for i, u in enumerate(g["u"]):
        end = P.end_state(P.prior_mean + u)
        assert np.array_equal(end, g["end_state"][i])
        assert P.last_n_fv == g["n_fv"][i]
        assert np.array_equal(P.G(u), g["G"][i])
        assert pot(u) == g["phi"][i]


def test_survey_kats():
    # values probed from the reference at survey time
    kat = {128: (1.953125, 1.95279744, 0.80078125, 2131.5932), 256: (2.1484375, 2.14843721, 0.88085938, 2657.0375)}
    for N, (g0, g4, y0, phi0) in kat.items():
        g = golden(f"burgers_forward_N{N}.npz")
        assert np.allclose(g["G"][0][:4], g0) and np.isclose(g["G"][0][4], g4)
        assert np.isclose(g["y"][0], y0) and np.isclose(g["phi"][0], phi0, rtol=1e-7)
    g = golden("burgers_forward_N256.npz")
    assert list(zip(g["left"], g["right"])) == [(58, 70), (90, 102), (154, 166), (186, 198), (218, 230)]
    assert list(g["n_fv"][:2]) == [640, 263]


def test_explicit_trapz_and_pairwise_sum_match_numpy():
    rng = np.random.default_rng(3)
    for n in list(range(1, 140)) + [204, 255, 256, 257, 1000]:
        a = rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)
        assert B.np_pairwise_sum(a) == a.sum()
    trapz = getattr(np, "trapezoid", None) or np.trapz
    for n in (2, 6, 10, 12, 52, 205):
        v = rng.standard_normal(n)
        assert B.trapz_window(v, 0.0078125) == trapz(v, dx=0.0078125)
        terms = 0.0078125 * (v[1:] + v[:-1]) / 2.0
        assert B.np_pairwise_sum(terms) == trapz(v, dx=0.0078125)


def test_large_grid_vector():
    g = golden("burgers_forward_N1024.npz")
    P = B.BurgersProblem(1024)
    u = g["u"][0]
    assert np.array_equal(P.G(u), g["G"][0])
    assert 1045 <= P.last_n_fv <= 1055   # ~max|w|*N FV steps (SURVEY.md section 8(c) quotes 1051)
    assert np.allclose(g["y"], g["G"][0])


def test_max_steps_cap_and_degenerate_state():
    P = B.BurgersProblem(32, max_steps=5)
    P.end_state(P.prior_mean)
    assert P.last_n_fv == 5
    # all-zero state: dt = inf, loop ends after one step with a NaN state (0*inf)
    u, t, n = B.rusanov_integrate(np.zeros(10), 0.1, 1)
    assert n == 1 and np.isinf(t)


@pytest.mark.parametrize("N,m", [(64, 4), (128, 8)])
def test_kl_extension_against_reference_solver_fixture(N, m):
    """KL/spectral extension: IC = Riemann + sum_k a_k sin(k pi (x-a)/(b-a)).  The fixture was produced
    by the reference's RusanovFVM / Measurer / EvolutionPotential driven by such an IC callable."""
    g = golden(f"burgers_kl_N{N}_m{m}.npz")
    P = B.BurgersProblem(N, kl_basis=g["basis"])
    pot = M.Potential(P, g["y"], 0.05 ** 2 * np.identity(5))
    for i, u in enumerate(g["u"]):
        assert np.array_equal(P.end_state(P.prior_mean + u), g["end_state"][i])
        assert np.array_equal(P.G(u), g["G"][i]) and pot(u) == g["phi"][i]
    assert np.array_equal(P.G_params(g["truth"]), g["y"])
