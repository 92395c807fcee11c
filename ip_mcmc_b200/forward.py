"""Device forward-model descriptors, accepted as ``observation_operator`` by EvolutionPotential.

In the reference the observation operator is an arbitrary Python callable (potential.py:48-49);
the two the north star names live in the report scripts.  Here they are descriptors of CUDA
kernels; calling one evaluates G(u) on the GPU.  Arbitrary Python callables are not supported
(no CPU fallback) -- EvolutionPotential raises TypeError for them.
"""
import ctypes as C

import numpy as np

from . import _lib


class ForwardModel:
    kind = None
    n_params = None
    n_obs = None
    stateful = False

    def _problem(self):
        from .engine import Problem
        if getattr(self, "_bare_problem", None) is None:
            self._bare_problem = Problem(self)
        return self._bare_problem

    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop("_bare_problem", None)
        st.pop("_abs_model", None)
        return st


class BurgersFVM(ForwardModel):
    """G(u) = Measurer(RusanovFVM(Burgers flux).integrate(PerturbedRiemannIC(prior_means + u), T)).

    Replaces FVMObservationOperator(PerturbedRiemannIC, prior_means, RusanovMCMC(flux, flux',
    domain, N, T), Measurer(points, interval, x))  (report/scripts/burgers/utilities.py:17-109,
    rusanov.py:15-109; wiring burgers_mcmc.py:86-114).  The grid tables are produced with the same
    NumPy calls as the reference (linspace with retstep: rusanov.py:18-25; searchsorted:
    utilities.py:93-98) so that they are bit-identical.

    numerics: "exact" (reference rounding order, bit-identical results) or "fused"
    (FMA-contracted update, 1e-13 ... 3e-11 relative agreement in G depending on the grid, fewer fp64 instructions).
    """
    kind = _lib.MODEL_BURGERS

    def __init__(self, domain=(-1, 1), N=200, T=1, prior_means=(1.5, 0.25, -0.5),
                 points=(-0.5, -0.25, 0.25, 0.5, 0.75), interval=0.1, numerics="exact", max_fv_steps=0,
                 kl_modes=0, monotone_shortcut=True):
        self.domain = (float(domain[0]), float(domain[1]))
        self.N = int(N)
        self.T = float(T)
        self.kl_modes = int(kl_modes)
        self.n_params = 3 + self.kl_modes
        pm = np.asarray(prior_means, dtype=np.float64)
        if pm.shape == (3,) and self.kl_modes:
            pm = np.concatenate([pm, np.zeros(self.kl_modes)])      # KL coefficients are centred
        self.prior_means = np.ascontiguousarray(pm)
        assert self.prior_means.shape == (self.n_params,), \
            "PerturbedRiemannIC takes (delta_1, delta_2, sigma) [+ kl_modes coefficients]"
        self.points = np.asarray(points, dtype=np.float64)
        self.interval = float(interval)
        if numerics not in ("exact", "fused"):
            raise ValueError("numerics must be 'exact' or 'fused'")
        self.numerics = numerics
        self.max_fv_steps = int(max_fv_steps)       # 0: the library default, see effective_max_fv_steps
        # FUSED numerics: take max|u| of a monotone state from its two end cells (bit-identical results; the
        # switch exists so that the tests can prove it)
        self.monotone_shortcut = bool(monotone_shortcut)
        a, b = self.domain
        dx0 = (b - a) / self.N
        self.x, self.dx = np.linspace(start=a - .5 * dx0, stop=b + .5 * dx0, num=self.N + 2, retstep=True)
        self.x = np.ascontiguousarray(self.x)
        xv = self.x[1:-1]
        self.dx_meas = xv[1] - xv[0]
        self.left_limits = np.searchsorted(xv, self.points - self.interval / 2, side="left").astype(np.int32)
        self.right_limits = np.searchsorted(xv, self.points + self.interval / 2, side="left").astype(np.int32)
        self.n_obs = self.points.shape[0]
        # KL / spectral extension (north star): w0(x) = Riemann(x) + sum_k a_k sin(k pi (x-a)/(b-a)),
        # k = 1..kl_modes, sampled at the N+2 cell centres like the reference samples its IC (rusanov.py:32)
        self.kl_basis = self.sine_basis(self.x, self.domain, self.kl_modes)

    @property
    def effective_max_fv_steps(self):
        """Cap on FV time steps per solve (include/ipmcmc.h: max_fv_steps; default 8 N + 256).  A capped
        solve reports Phi = NaN, i.e. the proposal is rejected and counted in `nonfinite`."""
        return self.max_fv_steps if self.max_fv_steps > 0 else 8 * self.N + 256

    @staticmethod
    def sine_basis(x, domain, m):
        a, b = domain
        k = np.arange(1, m + 1, dtype=np.float64)[:, None]
        return np.ascontiguousarray(np.sin(k * np.pi * (np.asarray(x)[None, :] - a) / (b - a)))

    @staticmethod
    def kl_prior_variances(m, scale=0.1, decay=2.0):
        """lambda_k = scale^2 * k^-decay: a power-law KL spectrum for the coefficients a_k."""
        return scale ** 2 * np.arange(1, m + 1, dtype=np.float64) ** (-decay)

    def _c_desc(self, keep):
        d = _lib.BurgersDesc()
        d.n_cells = self.N
        d.numerics = _lib.NUMERICS_FUSED if self.numerics == "fused" else _lib.NUMERICS_EXACT
        d.max_fv_steps = self.max_fv_steps
        d.n_params = self.n_params
        d.n_kl_modes = self.kl_modes
        d.flags = 0 if self.monotone_shortcut else _lib.BURGERS_NO_MONOTONE_SHORTCUT
        if self.kl_modes:
            d.kl_basis = _lib.as_double_p(self.kl_basis)
            keep.append(self.kl_basis)
        d.T, d.dx, d.dx_meas = self.T, float(self.dx), float(self.dx_meas)
        d.x = _lib.as_double_p(self.x)
        d.param_mean = _lib.as_double_p(self.prior_means)
        d.win_left = _lib.as_int32_p(self.left_limits)
        d.win_right = _lib.as_int32_p(self.right_limits)
        keep += [self.x, self.prior_means, self.left_limits, self.right_limits]
        return d

    def __call__(self, u):
        """G(u) for one parameter vector (utilities.py:40-41) -> ndarray[q]."""
        return self._problem().forward(np.asarray(u, dtype=np.float64).reshape(1, self.n_params))["G"][0].cpu().numpy()

    def batch(self, u, want_state=False):
        """G for [n, 3] parameter vectors; dict(G, work=(FV steps, 0), state=end states)."""
        return self._problem().forward(u, want_state=want_state)

    def at_parameters(self, params):
        """measurer(integrator(PerturbedRiemannIC(params))) for ABSOLUTE parameters -- how the
        reference generates its noise-free data, burgers_mcmc.py:104,116 (prior mean + (params -
        prior mean) would round differently)."""
        if getattr(self, "_abs_model", None) is None:
            self._abs_model = BurgersFVM(self.domain, self.N, self.T, np.zeros(self.n_params), self.points,
                                         self.interval, self.numerics, self.max_fv_steps, self.kl_modes,
                                         self.monotone_shortcut)
        return self._abs_model(params)


class Lorenz96Moments(ForwardModel):
    """G(u) = time-mean over one solve_ivp(RK45) run of length T of the 5K moment functions of the
    two-scale Lorenz-96 system with (F, h, b) = prior_means + u and fixed c.

    Replaces LorenzObservationOperator(K, J, T, c, prior_means, IC) (report/scripts/lorenz_mcmc.py:
    43-71) with Lorenz96 (lorenz.py:13-101) and moment_function (lorenz_mcmc.py:17-40).  Like the
    reference it is STATEFUL: every evaluation starts from the end state of the previous one
    (lorenz_mcmc.py:66); ``IC`` holds that state for single-vector calls, batched chains carry one
    state per chain on the device.
    """
    kind = _lib.MODEL_LORENZ
    n_params = 3
    stateful = True

    def __init__(self, K, J, T, c, prior_means, IC, rtol=1e-3, atol=1e-6, max_attempts=0, numerics="exact"):
        if numerics not in ("exact", "fused"):
            raise ValueError("numerics must be 'exact' or 'fused'")
        self.numerics = numerics
        self.K, self.J = int(K), int(J)
        self.T = float(T)
        self.c = float(c)
        self.prior_means = np.ascontiguousarray(prior_means, dtype=np.float64)
        assert self.prior_means.shape == (3,), "the operator infers (F, h, b)"
        self.n_var = self.K * (self.J + 1)
        self.IC = np.array(IC, dtype=np.float64)
        assert self.IC.shape == (self.n_var,), "IC must have K*(J+1) entries"
        self.rtol, self.atol = float(rtol), float(atol)
        self.max_attempts = int(max_attempts)
        self.n_obs = 5 * self.K

    def _c_desc(self, keep):
        d = _lib.LorenzDesc()
        d.K, d.J = self.K, self.J
        d.max_attempts = self.max_attempts
        d.numerics = _lib.NUMERICS_FUSED if self.numerics == "fused" else _lib.NUMERICS_EXACT
        d.T, d.c, d.rtol, d.atol = self.T, self.c, self.rtol, self.atol
        d.param_mean = _lib.as_double_p(self.prior_means)
        keep += [self.prior_means]
        return d

    def __call__(self, u):
        r = self._problem().forward(np.asarray(u, dtype=np.float64).reshape(1, 3), state=self.IC.reshape(1, -1))
        self.IC = r["state"][0].cpu().numpy()
        return r["G"][0].cpu().numpy()

    def batch(self, u, state):
        return self._problem().forward(u, state=state)
