"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/oracle_c.c, the plain-C restatement of the
reference's hot path (Burgers FV solve, Lorenz-96 RK45, Gaussian-misfit potential, Metropolis step).

It exists so that the statistical-parity tests can compare the device chains with CPU chains of the
SAME configuration as the benchmark (N = 256, thousands of steps; Lorenz T = 20) in seconds, which
the NumPy restatements (70 chain-steps/s/core, 0.2 for Lorenz) cannot.  Pinned by tests/test_oracle_c.py:
bit-identical to oracle/burgers_np.py / the reference's recorded chains for Burgers, to rounding per
RK attempt against oracle/lorenz_np.py (= scipy) for Lorenz.

Built by build() below into oracle/_build/liboracle_c.so (gcc -O2 -ffp-contract=off -pthread).
Only tests/, __graft_entry__ and bench.py's CPU legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_c.so")

RW, PCN = 0, 1
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return LIB


class _Potential(C.Structure):
    _fields_ = [("q", C.c_int), ("y", _dp), ("perm", _ip), ("scale", _dp), ("log_const", C.c_double)]


class _Burgers(C.Structure):
    _fields_ = [("N", C.c_int), ("max_steps", C.c_long), ("T", C.c_double), ("dx", C.c_double), ("dx_meas", C.c_double),
                ("x", _dp), ("left", _ip), ("right", _ip), ("prior_mean", _dp), ("pot", _Potential)]


class _Lorenz(C.Structure):
    _fields_ = [("K", C.c_int), ("J", C.c_int), ("max_attempts", C.c_long), ("T", C.c_double), ("c", C.c_double),
                ("rtol", C.c_double), ("atol", C.c_double), ("prior_mean", _dp), ("pot", _Potential)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_burgers_solve.restype = C.c_long
        L.orc_burgers_solve.argtypes = [C.POINTER(_Burgers), _dp, _dp, _dp]
        L.orc_burgers_phi.restype = C.c_double
        L.orc_burgers_phi.argtypes = [C.POINTER(_Burgers), _dp, _dp, _dp, C.POINTER(C.c_long)]
        L.orc_lorenz_rhs.restype = None
        L.orc_lorenz_rhs.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.orc_rk45_attempt.restype = C.c_double
        L.orc_rk45_attempt.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.orc_lorenz_phi.restype = C.c_double
        L.orc_lorenz_phi.argtypes = [C.POINTER(_Lorenz), _dp, _dp, _dp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_run_chains.restype = C.c_int
        L.orc_run_chains.argtypes = [C.POINTER(_Burgers), C.POINTER(_Lorenz), C.c_int, C.c_long, C.c_int, C.c_int,
                                     C.c_double, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, C.POINTER(C.c_ubyte),
                                     C.POINTER(C.c_long), C.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _potential(keep, y, noise_cov):
    """Constant tables of the Gaussian misfit for a DIAGONAL noise covariance, from scipy's own
    whitening (mcmc_np.psd_whitener): one non-zero per column of LP."""
    from . import mcmc_np
    LP, log_pdet, rank = mcmc_np.psd_whitener(noise_cov)
    nz = LP != 0
    if not np.all(nz.sum(axis=0) == 1):
        raise ValueError("the C oracle handles diagonal noise covariances only")
    q = LP.shape[0]
    perm = np.ascontiguousarray(np.argmax(nz, axis=0), dtype=np.int32)
    scale = _f64(LP[perm, np.arange(q)])
    yv = _f64(y)
    keep += [perm, scale, yv]
    p = _Potential()
    p.q = q
    p.y = _d(yv)
    p.perm = perm.ctypes.data_as(_ip)
    p.scale = _d(scale)
    p.log_const = float(rank * mcmc_np.LOG_2PI + log_pdet)
    return p


def _chains(bptr, lptr, IC, u0, normals, uniforms, proposer, accepter, step, prior_cov, recompute_phi_u, n_threads):
    normals = _f64(normals)
    uniforms = _f64(uniforms)
    if normals.ndim == 2:
        normals, uniforms = normals[None], uniforms[None]
    n_chains, n_steps, d = normals.shape
    assert d == 3 and uniforms.shape == (n_chains, n_steps)
    u0 = _f64(np.broadcast_to(np.asarray(u0, dtype=np.float64), (n_chains, 3)))
    chol = None
    if accepter == RW:
        import scipy.linalg
        chol = _f64(np.tril(scipy.linalg.cho_factor(np.atleast_2d(prior_cov), lower=True)[0]))
    states = np.empty((n_chains, n_steps, 3))
    phi_v = np.empty((n_chains, n_steps))
    acc = np.empty((n_chains, n_steps), dtype=np.uint8)
    work = np.zeros((n_chains, 2), dtype=np.int64)
    rc = lib().orc_run_chains(bptr, lptr, n_chains, n_steps, proposer, accepter, float(step),
                              _d(chol) if chol is not None else None, _d(u0), _d(IC) if IC is not None else None,
                              _d(normals), _d(uniforms), 1 if recompute_phi_u else 0, _d(states), _d(phi_v),
                              acc.ctypes.data_as(C.POINTER(C.c_ubyte)), work.ctypes.data_as(C.POINTER(C.c_long)),
                              int(n_threads or os.cpu_count() or 1))
    assert rc == 0
    return dict(u=states, phi_v=phi_v, accepted=acc.astype(bool), work=work)


class BurgersC:
    """The reference's Burgers inverse problem (burgers_mcmc.py:22-123) in C; grid tables from
    oracle.burgers_np (the same NumPy calls as the reference)."""

    def __init__(self, n_cells, y=None, noise_cov=None, max_steps=0, **kw):
        from . import burgers_np
        self.P = burgers_np.BurgersProblem(n_cells, **kw)
        self._keep = []
        s = _Burgers()
        s.N = n_cells
        s.max_steps = int(max_steps or 0)
        s.T, s.dx, s.dx_meas = float(self.P.T), float(self.P.dx), float(self.P.dx_meas)
        x = _f64(self.P.x)
        left = np.ascontiguousarray(self.P.left, dtype=np.int32)
        right = np.ascontiguousarray(self.P.right, dtype=np.int32)
        pm = _f64(self.P.prior_mean)
        self._keep += [x, left, right, pm]
        s.x, s.left, s.right, s.prior_mean = _d(x), left.ctypes.data_as(_ip), right.ctypes.data_as(_ip), _d(pm)
        q = len(left)
        s.pot = _potential(self._keep, np.zeros(q) if y is None else y, np.identity(q) if noise_cov is None else noise_cov)
        self.s = s
        self.q = q

    def forward(self, u):
        """dict(G, phi, state, n_fv) for ONE perturbation u (FVMObservationOperator, utilities.py:40-41)."""
        u = _f64(u)
        G, st, n = np.empty(self.q), np.empty(self.s.N), C.c_long()
        phi = lib().orc_burgers_phi(C.byref(self.s), _d(u), _d(G), _d(st), C.byref(n))
        return dict(G=G, phi=phi, state=st, n_fv=n.value)

    def run_chains(self, u0, normals, uniforms, proposer=PCN, accepter=PCN, step=0.25, prior_cov=None,
                   recompute_phi_u=False, n_threads=None):
        return _chains(C.byref(self.s), None, None, u0, normals, uniforms, proposer, accepter, step, prior_cov,
                       recompute_phi_u, n_threads)


class LorenzC:
    """LorenzObservationOperator + EvolutionPotential (lorenz_mcmc.py:43-71, 104-137) in C."""

    def __init__(self, K, J, T, c, prior_means, y=None, noise_cov=None, rtol=1e-3, atol=1e-6, max_attempts=0):
        self._keep = []
        s = _Lorenz()
        s.K, s.J, s.max_attempts = K, J, int(max_attempts or 0)
        s.T, s.c, s.rtol, s.atol = float(T), float(c), rtol, atol
        pm = _f64(prior_means)
        self._keep.append(pm)
        s.prior_mean = _d(pm)
        q = 5 * K
        s.pot = _potential(self._keep, np.zeros(q) if y is None else y, np.identity(q) if noise_cov is None else noise_cov)
        self.s = s
        self.K, self.J, self.nvar, self.q = K, J, K * (J + 1), q

    def rhs(self, theta4, state):
        out = np.empty(self.nvar)
        F, h, c, b = [float(v) for v in theta4]
        lib().orc_lorenz_rhs(self.K, self.J, F, h, c, b, _d(_f64(state)), _d(out))
        return out

    def attempt(self, theta4, y, f, h):
        yn, fn = np.empty(self.nvar), np.empty(self.nvar)
        err = lib().orc_rk45_attempt(self.K, self.J, _d(_f64(theta4)), _d(_f64(y)), _d(_f64(f)), float(h), self.s.rtol,
                                     self.s.atol, _d(yn), _d(fn))
        return yn, fn, err

    def forward(self, u, IC):
        """dict(G, phi, IC (advanced), n_acc, n_rej) for one u from the initial condition IC."""
        ic = np.array(IC, dtype=np.float64)
        G, a, r = np.empty(self.q), C.c_long(), C.c_long()
        phi = lib().orc_lorenz_phi(C.byref(self.s), _d(_f64(u)), _d(ic), _d(G), C.byref(a), C.byref(r))
        return dict(G=G, phi=phi, IC=ic, n_acc=a.value, n_rej=r.value)

    def run_chains(self, u0, IC, normals, uniforms, proposer=RW, accepter=RW, step=0.125, prior_cov=None,
                   n_threads=None):
        n_chains = 1 if np.ndim(normals) == 2 else np.shape(normals)[0]
        ic = _f64(np.broadcast_to(np.asarray(IC, dtype=np.float64), (n_chains, self.nvar))).copy()
        r = _chains(None, C.byref(self.s), ic, u0, normals, uniforms, proposer, accepter, step, prior_cov, True,
                    n_threads)
        r["IC"] = ic
        return r
