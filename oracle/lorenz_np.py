"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Lorenz-96 forward model.

Restates

  * Lorenz96.__call__ / _slow_variables / _fast_variables   report/scripts/lorenz.py:44-101
  * moment_function                                         report/scripts/lorenz_mcmc.py:17-40
  * LorenzObservationOperator (stateful IC)                 report/scripts/lorenz_mcmc.py:43-71
  * scipy.integrate.solve_ivp(method='RK45') as called at lorenz_mcmc.py:70-71 with all
    defaults: Dormand-Prince 5(4), rtol=1e-3, atol=1e-6, RMS error norm, SAFETY=0.9,
    MIN_FACTOR=0.2, MAX_FACTOR=10, exponent -1/5, factor<=1 after a rejection, last step clipped
    to t_bound, Hairer's initial-step heuristic.  The algorithm lives in the third-party
    dependency scipy (unpinned by the reference, requirements.txt:1-3; scipy 1.18.1 here):
    scipy/integrate/_ivp/rk.py:14-175, common.py:63-140, ivp.py (main loop, every accepted
    step is stored, t0 included).  The restatement uses the same NumPy expressions as scipy so
    it is bit-identical to solve_ivp on this stack (tests/test_oracle_lorenz.py), and it exposes
    the per-attempt quantities (y_new, error norm, step factor) the CUDA kernel is compared with.

Parity note ("Lorenz chaos"): trajectories over T=20 cannot agree to
1e-10 between ANY two implementations that round differently (error growth e^(lambda T));
parity is therefore pinned at the RHS / single RK attempt / controller level and on short
horizons, and statistically on G(u) and chain statistics.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np

# --------------------------------------------------------------------------------------------
# Lorenz-96 right-hand side (lorenz.py:44-101)
# --------------------------------------------------------------------------------------------


def lorenz_rhs(state, K, J, F, h, c, b):
    """d(state)/dt; state = [X_0..X_{K-1}, Y_{0,0..J-1}, ..., Y_{K-1,0..J-1}] (lorenz.py:48).

    slow (lorenz.py:73-88):  -X_k - (X_{k-1} X_{k-2} - X_{k-1} X_{k+1}) + F - (h*c)*mean(Y_k,:)
    fast (lorenz.py:90-101): c * ( -Y_j - b*(Y_{j+1} Y_{j+2} - Y_{j-1} Y_{j+1}) + (h/J)*X_k ),
                             periodic within each k block.
    """
    X = state[:K]
    out = np.empty_like(state)
    Xo = np.copy(X)
    Xo *= -1
    Xo -= np.roll(X, 1) * np.roll(X, 2) - np.roll(X, 1) * np.roll(X, -1)
    Xo += F
    if J:
        Y = np.reshape(state[K:], (K, J))
        fast_slow_fact = h * c
        # np.average over a contiguous row of J<8 elements is a sequential sum / J
        Xo -= fast_slow_fact * (np.add.reduce(Y, axis=1) / J)
        Yo = np.copy(Y)
        Yo *= -1
        Yo -= b * (np.roll(Y, -1, axis=1) * np.roll(Y, -2, axis=1)
                   - np.roll(Y, 1, axis=1) * np.roll(Y, -1, axis=1))
        Yo += (h / J * X)[:, None]
        Yo *= c
        out[K:] = Yo.reshape(-1)
    out[:K] = Xo
    return out


# --------------------------------------------------------------------------------------------
# Dormand-Prince 5(4) exactly as scipy.integrate RK45 (rk.py)
# --------------------------------------------------------------------------------------------
SAFETY = 0.9
MIN_FACTOR = 0.2
MAX_FACTOR = 10
RK45_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
RK45_A = np.array([
    [0, 0, 0, 0, 0],
    [1 / 5, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]])
RK45_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
RK45_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
ERROR_EXPONENT = -1 / 5
RTOL = 1e-3
ATOL = 1e-6


def rms_norm(x):
    """common.py:63-65."""
    return np.linalg.norm(x) / x.size ** 0.5


def select_initial_step(fun, t0, y0, t_bound, f0, order=4, rtol=RTOL, atol=ATOL):
    """common.py:68-140 (direction=+1, max_step=inf)."""
    interval_length = abs(t_bound - t0)
    if interval_length == 0.0:
        return 0.0
    scale = atol + np.abs(y0) * rtol
    d0 = rms_norm(y0 / scale)
    d1 = rms_norm(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = 1e-6
    else:
        h0 = 0.01 * d0 / d1
    h0 = min(h0, interval_length)
    y1 = y0 + h0 * 1 * f0
    f1 = fun(t0 + h0 * 1, y1)
    d2 = rms_norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1 / (order + 1))
    return min(100 * h0, h1, interval_length, np.inf)


def rk45_attempt(fun, t, y, f, h):
    """One Dormand-Prince attempt: rk.py:14-72 (rk_step) + error estimate (rk.py:111-116).
    Returns y_new, f_new, error_norm (RMS of err/scale), K."""
    n = y.shape[0]
    Kst = np.empty((7, n), dtype=np.float64)
    Kst[0] = f
    for s in range(1, 6):
        dy = np.dot(Kst[:s].T, RK45_A[s, :s]) * h
        Kst[s] = fun(t + RK45_C[s] * h, y + dy)
    y_new = y + h * np.dot(Kst[:-1].T, RK45_B)
    f_new = fun(t + h, y_new)
    Kst[-1] = f_new
    scale = ATOL + np.maximum(np.abs(y), np.abs(y_new)) * RTOL
    error_norm = rms_norm(np.dot(Kst.T, RK45_E) * h / scale)
    return y_new, f_new, error_norm, Kst


def step_factor(error_norm, accepted, step_rejected):
    """Step-size controller, rk.py:150-166."""
    if accepted:
        if error_norm == 0:
            factor = MAX_FACTOR
        else:
            factor = min(MAX_FACTOR, SAFETY * error_norm ** ERROR_EXPONENT)
        if step_rejected:
            factor = min(1, factor)
        return factor
    return max(MIN_FACTOR, SAFETY * error_norm ** ERROR_EXPONENT)


def rk45_solve(fun, y0, T, max_attempts=10 ** 7, record_attempts=False):
    """solve_ivp(fun, (0, T), y0, method='RK45') restated.  Returns dict(t, y[n, n_t] incl. t0,
    n_accepted, n_rejected, nfev, attempts=[(t, h, error_norm, accepted)]), status."""
    y = np.array(y0, dtype=np.float64)
    t = 0.0
    t_bound = float(T)
    f = fun(t, y)
    nfev = 1
    h_abs = select_initial_step(fun, t, y, t_bound, f)
    nfev += 1
    ts = [t]
    ys = [y]
    n_acc = n_rej = 0
    attempts = []
    status = 0
    while t != t_bound and status == 0:
        # ---- RungeKutta._step_impl (rk.py:118-175)
        min_step = 10 * np.abs(np.nextafter(t, np.inf) - t)
        if h_abs < min_step:
            h_abs = min_step
        step_accepted = False
        step_rejected = False
        while not step_accepted:
            if h_abs < min_step or n_acc + n_rej >= max_attempts:
                status = -1
                break
            h = h_abs
            t_new = t + h
            if t_new - t_bound > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = np.abs(h)
            y_new, f_new, error_norm, _ = rk45_attempt(fun, t, y, f, h)
            nfev += 6
            if error_norm < 1:
                h_abs *= step_factor(error_norm, True, step_rejected)
                step_accepted = True
                n_acc += 1
            else:
                h_abs *= step_factor(error_norm, False, step_rejected)
                step_rejected = True
                n_rej += 1
            if record_attempts:
                attempts.append((t, h, error_norm, step_accepted))
        if status != 0:
            break
        t = t_new
        y = y_new
        f = f_new
        ts.append(t)
        ys.append(y)
    return dict(t=np.array(ts), y=np.array(ys).T, n_accepted=n_acc, n_rejected=n_rej, nfev=nfev,
                attempts=attempts, status=status)


# --------------------------------------------------------------------------------------------
# moment function and observation operator (lorenz_mcmc.py:17-71)
# --------------------------------------------------------------------------------------------
def moment_function(y, K, J):
    """lorenz_mcmc.py:17-40, vectorised over time.  NOTE the reference takes Y_{k,0} (the first
    fast variable of block k) where the report describes the block mean (lorenz_mcmc.py:32)."""
    n_t = y.shape[1]
    f = np.empty((5 * K, n_t))
    X = y[:K, :]
    Y0 = y[K + J * np.arange(K), :]
    f[:K] = X
    f[K:2 * K] = Y0
    f[2 * K:3 * K] = X ** 2
    f[3 * K:4 * K] = X * Y0
    f[4 * K:5 * K] = Y0 ** 2
    return f


class LorenzProblem:
    """LorenzObservationOperator (lorenz_mcmc.py:43-71): F,h,b = prior_means + u; integrate
    (0,T) from the carried IC; IC <- y[:, -1]; return the unweighted mean of the moment function
    over ALL stored steps (t0 included)."""

    def __init__(self, K, J, T, c, prior_means, IC, use_scipy=False):
        self.K, self.J, self.T, self.c = K, J, T, c
        self.prior_means = np.asarray(prior_means, dtype=np.float64)
        self.IC = np.array(IC, dtype=np.float64)
        self.use_scipy = use_scipy
        self.n_accepted = 0
        self.n_rejected = 0
        self.n_solves = 0
        self.last = None

    def solve(self, u):
        F, h, b = self.prior_means + np.asarray(u, dtype=np.float64)
        K, J, c = self.K, self.J, self.c

        def fun(_, s):
            return lorenz_rhs(s, K, J, F, h, c, b)

        if self.use_scipy:
            from scipy.integrate import solve_ivp
            r = solve_ivp(fun=fun, t_span=(0, self.T), y0=self.IC, method='RK45')
            sol = dict(t=r.t, y=r.y, n_accepted=r.t.size - 1, n_rejected=(r.nfev - 2) // 6 - (r.t.size - 1),
                       nfev=r.nfev)
        else:
            sol = rk45_solve(fun, self.IC, self.T)
        self.n_accepted += sol["n_accepted"]
        self.n_rejected += sol["n_rejected"]
        self.n_solves += 1
        self.last = sol
        return sol

    def G(self, u):
        y = self.solve(u)["y"]
        self.IC = y[:, -1]
        return np.mean(moment_function(y, self.K, self.J), axis=1)

    __call__ = G
