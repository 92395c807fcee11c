"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's MCMC operators.

Restates, with the reference's floating-point operation order,

  * GaussianDistribution            ip_mcmc/ip_mcmc/distribution.py:90-146
      .sample  -> numpy Generator.multivariate_normal (method='svd'): w = (U*sqrt(s)) @ z
      .logpdf  -> scipy.stats.multivariate_normal._logpdf (scipy 1.18.1, _multivariate.py:566-591;
                  eigh whitening in _PSD, _multivariate.py:160-190)
      .L       -> scipy.linalg.cho_factor(lower=True)
  * EvolutionPotential.__call__     ip_mcmc/ip_mcmc/potential.py:53-54
  * ConstStepStandardRWProposer     ip_mcmc/ip_mcmc/proposer.py:14-30
  * ConstSteppCNProposer            ip_mcmc/ip_mcmc/proposer.py:59-82
  * VarStep*Proposer                ip_mcmc/ip_mcmc/proposer.py:33-56, 85-115
  * ProbabilisticAccepter           ip_mcmc/ip_mcmc/accepter.py:58-66   (a > U, strict, unclipped)
  * StandardRWAccepter              ip_mcmc/ip_mcmc/accepter.py:86-106  (I = Phi + .5*||L w||^2)
  * pCNAccepter                     ip_mcmc/ip_mcmc/accepter.py:109-122
  * ConstrainAccepter               ip_mcmc/ip_mcmc/accepter.py:39-55
  * MCMCSampler.run/_step           ip_mcmc/ip_mcmc/sampler.py:12-41

with the noise (proposal normals w, accept uniforms U) supplied by the caller -- the same
injection seam as the reference's MockRNG (test_utilities.py:11-26).

Pinned by tests/test_oracle_mcmc.py against the reference's own known-answer tests
(proposer_test.py, accepter_test.py, sampler_test.py, distribution_test.py) and against
chains recorded from the live reference (tests/golden/chain_*.npz).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np

LOG_2PI = np.log(2 * np.pi)   # scipy/stats/_multivariate.py: _LOG_2PI


# --------------------------------------------------------------------------------------------
# Gaussian helpers
# --------------------------------------------------------------------------------------------
def mvn_factor(cov):
    """Matrix A with  Generator.multivariate_normal(0, cov) == A @ z,  z ~ N(0, I)
    (numpy/random/_generator.pyx, method='svd': u, s, vh = svd(cov); x @ (u*sqrt(s)).T)."""
    cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
    u, s, vh = np.linalg.svd(cov)
    return u * np.sqrt(s)


def psd_whitener(cov):
    """(LP, log_pdet, rank) as scipy's _PSD computes them (_multivariate.py:160-190):
    s, u = eigh(cov, lower=True); eps = 1e6*dbl_eps*max|s|; LP = u * sqrt(1/s); logdet = sum log s."""
    import scipy.linalg
    cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
    s, u = scipy.linalg.eigh(cov, lower=True, check_finite=True)
    eps = 1e6 * np.finfo('d').eps * np.max(np.abs(s))
    d = s[s > eps]
    s_pinv = np.array([0 if abs(x) <= eps else 1 / x for x in s], dtype=float)
    LP = np.multiply(u, np.sqrt(s_pinv))
    return LP, np.sum(np.log(d)), len(d)


def gaussian_logpdf(dev, LP, log_pdet, rank):
    """scipy _logpdf: -0.5*(rank*LOG_2PI + log_pdet + sum(square(dev @ LP)))."""
    maha = np.sum(np.square(dev @ LP), axis=-1)
    return -0.5 * (rank * LOG_2PI + log_pdet + maha)


class Potential:
    """EvolutionPotential (potential.py:42-57): Phi(u) = -logpdf_noise(y - G(u))."""

    def __init__(self, G, y, noise_cov):
        self.G = G
        self.y = np.asarray(y, dtype=np.float64)
        self.LP, self.log_pdet, self.rank = psd_whitener(noise_cov)
        self.n_calls = 0

    def __call__(self, u):
        self.n_calls += 1
        return -gaussian_logpdf(self.y - self.G(u), self.LP, self.log_pdet, self.rank)


def prior_regulariser(w, L):
    """accepter.py:104-106: .5 * np.linalg.norm(L @ w)**2  (L = lower Cholesky of the prior
    covariance -- NOT its inverse; pinned by accepter_test.py:29-30)."""
    return .5 * np.linalg.norm(L @ w) ** 2


# --------------------------------------------------------------------------------------------
# one Metropolis step with injected noise
# --------------------------------------------------------------------------------------------
RW, PCN = 0, 1


def propose(kind, u, w, step):
    """kind RW:  v = u + sqrt(2*delta)*w   (proposer.py:23,29-30; step = delta)
       kind PCN: v = sqrt(1-beta^2)*u + beta*w  (proposer.py:77,81-82; step = beta)"""
    if kind == RW:
        return u + np.sqrt(2 * step) * w
    return np.sqrt(1 - step ** 2) * u + step * w


def propose_varstep(kind, u, w, step):
    """VarStep variants evaluate the same formulas with slightly different op order
    (proposer.py:53-56: sqrt(2)*sqrt(delta_i); proposer.py:110-115: sqrt(1-b**2))."""
    if kind == RW:
        return u + (np.sqrt(2) * np.sqrt(step)) * w
    return np.sqrt(1 - step ** 2) * u + step * w


def accept_probability(kind, phi_u, phi_v, u=None, v=None, L=None):
    """RW : exp(I(u)-I(v)), I(w)=Phi(w)+.5||L w||^2 (accepter.py:98-106)
       PCN: exp(Phi(u)-Phi(v))                      (accepter.py:121-122)."""
    with np.errstate(over="ignore", invalid="ignore"):
        if kind == RW:
            return np.exp((phi_u + prior_regulariser(u, L)) - (phi_v + prior_regulariser(v, L)))
        return np.exp(phi_u - phi_v)


def run_chain(potential, u0, normals, uniforms, proposer=PCN, accepter=PCN, step=0.25,
              prior_cov=None, constraint=None, varstep=False, recompute_phi_u=False, uniforms_by_step=False):
    """Replay MCMCSampler._step (sampler.py:35-41) len(uniforms)-or-len(normals) times with
    injected noise.  `step` is a scalar (Const* proposers) or an array indexed by step
    (VarStep*, already evaluated at i = 1, 2, ...).

    recompute_phi_u=True evaluates Phi(u) again every step, in the reference's order
    (Phi(u) first, then Phi(v); accepter.py:99-100,121-122) -- required for stateful forward
    models (Lorenz); for a deterministic G the cached value is bit-identical.

    Returns dict with per-step arrays: v, phi_v, phi_u, a, accepted, u (state AFTER the step),
    and the counters of a CountedAccepter wrapped OUTSIDE a ConstrainAccepter
    (calls counts every step, accepts only real accepts).
    """
    import scipy.linalg
    u = np.array(u0, dtype=np.float64)
    n = len(normals)
    d = u.shape[0]
    L = None
    if accepter == RW:
        L, _ = scipy.linalg.cho_factor(np.atleast_2d(prior_cov), lower=True)
        L = np.tril(L)
    out = dict(v=np.empty((n, d)), phi_v=np.full(n, np.nan), phi_u=np.full(n, np.nan),
               a=np.full(n, np.nan), accepted=np.zeros(n, dtype=bool), u=np.empty((n, d)))
    phi_u = None
    iu = 0
    for i in range(n):
        st = step[i] if np.ndim(step) else step
        v = (propose_varstep if varstep else propose)(proposer, u, normals[i], st)
        out["v"][i] = v
        ok = True
        if constraint is not None and not constraint(v):
            ok = False                       # accepter.py:52-55: no U drawn
        if ok:
            if phi_u is None or recompute_phi_u:
                phi_u = potential(u)
            phi_v = potential(v)
            a = accept_probability(accepter, phi_u, phi_v, u, v, L)
            # the reference draws U only when the constraint holds (accepter.py:52-55): a tape is consumed in
            # order; the engine's counter-based RNG indexes U by the step number instead (uniforms_by_step)
            U = uniforms[i] if uniforms_by_step else uniforms[iu]
            iu += 1
            acc = bool(a > U)                # accepter.py:61-62 (NaN -> False)
            out["phi_v"][i] = phi_v
            out["phi_u"][i] = phi_u
            out["a"][i] = a
            out["accepted"][i] = acc
            if acc:
                u = v
                phi_u = phi_v
        out["u"][i] = u
    out["calls"] = n
    out["accepts"] = int(out["accepted"].sum())
    out["uniforms_used"] = iu
    return out


def sampler_total_steps(n_samples, burn_in, sample_interval):
    """sampler.py:18-26 (pinned by sampler_test.py:14-16)."""
    return max(0, burn_in - sample_interval) + n_samples * sample_interval


def samples_from_states(states, n_samples, burn_in, sample_interval):
    """Rows of `states` (state after every step) that MCMCSampler.run records (sampler.py:23-28)."""
    pre = max(0, burn_in - sample_interval)
    idx = pre + sample_interval * (1 + np.arange(n_samples)) - 1
    return states[idx]


# --------------------------------------------------------------------------------------------
# chain statistics (sampler.py:43-54 and the ESS definition of SURVEY.md section 8(d))
# --------------------------------------------------------------------------------------------
def autocorr(x):
    """MCMCSampler.autocorr (sampler.py:43-54)."""
    x_ = x - np.mean(x)
    result = np.correlate(x_, x_, mode='full')
    result = result[-len(x):]
    if result[0] == 0:
        return np.ones_like(result)
    return result / result[0]
