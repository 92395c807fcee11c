"""Probability primitives with the reference's interface (ip_mcmc/ip_mcmc/distribution.py:8-146).

Only ``GaussianDistribution`` is on the hot path.  It is host-side
set-up code: it produces the small constant tables the CUDA engine needs --

  * ``sample_factor()``  the linear map A with  rng.multivariate_normal(0, C) == A @ z  (numpy's
    SVD method, distribution.py:114-118), used by the device proposal  w = A @ z, z ~ Philox N(0,I);
  * ``whitener()``       (LP, log_pdet, rank) of scipy's multivariate_normal.logpdf
    (distribution.py:111-112), used by the device potential;
  * ``L``                the lower Cholesky factor (distribution.py:104), used by the RW accepter.
"""
from abc import ABC, abstractmethod

import numpy as np
import scipy.linalg as la
from scipy.stats import multivariate_normal


class DistributionBase(ABC):
    """Interface of distribution.py:8-26; subclasses define the dimension attribute ``k``."""

    @abstractmethod
    def sample(self, rng):
        ...

    @abstractmethod
    def __call__(self, x):
        ...

    @abstractmethod
    def logpdf(self, x):
        ...


def _as_array(x, ndim):
    if np.isscalar(x):
        return np.array([x], ndmin=ndim, dtype=float)
    a = np.asarray(x, dtype=float)
    assert a.ndim == ndim, f"Dimension error: {a.ndim} instead of {ndim}."
    return a


class GaussianDistribution(DistributionBase):
    """N(mean, covariance); scalars are promoted to 1-D / 1x1 (distribution.py:94-109)."""

    def __init__(self, mean=0, covariance=1):
        self.mean = _as_array(mean, 1)
        self.covariance = _as_array(covariance, 2)
        self.k = self.mean.shape[0]
        assert self.covariance.shape == (self.k, self.k), "dimension error"
        self.L = np.tril(la.cholesky(self.covariance, lower=True))
        self.dist = multivariate_normal(mean=self.mean, cov=self.covariance)

    def __call__(self, x):
        return self.dist.pdf(x)

    def logpdf(self, x):
        return self.dist.logpdf(x)

    def sample(self, rng):
        return rng.multivariate_normal(mean=self.mean, cov=self.covariance)

    def apply_covariance(self, x):
        return self.covariance @ _as_array(x, 1)

    def apply_sqrt_covariance(self, x):
        return self.L @ _as_array(x, 1)

    def apply_precision(self, x):
        return la.cho_solve((self.L, True), _as_array(x, 1))

    def apply_sqrt_precision(self, x):
        return la.solve_triangular(self.L.T, _as_array(x, 1), lower=False)

    # ---- tables for the CUDA engine -----------------------------------------------------------
    def sample_factor(self):
        """A with multivariate_normal(0, C) = A @ z (numpy Generator, method='svd')."""
        u, s, _ = np.linalg.svd(self.covariance)
        return np.ascontiguousarray(u * np.sqrt(s))

    def whitener(self):
        """(LP, log_pdet, rank): logpdf(x) = -0.5*(rank*log(2 pi) + log_pdet + |(x-mean) @ LP|^2),
        computed as scipy does (eigh, relative eigenvalue cut-off 1e6*eps)."""
        s, u = la.eigh(self.covariance, lower=True)
        eps = 1e6 * np.finfo("d").eps * np.max(np.abs(s))
        if np.min(s) < -eps:
            raise ValueError("covariance is not positive semidefinite")
        keep = s > eps
        if not np.all(keep):
            raise np.linalg.LinAlgError("singular covariance")
        LP = np.multiply(u, np.sqrt(1.0 / s))
        return np.ascontiguousarray(LP), float(np.sum(np.log(s))), int(keep.sum())
