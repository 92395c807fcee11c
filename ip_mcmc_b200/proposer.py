"""Proposal kernels with the reference's interface (ip_mcmc/ip_mcmc/proposer.py:8-115).

``__call__(u, rng)`` keeps the reference semantics for a single state (O(d) host arithmetic; it
is what the reference's own unit tests exercise).  Chains are never advanced through it: the
sampler compiles a proposer into ``device_spec`` -- the (coef_u, coef_w) pair or per-step table of
  v = coef_u * u + coef_w * w,   w ~ N(0, C_prior)
that the fused CUDA kernel evaluates with Philox noise.
"""
from abc import ABC, abstractmethod

import numpy as np

from . import _lib
from .distribution import GaussianDistribution


class ProposerBase(ABC):
    @abstractmethod
    def __call__(self, u, rng):
        ...


class _GaussianStepProposer(ProposerBase):
    kind = None

    def _set_prior(self, prior):
        # only the covariance is used; a non-zero prior mean is ignored (proposer.py:26-27,78-79)
        self.w = GaussianDistribution(mean=np.zeros_like(prior.mean), covariance=prior.covariance)

    def _coefs(self, i):
        raise NotImplementedError

    def device_spec(self, n_steps):
        """dict(kind, coef_u, coef_w, schedule) for the next `n_steps` proposals."""
        raise NotImplementedError


class ConstStepStandardRWProposer(_GaussianStepProposer):
    """v = u + sqrt(2 delta) w (proposer.py:14-30)."""
    kind = _lib.PROPOSE_RW

    def __init__(self, delta, prior):
        self.prefactor = np.sqrt(2 * delta)
        self._set_prior(prior)

    def __call__(self, u, rng):
        return u + self.prefactor * self.w.sample(rng)

    def device_spec(self, n_steps):
        return dict(kind=self.kind, coef_u=1.0, coef_w=float(self.prefactor), schedule=None)


class ConstSteppCNProposer(_GaussianStepProposer):
    """v = sqrt(1 - beta^2) u + beta w (proposer.py:59-82)."""
    kind = _lib.PROPOSE_PCN

    def __init__(self, beta, prior):
        assert 0 <= beta <= 1, "beta has to be in [0,1]"
        self.beta = beta
        self.contraction = np.sqrt(1 - beta ** 2)
        self._set_prior(prior)

    def __call__(self, u, rng):
        return self.contraction * u + self.beta * self.w.sample(rng)

    def device_spec(self, n_steps):
        return dict(kind=self.kind, coef_u=float(self.contraction), coef_w=float(self.beta), schedule=None)


class VarStepStandardRWProposer(_GaussianStepProposer):
    """RW with delta = delta(i), i = number of proposals made so far, starting at 1
    (proposer.py:33-56).  Do not reuse an instance across chains."""
    kind = _lib.PROPOSE_RW

    def __init__(self, delta, prior):
        self.prefactor = np.sqrt(2)
        self.delta = delta
        self.i = 0
        self._set_prior(prior)

    def _stepsize(self, i):
        return self.prefactor * np.sqrt(self.delta(i))

    def __call__(self, u, rng):
        self.i += 1
        return u + self._stepsize(self.i) * self.w.sample(rng)

    def device_spec(self, n_steps):
        sched = np.empty((n_steps, 2))
        sched[:, 0] = 1.0
        sched[:, 1] = [self._stepsize(self.i + 1 + s) for s in range(n_steps)]
        self.i += n_steps
        return dict(kind=self.kind, coef_u=1.0, coef_w=0.0, schedule=sched)


class VarSteppCNProposer(_GaussianStepProposer):
    """pCN with beta = beta(i) (proposer.py:85-115)."""
    kind = _lib.PROPOSE_PCN

    def __init__(self, beta, prior):
        self.beta = beta
        self.i = 0
        self._set_prior(prior)

    def __call__(self, u, rng):
        self.i += 1
        b = self.beta(self.i)
        return np.sqrt(1 - b ** 2) * u + b * self.w.sample(rng)

    def device_spec(self, n_steps):
        b = np.array([self.beta(self.i + 1 + s) for s in range(n_steps)], dtype=float)
        assert np.all((0 <= b) & (b <= 1)), "beta has to be in [0,1]"
        self.i += n_steps
        return dict(kind=self.kind, coef_u=0.0, coef_w=0.0,
                    schedule=np.stack([np.sqrt(1 - b ** 2), b], axis=1))
