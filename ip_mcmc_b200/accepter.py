"""Accept/reject rules with the reference's interface (ip_mcmc/ip_mcmc/accepter.py:6-122).

``__call__(u, v, rng)`` keeps the reference semantics for a single pair of states (the potential
is evaluated by the CUDA forward kernels); the sampler compiles the accepter stack into the
configuration of the fused kernel via ``device_spec``.
"""
from abc import ABC, abstractmethod

import numpy as np

from . import _lib


class AccepterBase(ABC):
    @abstractmethod
    def __call__(self, u, v, rng):
        """Return True if v is accepted"""
        ...


class CountedAccepter(AccepterBase):
    """Counts calls/accepts of the wrapped accepter (accepter.py:13-36)."""

    def __init__(self, accepter):
        self.accepter = accepter
        self.calls = 0
        self.accepts = 0

    def __call__(self, u, v, rng):
        accepted = self.accepter(u, v, rng)
        self.calls += 1
        self.accepts += bool(accepted)
        return accepted

    def reset(self):
        self.calls = 0
        self.accepts = 0

    def ratio(self):
        if self.calls == 0:
            raise ValueError("No samples yet!")
        return self.accepts / self.calls


class BoxConstraint:
    """lo < v + shift < hi componentwise -- the form of every constraint in the reference's
    scripts (burgers_wasserstein_grid.py:48-56) and the one the fused kernel evaluates."""

    def __init__(self, lo, hi, shift=None):
        self.lo = np.asarray(lo, dtype=float)
        self.hi = np.asarray(hi, dtype=float)
        self.shift = np.zeros_like(self.lo) if shift is None else np.asarray(shift, dtype=float)

    def __call__(self, v):
        s = np.asarray(v, dtype=float) + self.shift
        return bool(np.all((s > self.lo) & (s < self.hi)))


class ConstrainAccepter(AccepterBase):
    """Reject without consulting the wrapped accepter -- hence without drawing U -- when the
    constraint is violated (accepter.py:39-55)."""

    def __init__(self, accepter, constraint):
        self.accepter = accepter
        self.is_valid = constraint

    def __call__(self, u, v, rng):
        if self.is_valid(v):
            return self.accepter(u, v, rng)
        return False


class ProbabilisticAccepter(AccepterBase):
    def __call__(self, u, v, rng):
        a = self.accept_probability(u, v)
        return a > rng.random()      # strict, un-clipped, NaN rejects (accepter.py:59-62)

    @abstractmethod
    def accept_probability(self, u, v):
        ...


class StandardRWAccepter(ProbabilisticAccepter):
    """a = exp(I(u) - I(v)), I(w) = Phi(w) + 0.5*||L w||^2 with L the Cholesky factor of the
    prior covariance (accepter.py:86-106; pinned by the reference's accepter_test.py:20-33)."""
    kind = _lib.ACCEPT_RW

    def __init__(self, potential, prior):
        self.theta = potential
        self.prior = prior

    def _I(self, w):
        return self.theta(w) + .5 * np.linalg.norm(self.prior.apply_sqrt_covariance(w)) ** 2

    def accept_probability(self, u, v):
        Iu = self._I(u)
        return np.exp(Iu - self._I(v))


class pCNAccepter(ProbabilisticAccepter):
    """a = exp(Phi(u) - Phi(v)) (accepter.py:109-122)."""
    kind = _lib.ACCEPT_PCN

    def __init__(self, potential):
        self.theta = potential

    def accept_probability(self, u, v):
        pu = self.theta(u)
        return np.exp(pu - self.theta(v))


def device_spec(accepter):
    """Unwrap CountedAccepter / ConstrainAccepter decorators around a StandardRW/pCN accepter.
    Returns dict(kind, potential, prior, constraint, counted), `counted` a list of
    (CountedAccepter, inside_constraint) pairs.  A counter that sits INSIDE the ConstrainAccepter -- the
    reference's Wasserstein scripts use ConstrainAccepter(CountedAccepter(StandardRWAccepter), is_valid),
    burgers_wasserstein_grid.py:150-160 -- never sees a constraint-rejected proposal (accepter.py:52-55),
    so its `calls` exclude them; one that wraps the ConstrainAccepter counts every step."""
    counted = []
    constraint = None
    a = accepter
    while True:
        if isinstance(a, CountedAccepter):
            counted.append((a, constraint is not None))
            a = a.accepter
        elif isinstance(a, ConstrainAccepter):
            if constraint is not None:
                raise TypeError("only one ConstrainAccepter level is supported on the device")
            if not isinstance(a.is_valid, BoxConstraint):
                raise TypeError("the fused kernel evaluates box constraints only: wrap the bounds in "
                                "ip_mcmc_b200.BoxConstraint(lo, hi, shift) (no CPU fallback for "
                                "arbitrary Python predicates)")
            constraint = a.is_valid
            a = a.accepter
        else:
            break
    if not isinstance(a, (StandardRWAccepter, pCNAccepter)):
        raise TypeError("accepter %r has no device implementation (supported: StandardRWAccepter, "
                        "pCNAccepter, optionally wrapped in CountedAccepter/ConstrainAccepter)" % (a,))
    return dict(kind=a.kind, potential=a.theta, prior=getattr(a, "prior", None), constraint=constraint,
                counted=counted, outer_counted=isinstance(accepter, CountedAccepter))


def credit_counters(spec, calls, accepts, constraint_rejects):
    """Add the device counters of a run to every CountedAccepter of the stack (device_spec)."""
    for c, inside in spec["counted"]:
        c.calls += int(calls) - (int(constraint_rejects) if inside else 0)
        c.accepts += int(accepts)
