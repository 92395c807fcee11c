"""Potentials with the reference's interface (ip_mcmc/ip_mcmc/potential.py:6-57)."""
from abc import ABC, abstractmethod

import numpy as np

from .forward import ForwardModel


class PotentialBase(ABC):
    """exp(-potential(u)) is the likelihood L(u; y) (potential.py:6-22)."""

    @abstractmethod
    def __call__(self, u):
        ...

    @abstractmethod
    def exp_minus_potential(self, u):
        ...


class EvolutionPotential(PotentialBase):
    """Phi(u) = -noise.logpdf(data - G(u)) for data = G(u) + eta, eta ~ noise (potential.py:42-57).

    `observation_operator` must be a device forward model (BurgersFVM, Lorenz96Moments); the
    potential is evaluated by the CUDA engine (ipmcmc_forward).  With a stateful operator
    (Lorenz) each call advances the operator's carried initial condition, as in the reference.
    """

    def __init__(self, observation_operator, data, noise_distribution):
        if not isinstance(observation_operator, ForwardModel):
            raise TypeError("observation_operator must be a device forward model (BurgersFVM or "
                            "Lorenz96Moments); arbitrary Python callables have no CUDA implementation and "
                            "there is no CPU fallback")
        self.G = observation_operator
        self.y = np.asarray(data, dtype=np.float64)
        self.rho = noise_distribution
        self._prob = None

    def problem(self):
        from .engine import Problem
        if self._prob is None:
            self._prob = Problem(self.G, self.y, self.rho)
        return self._prob

    def __call__(self, u):
        p = self.problem()
        u = np.asarray(u, dtype=np.float64).reshape(1, -1)
        if self.G.stateful:
            r = p.forward(u, state=self.G.IC.reshape(1, -1))
            self.G.IC = r["state"][0].cpu().numpy()
        else:
            r = p.forward(u)
        return float(r["phi"][0].item())

    def exp_minus_potential(self, u):
        return np.exp(-self(u))

    def batch(self, u, state=None):
        """Phi and G for [n, d] parameter vectors (cuda tensors in a dict)."""
        return self.problem().forward(u, state=state)
