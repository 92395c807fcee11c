"""Consumers of the batched engine from the reference's experiment scripts (SURVEY.md section 8(f),
rows 3-4): the `.npy` result cache, the 3-D posterior histogram and the grid-refinement /
chain-length study drivers.  The Wasserstein distance itself needs POT (not installed) and stays out.

  * load_or_compute           report/scripts/helpers.py:23-38 (same semantics, explicit data_dir)
  * histogram3d               np.histogramdd(samples, bins=20, range=...) as used by
                              burgers_wasserstein_grid.py:205-231, on whatever device the samples are
  * grid_refinement_study     burgers_wasserstein_grid.py:165-231 (N in {32,64,128,256}, RW proposals
                              with a PWLinear delta schedule, box constraint on the shock location)
"""
import os

import numpy as np
import torch

from .accepter import BoxConstraint, ConstrainAccepter, CountedAccepter, StandardRWAccepter
from .distribution import GaussianDistribution
from .forward import BurgersFVM
from .potential import EvolutionPotential
from .proposer import VarStepStandardRWProposer
from .sampler import MCMCSampler


def load_or_compute(name, function, args, data_dir="."):
    """np.load(data_dir/name.npy) if it exists, else function(*args), saved there (helpers.py:23-38)."""
    path = os.path.join(data_dir, name + ".npy")
    try:
        res = np.load(path)
    except FileNotFoundError:
        res = function(*args)
        np.save(path, res)
    return res


class PWLinear:
    """Step size decreasing linearly from start to end over `length` proposals, then constant
    (burgers_beta.py:131-147)."""

    def __init__(self, start_delta, end_delta, length):
        self.d_s, self.d_e, self.l = start_delta, end_delta, length
        self.slope = (start_delta - end_delta) / length

    def __call__(self, i):
        if i > self.l:
            return self.d_e
        return self.d_s - self.slope * i

    def __repr__(self):
        return f"pwl_{self.d_s}_{self.d_e}_{self.l}"


def histogram3d(samples, intervals, bins=20, density=True):
    """Normalised 3-D histogram of samples [n, 3] (numpy array or torch tensor on any device) over
    `intervals` [3, 2] -- np.histogramdd(samples, bins=bins, range=intervals, density=True) semantics
    (values on the upper edge fall into the last bin; outside values are dropped)."""
    x = torch.as_tensor(samples, dtype=torch.float64)
    x = x.reshape(-1, x.shape[-1])
    iv = torch.as_tensor(np.asarray(intervals, dtype=np.float64), device=x.device)
    lo, hi = iv[:, 0], iv[:, 1]
    inside = ((x >= lo) & (x <= hi)).all(dim=1)
    x = x[inside]
    idx = torch.floor((x - lo) / (hi - lo) * bins).to(torch.int64).clamp_(0, bins - 1)
    d = x.shape[1]
    flat = idx[:, 0]
    for j in range(1, d):
        flat = flat * bins + idx[:, j]
    h = torch.bincount(flat, minlength=bins ** d).to(torch.float64).reshape((bins,) * d)
    if density:
        vol = torch.prod((hi - lo) / bins)
        h = h / (h.sum() * vol)
    return h


def grid_refinement_study(grids=(32, 64, 128, 256), n_steps=10000, n_chains=64, burn_in=500,
                          prior_mean=(1.5, 0.25, -0.5), prior_std=0.25, noise_std=0.05,
                          truth=(0.025, -0.025, -0.02), schedule=None, seed=2, bins=20,
                          intervals=((-0.5, 0.5), (-0.5, 0.5), (-0.5, 0.5)), numerics="exact"):
    """Posterior histograms of (delta_1, delta_2, sigma) around the truth for several grids: the
    engine-side half of burgers_wasserstein_grid.py (the pairwise W1 distances need POT)."""
    prior_mean = np.asarray(prior_mean, dtype=np.float64)
    truth = np.asarray(truth, dtype=np.float64)
    out = {}
    for N in grids:
        f = BurgersFVM(N=N, prior_means=prior_mean, numerics=numerics)
        y = f.at_parameters(truth)
        prior = GaussianDistribution(prior_mean, prior_std ** 2 * np.identity(3))
        pot = EvolutionPotential(f, y, GaussianDistribution(np.zeros(f.n_obs), noise_std ** 2 * np.identity(f.n_obs)))
        sched = schedule if schedule is not None else PWLinear(0.1, 0.001, burn_in)
        proposer = VarStepStandardRWProposer(sched, prior)
        dom = f.domain
        box = BoxConstraint([-np.inf, -np.inf, dom[0]], [np.inf, np.inf, dom[1]], shift=[0.0, 0.0, prior_mean[2]])
        accepter = CountedAccepter(ConstrainAccepter(StandardRWAccepter(pot, prior), box))
        sampler = MCMCSampler(proposer, accepter, np.random.default_rng(seed))
        samples = sampler.run(np.zeros(3), n_steps - burn_in, burn_in=burn_in + 1, sample_interval=1,
                              n_chains=n_chains, return_device=True)
        centred = samples.reshape(-1, 3) + torch.as_tensor(prior_mean - truth, device=samples.device)
        out[N] = dict(histogram=histogram3d(centred, intervals, bins).cpu().numpy(),
                      acceptance=accepter.ratio(), pooled_mean=sampler.last_run["pooled_mean"] + prior_mean,
                      pooled_var=sampler.last_run["pooled_var"])
    return out
