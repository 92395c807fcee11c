"""Helpers shared by the -m gpu tests (all calls go through libipmcmc.so)."""
import numpy as np
import torch

import ip_mcmc_b200 as M
from ip_mcmc_b200 import _lib

TRUTH = np.array([0.025, -0.025, -0.02])
PRIOR_MEAN = np.array([1.5, 0.25, -0.5])
PRIOR_COV = 0.25 ** 2 * np.identity(3)
NOISE_COV = 0.05 ** 2 * np.identity(5)


def burgers_setup(N, numerics="exact", y=None, max_fv_steps=0):
    """The reference's Burgers inverse problem (burgers_mcmc.py:22-123) on the device."""
    f = M.BurgersFVM(N=N, numerics=numerics, max_fv_steps=max_fv_steps)
    if y is None:
        y = f.at_parameters(TRUTH)                 # noise-free data G(u*) (burgers_mcmc.py:116)
    noise = M.GaussianDistribution(np.zeros(5), NOISE_COV)
    prior = M.GaussianDistribution(PRIOR_MEAN, PRIOR_COV)
    pot = M.EvolutionPotential(f, y, noise)
    return f, pot, prior, y


def cuda(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def run_injected(pot, spec, u0, normals, uniforms, n_copies=1, recompute=False):
    """Replay a noise tape through the fused kernel; returns (states, steplog, vlog, chains)."""
    n = len(normals)
    d = normals.shape[1]
    ch = M.ChainBatch(pot.problem(), u0, n_chains=n_copies)
    w = cuda(np.stack([normals] * n_copies))
    U = cuda(np.stack([uniforms] * n_copies))
    trace = torch.empty((n_copies, n, d), dtype=torch.float64, device="cuda")
    slog = torch.empty((n_copies, n, 4), dtype=torch.float64, device="cuda")
    vlog = torch.empty((n_copies, n, d), dtype=torch.float64, device="cuda")
    ch.run(spec, n, trace=trace, steplog=slog, vlog=vlog, inject_w=w, inject_u=U)
    return trace.cpu().numpy(), slog.cpu().numpy(), vlog.cpu().numpy(), ch
