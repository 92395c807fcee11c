"""Consumers of the batched engine from the reference's experiment scripts (SURVEY.md section 8(f),
rows 3-4): the `.npy` result cache, the 3-D posterior histogram and the grid-refinement /
chain-length study drivers.  The Wasserstein distance itself needs POT (not installed) and stays out.

  * load_or_compute           report/scripts/helpers.py:23-38 (same semantics, explicit data_dir)
  * histogram3d               np.histogramdd(samples, bins=20, range=...) as used by
                              burgers_wasserstein_grid.py:205-231, on whatever device the samples are
  * grid_refinement_study     burgers_wasserstein_grid.py:165-231 (N in {32,64,128,256}, RW proposals
                              with a PWLinear delta schedule, box constraint on the shock location)
  * chain_length_study        burgers_wasserstein_chain.py:164-268 (one 100 000-step run, thinned, split into
                              nested sub-chains of halving length; 20^3 histograms accumulated ON THE DEVICE
                              launch by launch, the samples are never materialised)
  * DeviceHistogram           ipmcmc_histogram_accumulate: np.histogramdd counts kept on the device
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

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


class DeviceHistogram:
    """bins^d int64 counters on the device with np.histogramdd semantics, fed chunk by chunk
    (ipmcmc_histogram_accumulate): np.histogramdd(samples + shift, bins=bins, range=intervals)[0]."""

    def __init__(self, intervals, bins=20, shift=None, device=None):
        iv = np.asarray(intervals, dtype=np.float64)
        self.d, self.bins = iv.shape[0], int(bins)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.edges_host = np.stack([np.linspace(lo, hi, self.bins + 1) for lo, hi in iv])     # numpy's own edges
        self.edges = torch.as_tensor(self.edges_host).to(self.device)
        sh = np.zeros(self.d) if shift is None else np.asarray(shift, dtype=np.float64)
        self.shift = torch.as_tensor(sh).to(self.device)
        self.counts = torch.zeros((self.bins,) * self.d, dtype=torch.int64, device=self.device)

    def add(self, samples):
        """samples: cuda float64 tensor [..., >= d], C-contiguous; every row is one sample."""
        if samples.numel() == 0:
            return
        if not (samples.is_cuda and samples.dtype == torch.float64 and samples.is_contiguous()):
            raise ValueError("samples must be a contiguous float64 cuda tensor")
        stride = samples.shape[-1]
        n = samples.numel() // stride
        _lib.check(_lib.load().ipmcmc_histogram_accumulate(n, self.d, self.bins, C.c_void_p(samples.data_ptr()), stride,
                                                           C.c_void_p(self.shift.data_ptr()), C.c_void_p(self.edges.data_ptr()),
                                                           C.c_void_p(self.counts.data_ptr()),
                                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def normalised(self):
        """counts / counts.sum() -- what the scripts feed to the Wasserstein distance (chain.py:185)."""
        c = self.counts.to(torch.float64)
        return (c / c.sum().clamp(min=1)).cpu().numpy()


def chain_segments(n_kept, n_segments=4):
    """The nested sub-chains of burgers_wasserstein_chain.py:261-266 as [start, stop) ranges over the thinned
    samples: repeatedly l = len // 2, keep [l + 1, len), continue with [0, l)."""
    segs, length = [], int(n_kept)
    for _ in range(n_segments):
        l = length // 2
        segs.append((l + 1, length))
        length = l
    return segs


def chain_length_study(chain_length=100000, n_chains=64, N=128, sample_interval=20, n_segments=4, bins=20,
                       intervals=None, schedule=None, steps_per_launch=5000, prior_mean=(1.5, 0.25, -0.5),
                       prior_std=0.25, noise_std=0.05, truth=(0.025, -0.025, -0.02), seed=2, numerics="exact"):
    """The engine-side half of burgers_wasserstein_chain.py (164-268): ONE run of `chain_length` Metropolis
    steps per chain (VarStep RW with the script's PWLinear(0.05, 0.001, 250) schedule, box-constrained
    accepter with the counter INSIDE the constraint, burgers_wasserstein_chain.py:110-116, 160), thinned by
    `sample_interval` (samples_full[::interval], :255), shifted by the prior mean (:257-259) and split into
    nested sub-chains of halving length (:261-266).  Each sub-chain gets a bins^3 histogram over common
    `intervals`, pooled over the `n_chains` independent chains; the histograms are accumulated on the device
    after every launch of `steps_per_launch` steps, so the run keeps 4 x 20^3 counters instead of
    chain_length x n_chains x 3 samples.

    intervals=None reproduces the script's data-dependent support (min / max over the sub-chains, :170-175)
    with a first pass that only tracks the extrema; the second pass replays the SAME chains (counter-based
    Philox: the chains are a pure function of the seed) and bins them.

    Returns dict(histograms [n_segments, bins, bins, bins] normalised, counts, lengths (thinned samples per
    chain and sub-chain), intervals, acceptance, ground_truth_bin)."""
    prior_mean = np.asarray(prior_mean, dtype=np.float64)
    truth = np.asarray(truth, dtype=np.float64)
    f = BurgersFVM(N=N, prior_means=prior_mean, numerics=numerics)
    y = f.at_parameters(truth)
    prior = GaussianDistribution(prior_mean, prior_std ** 2 * np.identity(3))
    pot = EvolutionPotential(f, y, GaussianDistribution(np.zeros(f.n_obs), noise_std ** 2 * np.identity(f.n_obs)))
    sched = schedule if schedule is not None else PWLinear(0.05, 0.001, 250)
    dom = f.domain
    box = BoxConstraint([-np.inf, -np.inf, dom[0]], [np.inf, np.inf, dom[1]], shift=[0.0, 0.0, prior_mean[2]])
    counted = CountedAccepter(StandardRWAccepter(pot, prior))
    n_kept = len(range(0, chain_length, sample_interval))            # samples_full[::interval]
    segs = chain_segments(n_kept, n_segments)
    dev = pot.problem().device

    def one_pass(hists):
        """Run the chains; per launch either track the extrema of the kept samples (hists None) or bin them."""
        from .engine import ChainBatch, F64
        sampler = MCMCSampler(VarStepStandardRWProposer(sched, prior), ConstrainAccepter(counted, box),
                              np.random.default_rng(seed))
        spec, _, a = sampler._compile(chain_length, 0, 1, None)
        chains = ChainBatch(pot.problem(), np.zeros(3), n_chains=n_chains)
        lo = torch.full((3,), float("inf"), dtype=F64, device=dev)
        hi = torch.full((3,), float("-inf"), dtype=F64, device=dev)
        done = 0
        while done < chain_length:
            n = min(steps_per_launch, chain_length - done)
            trace = torch.empty((n_chains, n, 3), dtype=F64, device=dev)
            chains.run(spec, n, trace=trace)
            # thinned samples of this launch: global step indices done + i with (done + i) % interval == 0
            first = (-done) % sample_interval
            kept = trace[:, first::sample_interval]                         # [n_chains, k, 3]
            k0 = (done + first) // sample_interval                          # index of the first kept sample
            for s, (a0, a1) in enumerate(segs):
                i0, i1 = max(a0 - k0, 0), min(a1 - k0, kept.shape[1])
                if i1 <= i0:
                    continue
                part = kept[:, i0:i1].contiguous()
                if hists is None:
                    flat = part.reshape(-1, 3)
                    lo = torch.minimum(lo, flat.amin(0))
                    hi = torch.maximum(hi, flat.amax(0))
                else:
                    hists[s].add(part)
            done += n
        c = chains.counters.sum(0).cpu().numpy()
        return lo.cpu().numpy(), hi.cpu().numpy(), c

    if intervals is None:
        lo, hi, _ = one_pass(None)
        intervals = np.stack([lo + prior_mean, hi + prior_mean], axis=1)
    intervals = np.asarray(intervals, dtype=np.float64)
    hists = [DeviceHistogram(intervals, bins, shift=prior_mean, device=dev) for _ in segs]
    _, _, c = one_pass(hists)
    gt, _ = np.histogramdd(truth.reshape(1, 3), bins=bins, range=intervals)
    return dict(histograms=np.stack([h.normalised() for h in hists]),
                counts=np.stack([h.counts.cpu().numpy() for h in hists]),
                lengths=[a1 - a0 for a0, a1 in segs], segments=segs, intervals=intervals,
                acceptance=c[1] / max(c[0] - c[5], 1),              # the counter sits inside the constraint
                constraint_rejects=int(c[5]), ground_truth_bin=np.argwhere(gt > 0)[0] if gt.sum() else None)
