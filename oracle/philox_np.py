"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the engine's counter-based RNG.

The reference draws from numpy's PCG64 (`np.random.default_rng(seed)`, burgers_mcmc.py:101,
lorenz_mcmc.py:83) through ONE shared generator, which makes the draw order part of its
behaviour and cannot be reproduced by thousands of parallel chains.  The engine replaces it by
Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), keyed by
(seed, global chain id) and counted by (step, slot): draws are order-free and invariant to the
GPU count ("Threading").  This file is the bit-exact CPU statement of
that generator and of the integer->double conversions; tests compare the CUDA draws with it.

Pinned against the Random123 known-answer vectors for philox4x32-10 (tests/test_oracle_philox.py).

Layout (must match ip_mcmc_b200/csrc/philox.cuh):
  key     = (seed & 0xffffffff, chain_id & 0xffffffff)
  counter = (step & 0xffffffff, step >> 32, slot, (seed >> 32) & 0xffffffff)
  slot    = component index i for the proposal normal xi_i, SLOT_UNIFORM for the accept uniform
  uniform U  = ((x1<<32 | x0) >> 11) * 2^-53            in [0,1)   (numpy's `random()` mapping)
  normal  xi = sqrt(-2 ln u1) * cos(2 pi u2),  u1 = (((x1<<32|x0) >> 11) + 1) * 2^-53 in (0,1],
                                               u2 =  ((x3<<32|x2) >> 11)      * 2^-53 in [0,1)
"""
import numpy as np

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
SLOT_UNIFORM = 0xFFFFFFFF
MASK32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """counter: 4 python ints (32-bit), key: 2 python ints. Returns 4 ints."""
    c0, c1, c2, c3 = [int(c) & MASK32 for c in counter]
    k0, k1 = [int(k) & MASK32 for k in key]
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK32
        hi1, lo1 = p1 >> 32, p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + PHILOX_W0) & MASK32
        k1 = (k1 + PHILOX_W1) & MASK32
    return c0, c1, c2, c3


def draw_words(seed, chain, step, slot):
    key = (seed & MASK32, chain & MASK32)
    ctr = (step & MASK32, (step >> 32) & MASK32, slot & MASK32, (seed >> 32) & MASK32)
    return philox4x32_10(ctr, key)


def uniform(seed, chain, step):
    x0, x1, _, _ = draw_words(seed, chain, step, SLOT_UNIFORM)
    return np.float64(((x1 << 32 | x0) >> 11)) * np.float64(2.0 ** -53)


def normal(seed, chain, step, slot):
    x0, x1, x2, x3 = draw_words(seed, chain, step, slot)
    u1 = np.float64(((x1 << 32 | x0) >> 11) + 1) * np.float64(2.0 ** -53)
    u2 = np.float64(((x3 << 32 | x2) >> 11)) * np.float64(2.0 ** -53)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def chain_noise(seed, chain, first_step, n_steps, d):
    """(normals[n_steps, d], uniforms[n_steps]) the engine draws for one chain."""
    z = np.empty((n_steps, d))
    U = np.empty(n_steps)
    for s in range(n_steps):
        for i in range(d):
            z[s, i] = normal(seed, chain, first_step + s, i)
        U[s] = uniform(seed, chain, first_step + s)
    return z, U
