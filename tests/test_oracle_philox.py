"""Pins oracle/philox_np.py to the Random123 known-answer vectors for philox4x32-10
(Random123 kat_vectors: 'philox4x32 10 ...')."""
import numpy as np

from oracle import philox_np as P


def test_random123_kat():
    assert P.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    f = 0xffffffff
    assert P.philox4x32_10((f, f, f, f), (f, f)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert P.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
                           (0xa4093822, 0x299f31d0)) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_uniform_normal_ranges_and_moments():
    z, U = P.chain_noise(seed=2, chain=5, first_step=0, n_steps=4000, d=2)
    assert U.min() >= 0 and U.max() < 1
    assert abs(U.mean() - 0.5) < 0.02
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05
    assert np.all(np.isfinite(z))


def test_streams_are_keyed_by_chain_and_seed():
    a = P.normal(2, 0, 0, 0)
    assert a != P.normal(2, 1, 0, 0) and a != P.normal(3, 0, 0, 0) and a != P.normal(2, 0, 1, 0)
    assert a != P.normal(2 + (1 << 32), 0, 0, 0)
    assert a == P.normal(2, 0, 0, 0)
