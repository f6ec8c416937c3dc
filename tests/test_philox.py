"""Philox4x32-10 restatement in the oracle against the Random123 v1.09 known-answer vectors
(kat_vectors: `philox4x32 10 ...`)."""
import numpy as np

from oracle.copter_oracle import philox4x32_10, reset_force

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_random123_known_answers():
    for ctr, key, out in KAT:
        got = philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == out


def test_vectorised_matches_scalar():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2**32, (4, 64), dtype=np.uint64)
    v = philox4x32_10(c[0], c[1], c[2], c[3], 123, 456)
    for i in range(64):
        s = philox4x32_10(*[np.array([c[j, i]]) for j in range(4)], 123, 456)
        assert all(int(s[j][0]) == int(v[j][i]) for j in range(4))


def test_reset_force_range_and_determinism():
    ids = np.arange(100000)
    f = reset_force(7, ids, np.zeros_like(ids), 30.0)
    assert f.shape == (100000, 3) and f.min() >= -30 and f.max() < 30
    assert abs(f.mean()) < 0.2 and abs(f.std() - 60 / np.sqrt(12)) < 0.2
    # the stream of env i does not depend on how many envs are drawn with it
    g = reset_force(7, ids[500:600], np.zeros(100, np.int64), 30.0)
    assert np.array_equal(f[500:600], g)
    # episodes and seeds decorrelate
    assert not np.array_equal(f, reset_force(7, ids, np.ones_like(ids), 30.0))
    assert not np.array_equal(f, reset_force(8, ids, np.zeros_like(ids), 30.0))
    # fp32 draw is the single rounding of the exact fp64 draw
    assert np.array_equal(reset_force(7, ids, np.zeros_like(ids), 30.0, np.float32), f.astype(np.float32))
