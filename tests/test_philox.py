"""Philox4x32-10 known-answer vectors (Random123 kat_vectors, philox4x32 10 rounds)."""
import numpy as np

from oracle.philox import agent_draws, philox4x32_10

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_agent_keying_is_elementwise():
    ids = np.array([0, 1, 2**33 + 5, 123456789], dtype=np.int64)
    all4 = agent_draws(0xDEADBEEFCAFE, 7, ids, 1)
    for i, a in enumerate(ids):
        one = agent_draws(0xDEADBEEFCAFE, 7, np.array([a]), 1)
        assert [int(x[0]) for x in one] == [int(x[i]) for x in all4]
    assert not np.array_equal(agent_draws(1, 7, ids, 0)[0], agent_draws(1, 8, ids, 0)[0])
