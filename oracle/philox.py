"""Philox4x32-10 in numpy -- TEST INFRASTRUCTURE (oracle), never imported by the product path.

The CUDA kernels draw every random number from Philox4x32-10 (Salmon et al., SC'11,
"Parallel random numbers: as easy as 1, 2, 3") with
    key     = (seed_lo, seed_hi)
    counter = (id_lo, id_hi, step, site)
where id = agent id for the disease-course sites (one block per agent: `agent_draws`) and
id = agent id >> 2 for the per-sub-step sites (one block per four consecutive agents, word
agent id & 3 belongs to the agent: `group_draws`), so a GPU lane that owns four agents
computes one block for all of them.  This module restates the same generator with numpy uint64
arithmetic so the oracle step consumes bit-identical draws (every
reference np.random call site becomes one word of one of these blocks).

Known-answer vectors from the Random123 distribution (kat_vectors, philox4x32 10 rounds)
are checked in tests/test_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
SH32 = np.uint64(32)

# Draw sites (the `site` word of the counter).
SITE_INFECT = 0    # group block: infection draw of the sub-step            (base_model.py:143, 167)
SITE_COURSE_A = 1  # agent block: x0 symptomatic?, x1 incubation / exposed->contagious dwell, x2 hospital?, x3 critical?
SITE_COURSE_B = 2  # agent block: x0 critical fatality, x1 hospital fatality, x2 recovery dwell, x3 spare
SITE_MOVER = 3     # group block: go_out mover draw                         (base_model.py:317)
SITE_OD = 4        # group block: OD destination draw                       (base_model.py:374)
SITE_STAY = 5      # group block: length-of-stay normal                     (base_model.py:431)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable integer arrays (values < 2**32).

    Returns four uint32 arrays.
    """
    c0 = np.asarray(c0).astype(np.uint64) & MASK
    c1 = np.asarray(c1).astype(np.uint64) & MASK
    c2 = np.asarray(c2).astype(np.uint64) & MASK
    c3 = np.asarray(c3).astype(np.uint64) & MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0            # 32x32 -> 64 bit products, exact in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> SH32, p0 & MASK
        hi1, lo1 = p1 >> SH32, p1 & MASK
        c0 = hi1 ^ c1 ^ np.uint64(k0)
        c1 = lo1
        c2 = hi0 ^ c3 ^ np.uint64(k1)
        c3 = lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def agent_draws(seed, step, agent_ids, site):
    """The 4 uint32 lanes for (seed, step, agent, site)."""
    a = np.asarray(agent_ids).astype(np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10(a & MASK, a >> SH32, np.uint64(int(step) & 0xFFFFFFFF), np.uint64(site),
                         seed & 0xFFFFFFFF, seed >> 32)


def group_draws(seed, step, agent_ids, site):
    """One uint32 per agent: word (agent & 3) of the block of group (agent >> 2) at (seed, step, site)."""
    a = np.asarray(agent_ids).astype(np.uint64)
    x = agent_draws(seed, step, a >> np.uint64(2), site)
    w = (a & np.uint64(3)).astype(np.int64)
    return np.choose(w, x) if a.ndim else x[int(w)]
