"""Household-aligned sharding of the agent store across ranks (one process per GPU).

Agents are cut into `world_size` contiguous ranges of the canonical order, snapped to household
boundaries, so a household never straddles two GPUs and the household term of the infection draw stays
local.  District pools are global: each rank accumulates partial district x status tables over its own
agents (travellers contribute to their destination row from their home rank -- no agent ever migrates)
and one all-reduce per sub-step sums them.  Random draws are keyed by the
global canonical agent id, so trajectories do not depend on the number of ranks.
"""
import numpy as np

from . import synth
from .spec import ModelSpec, Population


def nominal_bounds(n_agents, world_size):
    return [(n_agents * r) // world_size for r in range(world_size + 1)]


def snap_to_household(household, pos):
    """Smallest index >= pos that starts a household (pos may equal len)."""
    n = household.shape[0]
    while 0 < pos < n and household[pos] == household[pos - 1]:
        pos += 1
    return pos


def split_population(pop: Population, world_size):
    """[(lo, hi)] canonical ranges for an in-memory population."""
    nb = nominal_bounds(pop.n, world_size)
    cuts = [snap_to_household(pop.household, p) for p in nb]
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def slice_population(pop: Population, lo, hi):
    sl = slice(lo, hi)
    hh = pop.household[sl]
    return Population(district=pop.district[sl], household=hh - (hh[0] if hi > lo else 0), age=pop.age[sl],
                      status=pop.status[sl], econ_kind=pop.econ_kind[sl], district_mover=pop.district_mover[sl],
                      move_prob_weekday=pop.move_prob_weekday[sl], move_prob_other=pop.move_prob_other[sl],
                      econ_pool=None if pop.econ_pool is None else pop.econ_pool[sl],
                      extras={k: v[sl] for k, v in pop.extras.items()})


def district_counts(n_agents, spec: ModelSpec, seed=0):
    shares = synth.district_shares(spec.n_districts, seed)
    counts = np.floor(shares * n_agents).astype(np.int64)
    counts[np.argmax(counts)] += n_agents - counts.sum()
    return counts


def make_synthetic_shard(n_agents, spec: ModelSpec, seed, world_size, rank):
    """This rank's shard of the synthetic population without materialising the rest.

    Returns (pop_shard, agent_id_base, district_start_global).  Districts are generated from per-district
    RNG streams, so every rank that touches a district sees the same households and the cuts agree."""
    counts = district_counts(n_agents, spec, seed)
    dstart = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nb = nominal_bounds(n_agents, world_size)
    lo_nom, hi_nom = nb[rank], nb[rank + 1]
    d_lo = int(np.searchsorted(dstart, lo_nom, side='right') - 1)
    d_hi = int(np.searchsorted(dstart, max(hi_nom, lo_nom + 1) - 1, side='right'))   # exclusive
    d_hi = min(max(d_hi, d_lo + 1), spec.n_districts)
    while d_hi < spec.n_districts and counts[d_hi - 1] == 0:
        d_hi += 1
    part = synth.make_population(n_agents, spec, seed=seed, district_range=(d_lo, d_hi))
    base = int(dstart[d_lo])
    lo = snap_to_household(part.household, lo_nom - base)
    hi = part.n if rank == world_size - 1 and d_hi == spec.n_districts else snap_to_household(part.household, min(hi_nom - base, part.n))
    return slice_population(part, lo, hi), base + lo, dstart
