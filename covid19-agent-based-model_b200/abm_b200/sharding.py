"""Household-aligned sharding of the agent store across ranks (one process per GPU).

Agents are cut into `world_size` contiguous ranges of the canonical order, snapped to household
boundaries, so a household never straddles two GPUs and the household term of the infection draw stays
local.  District pools are global: each rank accumulates partial district x status tables over its own
agents (travellers contribute to their destination row from their home rank -- no agent ever migrates)
and one all-reduce per sub-step sums them.  Random draws are keyed by the
global canonical agent id, so trajectories do not depend on the number of ranks.

Two ways to cut (both household-aligned, both give bit-identical trajectories):
  contiguous  rank r holds the r-th of `world_size` ranges of the canonical order (`split_population`,
              `make_synthetic_shard`): a few whole districts per rank;
  dealt       rank r holds the r-th piece of EVERY district (`deal_population`, `make_synthetic_shard_dealt`): every
              rank sees the same mix of districts, hence the same epidemic phase and the same amount of rare-event
              work per sub-step -- less waiting in the per-sub-step table exchange.  The shard is then a sequence of
              id runs; the slot map takes the per-agent global ids (layout.build_slot_map(agent_ids=...)).
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


def deal_population(pop: Population, world_size):
    """Per rank, the indices (into the canonical order) of its dealt shard: piece r of every home district, pieces
    cut at household boundaries near the r/world_size quantiles of the district."""
    D = int(pop.district.max()) + 1 if pop.n else 0
    dstart = np.searchsorted(pop.district, np.arange(D + 1)).astype(np.int64)
    parts = [[] for _ in range(world_size)]
    for d in range(D):
        lo, hi = int(dstart[d]), int(dstart[d + 1])
        if hi == lo:
            continue
        hh = pop.household[lo:hi]
        cuts = [snap_to_household(hh, ((hi - lo) * r) // world_size) for r in range(world_size)] + [hi - lo]
        for r in range(world_size):
            if cuts[r + 1] > cuts[r]:
                parts[r].append(np.arange(lo + cuts[r], lo + cuts[r + 1], dtype=np.int64))
    return [np.concatenate(p) if p else np.zeros(0, dtype=np.int64) for p in parts]


def take_population(pop: Population, idx):
    """The agents `idx` (increasing, whole households) as a Population with dense local household numbers."""
    hh = pop.household[idx]
    dense = np.cumsum(np.r_[0, hh[1:] != hh[:-1]]) if idx.shape[0] else hh
    return Population(district=pop.district[idx], household=dense.astype(np.int64), age=pop.age[idx], status=pop.status[idx],
                      econ_kind=pop.econ_kind[idx], district_mover=pop.district_mover[idx],
                      move_prob_weekday=pop.move_prob_weekday[idx], move_prob_other=pop.move_prob_other[idx],
                      econ_pool=None if pop.econ_pool is None else pop.econ_pool[idx],
                      extras={k: v[idx] for k, v in pop.extras.items()})


def make_synthetic_shard_dealt(n_agents, spec: ModelSpec, seed, world_size, rank):
    """This rank's dealt shard of the synthetic population: (pop_shard, agent_ids, district_start_global).

    Every district is generated from its own RNG stream (the same agents `make_population` builds in one piece), cut
    into `world_size` household-aligned pieces, and piece `rank` is kept -- one district in memory at a time."""
    counts = district_counts(n_agents, spec, seed)
    dstart = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    pieces, ids, hh_next = [], [], 0
    for d in range(spec.n_districts):
        if counts[d] == 0:
            continue
        part = synth.make_population(n_agents, spec, seed=seed, district_range=(d, d + 1))
        n = part.n
        cuts = [snap_to_household(part.household, (n * r) // world_size) for r in range(world_size)] + [n]
        lo, hi = cuts[rank], cuts[rank + 1]
        if hi <= lo:
            continue
        piece = slice_population(part, lo, hi)
        piece.household = piece.household + hh_next
        hh_next = int(piece.household[-1]) + 1
        pieces.append(piece)
        ids.append(np.arange(lo, hi, dtype=np.int64) + int(dstart[d]))
    cat = lambda name: np.concatenate([getattr(p, name) for p in pieces])
    pop = Population(district=cat('district'), household=cat('household'), age=cat('age'), status=cat('status'),
                     econ_kind=cat('econ_kind'), district_mover=cat('district_mover'),
                     move_prob_weekday=cat('move_prob_weekday'), move_prob_other=cat('move_prob_other'),
                     extras={k: np.concatenate([p.extras[k] for p in pieces]) for k in pieces[0].extras})
    return pop, np.concatenate(ids), dstart
