"""Structure-of-arrays agent store: host-side packing / unpacking.

Device layout (one slot per agent, slots grouped in tiles of TILE = 1024):
  hot, read every sub-step (8 B/agent):
    flags  uint32   bit fields of constants.F_*  (epidemic / clinical / symptomatic state, location kind,
                    away, district_mover, economic status, age, household-head bit, outcome bits, ...)
    aux    uint32   low 16: sub-step index at which the next scheduled event fires (FIRE_NEVER = none)
                    high 16: current district id
  cold, touched only when an event fires, on infection, or on the two movement sub-steps for travellers:
    t_infected, t_contagious, t_onset, t_recovered, t_return   int32 minute ticks
    inf_at uint32   (location kind << 16) | district at infection
    move_class uint16  index into the movement-probability threshold table
    econ_pool  uint32  outside-pool row for EKIND_POOL agents (schools), allocated only when used
Agents are in canonical order (home district, household).  Filler slots are appended at the end of a
tile whenever the next household (of <= BIG_HOUSEHOLD members) would straddle a tile boundary, so a
thread block always sees whole households; households larger than that are flagged F_BIGHH and use a
global-atomic path.  `tile_base[t]` is the canonical id of the first agent of tile t.

This mirrors what Country.initialize_agent_vectors / initialize_epidemic_vectors build
(/root/reference/src/covid19_abm/base_model.py:527-634) with 31 eight-byte vectors per agent.
"""
from dataclasses import dataclass

import numpy as np

from . import constants as C
from .spec import Population

T_NEVER = int(C.T_NEVER)


@dataclass
class SlotMap:
    n_agents: int
    n_tiles: int
    tile_base: np.ndarray      # int64 [n_tiles + 1] canonical (global) id of first agent per tile
    slot_of: np.ndarray        # int64 [n_agents] local agent -> slot
    agent_of: np.ndarray       # int64 [n_slots] slot -> local agent, -1 for fillers
    big_start: np.ndarray      # int64 [n_big + 1] global id of first member of each big household (+ sentinel)
    big_size: np.ndarray       # int64 [n_big]
    agent_id_base: int

    @property
    def n_slots(self):
        return self.n_tiles * C.TILE


def build_slot_map(household: np.ndarray, agent_id_base: int = 0) -> SlotMap:
    """Greedy tile packing.  Units are whole small households and single members of big ones."""
    n = int(household.shape[0])
    if n == 0:
        raise ValueError('empty population')
    head = np.empty(n, dtype=bool)
    head[0] = True
    head[1:] = household[1:] != household[:-1]
    starts = np.nonzero(head)[0]
    sizes = np.diff(np.append(starts, n))
    big = sizes > C.BIG_HOUSEHOLD
    # placement units
    unit_size = np.where(big, 1, sizes)
    unit_rep = np.where(big, sizes, 1)
    unit_len = np.repeat(unit_size, unit_rep)                 # length of each unit
    unit_end = np.cumsum(unit_len)                            # agents consumed after each unit
    tile_first = [0]                                          # first agent of each tile
    pos = 0
    while pos < n:
        # last unit whose end <= pos + TILE
        j = int(np.searchsorted(unit_end, pos + C.TILE, side='right'))
        nxt = int(unit_end[j - 1]) if j > 0 else pos
        if nxt <= pos:
            raise RuntimeError('placement unit larger than a tile')
        pos = nxt
        tile_first.append(pos)
    tile_first = np.array(tile_first, dtype=np.int64)
    n_tiles = tile_first.shape[0] - 1
    tile_of_agent = np.searchsorted(tile_first, np.arange(n, dtype=np.int64), side='right') - 1
    slot_of = tile_of_agent * C.TILE + (np.arange(n, dtype=np.int64) - tile_first[tile_of_agent])
    agent_of = np.full(n_tiles * C.TILE, -1, dtype=np.int64)
    agent_of[slot_of] = np.arange(n, dtype=np.int64)
    big_start = np.append(starts[big], n).astype(np.int64) + agent_id_base
    return SlotMap(n_agents=n, n_tiles=n_tiles, tile_base=tile_first + agent_id_base, slot_of=slot_of,
                   agent_of=agent_of, big_start=big_start, big_size=sizes[big].astype(np.int64),
                   agent_id_base=int(agent_id_base))


def pack_flags(pop: Population, sm: SlotMap):
    """Initial hot record (everyone susceptible, at home): flags / aux arrays over slots."""
    n = pop.n
    head = np.empty(n, dtype=bool)
    head[0] = True
    head[1:] = pop.household[1:] != pop.household[:-1]
    starts = np.nonzero(head)[0]
    sizes = np.diff(np.append(starts, n))
    big_agent = np.repeat(sizes > C.BIG_HOUSEHOLD, sizes)
    f = np.zeros(n, dtype=np.uint32)
    f |= np.uint32(C.LOC_HOME) << np.uint32(C.F_LOC_SHIFT)
    f |= (pop.district_mover.astype(np.uint32) & 1) * np.uint32(C.F_DMOVER)
    f |= pop.status.astype(np.uint32) << np.uint32(C.F_STATUS_SHIFT)
    f |= pop.age.astype(np.uint32) << np.uint32(C.F_AGE_SHIFT)
    f |= head.astype(np.uint32) * np.uint32(C.F_HEAD)
    f |= pop.econ_kind.astype(np.uint32) << np.uint32(C.F_EKIND_SHIFT)
    f |= big_agent.astype(np.uint32) * np.uint32(C.F_BIGHH)
    filler = np.uint32(C.F_FILLER | C.F_HEAD | (C.LOC_DEAD << C.F_LOC_SHIFT) | (C.STATE_DEAD << C.F_EPI_SHIFT)
                       | (C.CLINICAL_RELEASED_OR_DEAD << C.F_CLIN_SHIFT))
    flags = np.full(sm.n_slots, filler, dtype=np.uint32)
    flags[sm.slot_of] = f
    aux = np.full(sm.n_slots, C.FIRE_NEVER, dtype=np.uint32)
    aux[sm.slot_of] = (pop.district.astype(np.uint32) << np.uint32(16)) | np.uint32(C.FIRE_NEVER)
    return flags, aux


def update_static_fields(flags: np.ndarray, pop: Population, sm: SlotMap):
    """Re-encode the fields scenario hooks mutate on the host (district_mover, econ-location kind)
    into an existing flags array (base_model.py:642-731 helpers), keeping all dynamic state."""
    keep = ~np.uint32(C.F_DMOVER | (C.F_EKIND_MASK << C.F_EKIND_SHIFT))
    f = flags[sm.slot_of] & keep
    f |= (pop.district_mover.astype(np.uint32) & 1) * np.uint32(C.F_DMOVER)
    f |= pop.econ_kind.astype(np.uint32) << np.uint32(C.F_EKIND_SHIFT)
    flags[sm.slot_of] = f
    return flags


def unpack_state(dev: dict, pop: Population, sm: SlotMap, n_districts: int, n_households_before: int = 0):
    """Device arrays (numpy copies over slots) -> reference-style vectors over local agents, in the
    representation of oracle.abm_oracle.OracleModel.state_dict (ticks; T_NEVER = datetime.max)."""
    s = sm.slot_of
    f = dev['flags'][s]
    aux = dev['aux'][s]
    D = n_districts
    epi = ((f >> C.F_EPI_SHIFT) & C.F_EPI_MASK).astype(np.int8)
    clin = ((f >> C.F_CLIN_SHIFT) & C.F_CLIN_MASK).astype(np.int8)
    symc = ((f >> C.F_SYM_SHIFT) & C.F_SYM_MASK).astype(np.int8)
    kind = ((f >> C.F_LOC_SHIFT) & C.F_LOC_MASK).astype(np.int64)
    cur = (aux >> 16).astype(np.int64)
    household_ids = pop.household.astype(np.int64) + D
    H = pop.n_households
    econ_loc = np.where(pop.econ_kind == C.EKIND_HOUSEHOLD, household_ids, pop.district.astype(np.int64))
    if pop.econ_pool is not None:
        econ_loc = np.where(pop.econ_kind == C.EKIND_POOL, D + H + (pop.econ_pool.astype(np.int64) - D), econ_loc)

    def loc_from(kind, cur):
        return np.select([kind == C.LOC_HOME, kind == C.LOC_ECON, kind == C.LOC_DIST, kind == C.LOC_HOSP],
                         [household_ids, econ_loc, cur, -cur], default=-1)

    away = (f & C.F_AWAY) != 0
    infected = dev['t_infected'][s] != T_NEVER
    onset = dev['t_onset'][s].astype(np.int64)
    has_h = (f & C.F_OUT_H) != 0
    has_c = (f & C.F_OUT_C) != 0
    has_d = (f & C.F_OUT_D) != 0
    never = np.int64(T_NEVER)
    inf_at = dev['inf_at'][s]
    out = dict(
        epidemic_state=epi, clinical_state=clin, infected_symptomatic_status=(symc - 1).astype(np.int8),
        current_location_ids=loc_from(kind, cur), current_district_ids=cur,
        return_district_at=np.where(away, dev['t_return'][s], T_NEVER).astype(np.int32),
        date_infected=dev['t_infected'][s].astype(np.int32),
        date_start_contagious=dev['t_contagious'][s].astype(np.int32),
        date_start_symptomatic=dev['t_onset'][s].astype(np.int32),
        date_recovered=dev['t_recovered'][s].astype(np.int32),
        date_start_hospitalized=np.where(has_h, onset + C.ONSET_TO_HOSPITAL_TICKS, never).astype(np.int32),
        date_end_hospitalized=np.where(has_h, onset + C.ONSET_TO_HOSPITAL_TICKS + C.HOSPITAL_TICKS, never).astype(np.int32),
        date_start_critical=np.where(has_c, onset + C.ONSET_TO_HOSPITAL_TICKS + C.HOSPITAL_TICKS, never).astype(np.int32),
        date_end_critical=np.where(has_c, onset + C.ONSET_TO_HOSPITAL_TICKS + C.HOSPITAL_TICKS + C.CRITICAL_TICKS,
                                   never).astype(np.int32),
        date_died=np.where(has_d, onset + C.ONSET_TO_DEATH_TICKS, never).astype(np.int32),
        infected_at_district_ids=np.where(infected, (inf_at & 0xFFFF).astype(np.int64), -1),
        infected_at_location_ids=np.where(infected, loc_from((inf_at >> 16).astype(np.int64) & 7,
                                                             (inf_at & 0xFFFF).astype(np.int64)), -1),
    )
    return out
