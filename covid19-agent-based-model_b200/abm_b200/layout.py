"""Structure-of-arrays agent store: host-side packing / unpacking.

Device layout (one slot per agent; sub-tiles of SUBTILE = 128 slots, one per warp; tiles of TILE = 1024
slots, one per thread block):
  hot, read every sub-step (8 B/agent):
    flags  uint32   bit fields of constants.F_*  (epidemic / clinical / symptomatic state, location kind,
                    away, district_mover, economic status, age, household-head bit, outcome bits, ...)
    aux    uint32   high 16: sub-step index at which the next scheduled event fires (FIRE_NEVER = none)
                    low 16: current district id
  cold, touched only when an event fires, on infection, or on the two movement sub-steps for travellers:
    t_infected, t_contagious, t_onset, t_recovered, t_return   int32 minute ticks
    inf_at uint32   (location kind << 16) | district at infection
    move_class uint16  index into the movement-probability threshold table (read at go_out only)
    hh_index   uint8   index of the slot's household within its sub-tile (static; read by warps that hold
                       a contagious agent at home)
    econ_pool  uint32  outside-pool row for EKIND_POOL agents (schools), allocated only when used
Agents are in canonical order (home district, household).  A sub-tile holds whole households (of
<= BIG_HOUSEHOLD members) of ONE home district, and slot offset o of a sub-tile holds canonical agent id
base + o with base a multiple of 4: when the id b of its first agent is not, the first b & 3 slots are
fillers.  So a warp sees whole households, the home district is a per-warp scalar, and the four agents
of a lane share one Philox group (id >> 2).  Remaining slots at the end of a sub-tile are fillers too.
Households larger than BIG_HOUSEHOLD are flagged F_BIGHH, placed member by member and use a
global-atomic path.  `sub_info[u]` = base | home_district << 48.

This mirrors what Country.initialize_agent_vectors / initialize_epidemic_vectors build
(/root/reference/src/covid19_abm/base_model.py:527-634) with 31 eight-byte vectors per agent.
"""
from dataclasses import dataclass

import numpy as np

from . import constants as C
from .spec import Population

T_NEVER = int(C.T_NEVER)
SUBS_PER_TILE = C.TILE // C.SUBTILE


@dataclass
class SlotMap:
    n_agents: int
    n_tiles: int               # thread-block tiles of TILE slots
    n_subtiles: int            # n_tiles * SUBS_PER_TILE (trailing ones may be empty)
    sub_first: np.ndarray      # int64 [n_subtiles + 1] local index of the first agent of each sub-tile
    sub_lead: np.ndarray       # int64 [n_subtiles] filler slots at the start of each sub-tile (0..3)
    sub_home: np.ndarray       # int64 [n_subtiles] home district of the agents of each sub-tile
    slot_of: np.ndarray        # int64 [n_agents] local agent -> slot
    big_start: np.ndarray      # int64 [n_big + 1] global id of first member of each big household (+ sentinel)
    big_size: np.ndarray       # int64 [n_big]
    agent_id_base: int
    sub_id: np.ndarray = None  # int64 [n_subtiles] global id of the first agent of each sub-tile (None: contiguous ids)
    _agent_of: np.ndarray = None

    @property
    def n_slots(self):
        return self.n_tiles * C.TILE

    @property
    def agent_of(self):
        """int64 [n_slots] slot -> local agent, -1 for fillers."""
        if self._agent_of is None:
            a = np.full(self.n_slots, -1, dtype=np.int64)
            a[self.slot_of] = np.arange(self.n_agents, dtype=np.int64)
            self._agent_of = a
        return self._agent_of

    @property
    def sub_info(self):
        """uint64 [n_subtiles]: canonical id of slot 0 (a multiple of 4) | home district << 48."""
        first_id = self.sub_first[:-1] + self.agent_id_base if self.sub_id is None else self.sub_id
        base = first_id - self.sub_lead
        return (base.astype(np.uint64) | (self.sub_home.astype(np.uint64) << np.uint64(48)))


def build_slot_map(household: np.ndarray, district: np.ndarray, agent_id_base: int = 0, agent_ids=None) -> SlotMap:
    """Greedy sub-tile packing.  Units are whole small households and single members of big ones.

    `agent_ids` (optional, int64 [n], strictly increasing): the global canonical id of every agent when the shard is
    not one contiguous range of the canonical order (e.g. one piece of every district); ids then run contiguously
    inside a run and a new run starts a new sub-tile, so that slot offset o of a sub-tile still holds id base + o."""
    n = int(household.shape[0])
    if n == 0:
        raise ValueError('empty population')
    if agent_ids is None:
        ids = None
    else:
        ids = np.ascontiguousarray(agent_ids, dtype=np.int64)
        if ids.shape[0] != n or (n > 1 and np.any(np.diff(ids) <= 0)):
            raise ValueError('agent_ids must be strictly increasing, one per agent')
        agent_id_base = int(ids[0])
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
    unit_start = unit_end - unit_len
    m = unit_start.shape[0]
    # a sub-tile never continues into another home district
    d_unit = np.asarray(district)[unit_start]
    new_run = d_unit[1:] != d_unit[:-1]                                  # units that start a district ...
    if ids is not None:                                                  # ... or a new run of consecutive ids
        new_run = new_run | (ids[unit_start[1:]] != ids[unit_start[1:] - 1] + 1)
    brk = np.nonzero(np.r_[True, new_run])[0]
    next_brk = np.r_[brk[1:], m][np.searchsorted(brk, np.arange(m), side='right') - 1]
    # capacity of a sub-tile that starts with unit i: the leading (id & 3) slots are fillers
    lead_unit = ((unit_start + agent_id_base) if ids is None else ids[unit_start]) & 3
    nxt = np.minimum(np.searchsorted(unit_end, unit_start + (C.SUBTILE - lead_unit), side='right'), next_brk)
    if np.any(nxt <= np.arange(m)):
        raise RuntimeError('placement unit larger than a sub-tile')
    nxt_l = nxt.tolist()
    chain = [0]
    i = 0
    while i < m:                      # one hop per sub-tile
        i = nxt_l[i]
        chain.append(i)
    chain = np.array(chain, dtype=np.int64)
    first_unit = chain[:-1]
    n_sub_used = first_unit.shape[0]
    sub_first = np.r_[unit_start[first_unit], n].astype(np.int64)
    sub_lead = lead_unit[first_unit].astype(np.int64)
    sub_home = d_unit[first_unit].astype(np.int64)
    n_tiles = (n_sub_used + SUBS_PER_TILE - 1) // SUBS_PER_TILE
    pad = n_tiles * SUBS_PER_TILE - n_sub_used
    counts = np.diff(sub_first)
    if np.any(counts + sub_lead > C.SUBTILE):
        raise RuntimeError('sub-tile overflow')
    shift = np.arange(n_sub_used, dtype=np.int64) * C.SUBTILE + sub_lead - sub_first[:-1]
    slot_of = np.arange(n, dtype=np.int64) + np.repeat(shift, counts)
    # empty trailing sub-tiles: all fillers, ids past the last agent (rounded up to a multiple of 4)
    end_id = n + agent_id_base if ids is None else int(ids[-1]) + 1
    pad_lead = np.full(pad, end_id & 3, dtype=np.int64)       # keeps base = end_id - lead a multiple of 4
    sub_first = np.r_[sub_first, np.full(pad, n, dtype=np.int64)]
    sub_lead = np.r_[sub_lead, pad_lead]
    sub_home = np.r_[sub_home, np.zeros(pad, dtype=np.int64)]
    if ids is None:
        big_start = np.append(starts[big], n).astype(np.int64) + agent_id_base
        sub_id = None
    else:
        big_start = np.append(ids[starts[big]], end_id).astype(np.int64)
        sub_id = np.r_[ids[sub_first[:n_sub_used]], np.full(pad, end_id, dtype=np.int64)]
    return SlotMap(n_agents=n, n_tiles=n_tiles, n_subtiles=n_tiles * SUBS_PER_TILE, sub_first=sub_first,
                   sub_lead=sub_lead, sub_home=sub_home, slot_of=slot_of, big_start=big_start,
                   big_size=sizes[big].astype(np.int64), agent_id_base=int(agent_id_base), sub_id=sub_id)


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
    never = np.uint32(C.FIRE_NEVER << C.AUX_FIRE_SHIFT)
    aux = np.full(sm.n_slots, never, dtype=np.uint32)
    aux[sm.slot_of] = pop.district.astype(np.uint32) | never
    return flags, aux


def household_index(flags: np.ndarray):
    """uint8 [n_slots]: number of household heads (fillers count as heads) up to each slot within its
    sub-tile, minus one.  255 for members of a big household that continue at the start of a sub-tile
    (they never use it)."""
    head = ((flags & np.uint32(C.F_HEAD)) != 0).reshape(-1, C.SUBTILE)
    idx = np.cumsum(head, axis=1, dtype=np.int32) - 1
    return (idx & 0xFF).astype(np.uint8).reshape(-1)


def update_static_fields(flags: np.ndarray, pop: Population, sm: SlotMap):
    """Re-encode the fields scenario hooks mutate on the host (district_mover, econ-location kind)
    into an existing flags array (base_model.py:642-731 helpers), keeping all dynamic state."""
    keep = ~np.uint32(C.F_DMOVER | (C.F_EKIND_MASK << C.F_EKIND_SHIFT))
    f = flags[sm.slot_of] & keep
    f |= (pop.district_mover.astype(np.uint32) & 1) * np.uint32(C.F_DMOVER)
    f |= pop.econ_kind.astype(np.uint32) << np.uint32(C.F_EKIND_SHIFT)
    flags[sm.slot_of] = f
    return flags


def unpack_state(dev: dict, pop: Population, sm: SlotMap, n_districts: int, household_ids=None, econ_loc=None):
    """Device arrays (numpy copies over slots) -> reference-style vectors over local agents, in the
    representation of oracle.abm_oracle.OracleModel.state_dict (ticks; T_NEVER = datetime.max).
    `household_ids` / `econ_loc` override the location ids of the agents' households / economic-activity
    locations (the drop-in `Country` numbers households by sorted name, base_model.py:577-581)."""
    s = sm.slot_of
    f = dev['flags'][s]
    aux = dev['aux'][s]
    D = n_districts
    epi = ((f >> C.F_EPI_SHIFT) & C.F_EPI_MASK).astype(np.int8)
    clin = ((f >> C.F_CLIN_SHIFT) & C.F_CLIN_MASK).astype(np.int8)
    sym_status = np.where((f & C.F_SYMPT) != 0, 1, np.where((f & C.F_ASYMPT) != 0, 0, -1)).astype(np.int8)
    kind = ((f >> C.F_LOC_SHIFT) & C.F_LOC_MASK).astype(np.int64)
    cur = (aux & 0xFFFF).astype(np.int64)
    if household_ids is None:
        household_ids = pop.household.astype(np.int64) + D
    H = pop.n_households
    if econ_loc is None:
        econ_loc = np.where(pop.econ_kind == C.EKIND_HOUSEHOLD, household_ids, pop.district.astype(np.int64))
        if pop.econ_pool is not None:
            econ_loc = np.where(pop.econ_kind == C.EKIND_POOL, D + H + (pop.econ_pool.astype(np.int64) - D), econ_loc)

    def loc_from(kind, cur):
        return np.select([kind == C.LOC_HOME, kind == C.LOC_HOMEPOOL, kind == C.LOC_DEST, kind == C.LOC_POOL,
                          kind == C.LOC_HOSP],
                         [household_ids, pop.district.astype(np.int64), cur, econ_loc, -cur], default=-1)

    away = (f & C.F_AWAY) != 0
    infected = dev['t_infected'][s] != T_NEVER
    onset = dev['t_onset'][s].astype(np.int64)
    has_h = (f & C.F_OUT_H) != 0
    has_c = (f & C.F_OUT_C) != 0
    has_d = (f & C.F_OUT_D) != 0
    never = np.int64(T_NEVER)
    inf_at = dev['inf_at'][s]
    out = dict(
        epidemic_state=epi, clinical_state=clin, infected_symptomatic_status=sym_status,
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
