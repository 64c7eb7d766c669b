import numpy as np

from abm_b200 import constants as C, layout, synth


def test_slot_map_keeps_households_inside_tiles():
    spec = synth.make_spec(n_districts=5, seed=1)
    pop = synth.make_population(40_000, spec, seed=1)
    hh = pop.household.copy()
    hh[5000:5700] = hh[5000]                      # one camp-sized household
    _, hh = np.unique(hh, return_inverse=True)
    sm = layout.build_slot_map(hh.astype(np.int64), agent_id_base=1_000_000)
    assert sm.n_slots % C.TILE == 0 and sm.n_slots >= pop.n
    assert sm.n_slots - pop.n < C.TILE + 0.01 * pop.n
    assert np.array_equal(np.sort(sm.slot_of), sm.slot_of)           # canonical order preserved
    tile = sm.slot_of // C.TILE
    sizes = np.bincount(hh)
    small = sizes[hh] <= C.BIG_HOUSEHOLD
    first = np.r_[True, hh[1:] != hh[:-1]]
    start_tile = tile[first][hh]                                       # tile of each agent's household head
    assert np.all(tile[small] == start_tile[small])
    assert sm.big_size.tolist() == [sizes.max()] and sm.big_start[0] == 1_000_000 + np.argmax(first & (sizes[hh] > 32))
    # tile_base is the canonical id of the first agent in each tile; fillers only at tile ends
    for t in range(sm.n_tiles):
        cnt = sm.tile_base[t + 1] - sm.tile_base[t]
        a = sm.agent_of[t * C.TILE:(t + 1) * C.TILE]
        assert np.array_equal(a[:cnt], np.arange(cnt) + sm.tile_base[t] - 1_000_000) and np.all(a[cnt:] == -1)


def test_pack_unpack_roundtrip_initial_state():
    spec = synth.make_spec(n_districts=4, seed=2)
    pop = synth.make_population(5_000, spec, seed=2)
    sm = layout.build_slot_map(pop.household)
    flags, aux = layout.pack_flags(pop, sm)
    never = np.full(sm.n_slots, C.T_NEVER, np.int32)
    dev = dict(flags=flags, aux=aux, t_infected=never, t_contagious=never, t_onset=never, t_recovered=never,
               t_return=never, inf_at=np.zeros(sm.n_slots, np.uint32))
    st = layout.unpack_state(dev, pop, sm, spec.n_districts)
    assert np.all(st['epidemic_state'] == 0) and np.all(st['infected_symptomatic_status'] == -1)
    assert np.array_equal(st['current_location_ids'], pop.household + spec.n_districts)
    assert np.array_equal(st['current_district_ids'], pop.district)
    assert np.all(st['date_died'] == C.T_NEVER) and np.all(st['infected_at_location_ids'] == -1)
    fill = flags[sm.agent_of < 0]
    assert np.all(fill & C.F_FILLER) and np.all((fill & 7) == C.STATE_DEAD)
