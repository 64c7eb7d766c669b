import numpy as np

from abm_b200 import constants as C, layout, synth


def test_slot_map_invariants():
    """Sub-tiles hold whole small households of one home district; slot offset o <-> id base + o, base % 4 == 0."""
    spec = synth.make_spec(n_districts=5, seed=1)
    pop = synth.make_population(40_000, spec, seed=1)
    hh = pop.household.copy()
    hh[5000:5700] = hh[5000]                      # one camp-sized household
    _, hh = np.unique(hh, return_inverse=True)
    hh = hh.astype(np.int64)
    base = 1_000_003                              # odd id base: leading fillers are needed
    sm = layout.build_slot_map(hh, pop.district, agent_id_base=base)
    S = C.SUBTILE
    assert sm.n_slots % C.TILE == 0 and sm.n_slots >= pop.n and sm.n_subtiles * S == sm.n_slots
    assert sm.n_slots - pop.n < C.TILE + 0.05 * pop.n
    assert np.all(np.diff(sm.slot_of) > 0)                             # canonical order preserved
    sub = sm.slot_of // S
    sizes = np.bincount(hh)
    small = sizes[hh] <= C.BIG_HOUSEHOLD
    first = np.r_[True, hh[1:] != hh[:-1]]
    start_sub = sub[first][hh]                                         # sub-tile of each agent's household head
    assert np.all(sub[small] == start_sub[small])                      # small households do not straddle
    for u in np.unique(sub):                                           # one home district per sub-tile
        d = pop.district[sub == u]
        assert d.min() == d.max() == sm.sub_home[u]
    info = sm.sub_info
    sub_base = (info & np.uint64((1 << 48) - 1)).astype(np.int64)
    assert np.all(sub_base % 4 == 0)
    assert np.array_equal((info >> np.uint64(48)).astype(np.int64), sm.sub_home)
    ids = np.arange(pop.n) + base
    assert np.array_equal(sub_base[sub] + sm.slot_of % S, ids)         # slot offset o holds id base + o
    assert sm.big_size.tolist() == [sizes.max()] and sm.big_start[0] == base + np.argmax(first & (sizes[hh] > 32))
    a = sm.agent_of
    assert np.array_equal(a[sm.slot_of], np.arange(pop.n)) and (a >= 0).sum() == pop.n


def test_slot_map_is_shard_invariant():
    """A household-aligned shard gets the same sub-tile bases as the same agents in the whole population
    when cut at a district boundary (ids, hence Philox groups, do not depend on the number of ranks)."""
    spec = synth.make_spec(n_districts=4, seed=3)
    pop = synth.make_population(20_000, spec, seed=3)
    cut = int(np.searchsorted(pop.district, 2))
    whole = layout.build_slot_map(pop.household, pop.district)
    part = layout.build_slot_map(pop.household[cut:] - pop.household[cut], pop.district[cut:], agent_id_base=cut)
    mask = (1 << 48) - 1
    wb = (whole.sub_info & np.uint64(mask)).astype(np.int64)
    pb = (part.sub_info & np.uint64(mask)).astype(np.int64)
    used = np.unique(part.slot_of // C.SUBTILE)
    assert set(pb[used]).issubset(set(wb))


def test_pack_unpack_roundtrip_initial_state():
    spec = synth.make_spec(n_districts=4, seed=2)
    pop = synth.make_population(5_000, spec, seed=2)
    sm = layout.build_slot_map(pop.household, pop.district)
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


def test_slot_map_properties_on_random_households():
    """Property test: any grouping into households (sizes 1..90, i.e. below and above the 32-member tile), any
    district boundaries and any id base give a slot map that keeps the canonical order, never splits a small household
    or a lane's four-id group across sub-tiles, keeps one home district per sub-tile and numbers slots by id."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(sizes=st.lists(st.integers(1, 90), min_size=1, max_size=120),
           n_districts=st.integers(1, 5), base=st.integers(0, 2 ** 40), seed=st.integers(0, 2 ** 31))
    def check(sizes, n_districts, base, seed):
        rng = np.random.default_rng(seed)
        hh = np.repeat(np.arange(len(sizes), dtype=np.int64), sizes)
        hh_district = np.sort(rng.integers(0, n_districts, len(sizes)))          # households grouped by district
        district = hh_district[hh].astype(np.int32)
        sm = layout.build_slot_map(hh, district, agent_id_base=base)
        S = C.SUBTILE
        n = hh.shape[0]
        assert sm.n_slots % C.TILE == 0 and sm.n_subtiles * S == sm.n_slots
        assert np.all(np.diff(sm.slot_of) > 0) and sm.slot_of[-1] < sm.n_slots
        sub = sm.slot_of // S
        size_of = np.asarray(sizes)[hh]
        head_sub = sub[np.r_[0, np.cumsum(sizes)[:-1]]][hh]
        assert np.all(sub[size_of <= C.BIG_HOUSEHOLD] == head_sub[size_of <= C.BIG_HOUSEHOLD])
        sub_base = (sm.sub_info & np.uint64((1 << 48) - 1)).astype(np.int64)
        home = (sm.sub_info >> np.uint64(48)).astype(np.int64)
        assert np.all(sub_base % 4 == 0)
        assert np.array_equal(sub_base[sub] + sm.slot_of % S, np.arange(n) + base)
        assert np.array_equal(home[sub], district)
        big = [i for i, s in enumerate(sizes) if s > C.BIG_HOUSEHOLD]
        assert sm.big_size.tolist() == [sizes[i] for i in big]
        starts = np.r_[0, np.cumsum(sizes)[:-1]]
        assert sm.big_start[:-1].tolist() == [int(starts[i]) + base for i in big]
        a = sm.agent_of
        assert np.array_equal(a[sm.slot_of], np.arange(n)) and (a >= 0).sum() == n
        # the household index inside the sub-tile that the kernel's segmented sum uses
        pop = synth.Population(district=district, household=hh, age=np.zeros(n, np.int32), status=np.zeros(n, np.int32),
                               econ_kind=np.zeros(n, np.int8), district_mover=np.zeros(n, np.int8),
                               move_prob_weekday=np.ones(n), move_prob_other=np.ones(n))
        flags, aux = layout.pack_flags(pop, sm)
        idx = layout.household_index(flags).view(np.uint8)
        real = a >= 0
        per_slot_hh = np.where(real, hh[np.maximum(a, 0)], -1)
        for u in np.unique(sub):
            sl = slice(u * S, (u + 1) * S)
            small = real[sl] & (np.asarray(sizes)[np.maximum(per_slot_hh[sl], 0)] <= C.BIG_HOUSEHOLD)
            if small.any():
                hh_here, ix_here = per_slot_hh[sl][small], idx[sl][small]
                # same household <-> same index, and indices are dense from 0 in slot order
                assert len(set(zip(hh_here.tolist(), ix_here.tolist()))) == len(set(hh_here.tolist())) == len(set(ix_here.tolist()))

    check()


def test_slot_map_of_a_dealt_shard():
    """A shard made of one piece of every district (sharding.deal_population): per-agent global ids, a new sub-tile at
    every id run, slot offset o still holds id base + o with base % 4 == 0, big households keep their global start."""
    from abm_b200 import sharding
    spec = synth.make_spec(n_districts=7, seed=4)
    pop = synth.make_population(30_000, spec, seed=4)
    hh = pop.household.copy()
    hh[12_000:12_150] = hh[12_000]                  # a camp-sized household inside some district
    _, hh = np.unique(hh, return_inverse=True)
    pop.household = hh.astype(np.int64)
    S = C.SUBTILE
    whole = layout.build_slot_map(pop.household, pop.district)
    covered = np.zeros(pop.n, dtype=int)
    for world in (2, 3):
        for idx in sharding.deal_population(pop, world):
            shard = sharding.take_population(pop, idx)
            assert np.array_equal(np.unique(shard.district), np.arange(7))            # a piece of every district
            sm = layout.build_slot_map(shard.household, shard.district, agent_ids=idx)
            sub = sm.slot_of // S
            sub_base = (sm.sub_info & np.uint64((1 << 48) - 1)).astype(np.int64)
            assert np.all(sub_base % 4 == 0)
            assert np.array_equal(sub_base[sub] + sm.slot_of % S, idx)
            assert np.array_equal((sm.sub_info >> np.uint64(48)).astype(np.int64)[sub], shard.district)
            sizes = np.bincount(shard.household)
            head = np.r_[True, shard.household[1:] != shard.household[:-1]]
            small = sizes[shard.household] <= C.BIG_HOUSEHOLD
            assert np.all(sub[small] == sub[head][shard.household][small])
            # households are never cut by the deal, and the camp keeps its global first id
            assert pop.household[idx[0]] != pop.household[idx[0] - 1] if idx[0] else True
            if sm.big_size.size:
                assert sm.big_size.tolist() == [150] and sm.big_start[0] == whole.big_start[0]
            if world == 2:
                covered[idx] += 1
    assert np.all(covered == 1)
