"""Device-side ingestion (abm_b200/ingest.py) builds exactly the store engine.HostStore builds on the host: every packed
array, the slot map, the movement classes and the canonical Population, from columns in shuffled PERSON order.  Runs the
torch code on the CPU device here; tests/test_gpu_ingest.py repeats it on the GPU and steps the result."""
import numpy as np
import pytest

from abm_b200 import synth
from abm_b200.engine import HostStore
from abm_b200.ingest import DeviceStore
from abm_b200.spec import Population

COLUMNS = ('district', 'household', 'age', 'status', 'econ_kind', 'district_mover', 'move_prob_weekday', 'move_prob_other')


def person_order_columns(pop, seed):
    """Census-like input: persons in random order, household codes arbitrary integers (not dense, not grouped)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(pop.n)
    code = rng.permutation(pop.n_households).astype(np.int64) * 7 + 3
    cols = {name: getattr(pop, name)[perm] for name in COLUMNS if name != 'household'}
    cols['household'] = code[pop.household][perm]
    return cols


def host_reference(spec, cols, base=0):
    """What the drop-in does on the host (covid19_abm/base_model.py `_ensure_device`): stable sort by (district, household),
    dense household index, HostStore."""
    n = cols['district'].shape[0]
    order = np.lexsort((np.arange(n), cols['household'], cols['district']))
    key = cols['district'][order].astype(np.int64) * (int(cols['household'].max()) + 1) + cols['household'][order]
    dense = np.cumsum(np.r_[True, key[1:] != key[:-1]]) - 1
    pop = Population(district=cols['district'][order].astype(np.int32), household=dense.astype(np.int64),
                     **{name: cols[name][order] for name in COLUMNS if name not in ('district', 'household')})
    return order, pop, HostStore(spec, pop, agent_id_base=base)


@pytest.mark.parametrize('n,base,big', [(30_000, 0, False), (12_345, 6, True), (2_000, 1, False)])
def test_device_store_equals_host_store(n, base, big):
    spec = synth.make_spec(n_districts=9, R0=2.0, seed=3)
    pop = synth.make_population(n, spec, seed=3)
    if big:                      # a few camps above BIG_HOUSEHOLD members
        hh = pop.household.copy()
        for lo in (500, 4000):
            hh[lo:lo + 90] = hh[lo]
        _, pop.household = np.unique(hh, return_inverse=True)
        pop.household = pop.household.astype(np.int64)
    pop.move_prob_weekday = pop.move_prob_weekday * np.where(pop.district % 2 == 0, 0.75, 1.0)     # more movement classes
    cols = person_order_columns(pop, seed=n)
    order, ref_pop, want = host_reference(spec, cols, base)
    got = DeviceStore(spec, device='cpu', agent_id_base=base, **cols)
    assert np.array_equal(got.order, order)
    for name in COLUMNS:
        assert np.array_equal(getattr(got.pop, name), getattr(ref_pop, name)), name
    for name in ('n_tiles', 'n_subtiles', 'agent_id_base', 'n_agents'):
        assert getattr(got.sm, name) == getattr(want.sm, name), name
    for name in ('sub_first', 'sub_lead', 'sub_home', 'big_start', 'big_size', 'slot_of', 'sub_info'):
        assert np.array_equal(getattr(got.sm, name), getattr(want.sm, name)), name
    assert np.array_equal(got.move_thr, want.move_thr) and np.array_equal(got.move_class, want.move_class)
    assert set(got.host) == set(want.host)
    for k, v in want.host.items():
        assert np.array_equal(got.host[k].numpy(), v.numpy()), k
    assert got.nbytes == want.nbytes


def test_household_across_districts_is_refused():
    spec = synth.make_spec(n_districts=4, seed=1)
    pop = synth.make_population(3_000, spec, seed=1)
    cols = person_order_columns(pop, seed=1)
    cols['district'] = cols['district'].copy()
    same = np.nonzero(cols['household'] == cols['household'][0])[0]
    if same.size > 1:
        cols['district'][same[0]] = (cols['district'][same[0]] + 1) % 4
        with pytest.raises(NotImplementedError, match='different home districts'):
            DeviceStore(spec, device='cpu', **cols)
