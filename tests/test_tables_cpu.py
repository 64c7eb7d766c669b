"""Host-built integer tables and the oracle's integer arithmetic: exactness / accuracy properties (no GPU)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from abm_b200 import constants as C, layout, synth, tables
from oracle import abm_oracle as ora


def test_bernoulli_thresholds_are_exact():
    """(x >> 1) < ceil(p 2^31)  <=>  (x >> 1) / 2^31 < p, for probabilities on and off the 31-bit grid."""
    rng = np.random.default_rng(0)
    p = np.r_[0.0, 1.0, 0.5, 0.8, 0.6, 1e-9, 1 - 1e-12, rng.random(200), rng.integers(0, 2 ** 31, 50) / 2.0 ** 31]
    thr = tables.bern_threshold(p).astype(np.int64)
    x = np.r_[0, 1, 2 ** 31 - 1, rng.integers(0, 2 ** 31, 3000)].astype(np.int64)
    for pi, ti in zip(p, thr):
        assert np.array_equal(x < ti, x / 2.0 ** 31 < pi), pi
        for edge in (ti - 1, ti):                    # the two integers around the threshold
            if 0 <= edge < 2 ** 31:
                assert (edge < ti) == (edge / 2.0 ** 31 < pi), (pi, edge)


def test_hazard_threshold_is_monotone_and_accurate():
    """round((1 - exp(-h)) 2^31) in integer arithmetic: within 2 units of the float value, never decreasing."""
    omexp = tables.omexp_table()
    rng = np.random.default_rng(1)
    h = np.sort(np.r_[0, 1, 2 ** 24 - 1, 2 ** 24, rng.integers(0, 2 ** 32, 4000), rng.integers(0, 40 << 32, 4000)].astype(np.uint64))
    thr = ora.hazard_threshold_fx(omexp, h).astype(np.int64)
    want = -np.expm1(-h.astype(np.float64) / 2.0 ** 32) * 2.0 ** 31
    assert np.all(np.diff(thr) >= 0)
    assert np.abs(thr - want).max() <= 2.0
    assert thr[0] == 0 and thr[-1] == 2 ** 31
    assert ora.hazard_threshold_fx(omexp, np.array([np.uint64(tables.OMEXP_N) << np.uint64(24)]))[0] == 2 ** 31


def test_quantile_tables_reproduce_the_distributions():
    """Dwell-time tables: sampling through them gives the reference's Gamma / exponential moments (params.py:145-189)."""
    dwell = tables.dwell_tables()
    x = np.random.default_rng(2).integers(0, 2 ** 32, 400_000, dtype=np.uint64).astype(np.uint32)
    day = C.TICKS_PER_DAY
    for row, (mean, var) in enumerate([
            (C.INCUBATION_GAMMA[0] * C.INCUBATION_GAMMA[1], C.INCUBATION_GAMMA[0] * C.INCUBATION_GAMMA[1] ** 2),
            (C.ASYM_TO_CONTAGIOUS_MEAN_DAYS, C.ASYM_TO_CONTAGIOUS_MEAN_DAYS ** 2),
            (C.SYM_RECOVERY_GAMMA[0] * C.SYM_RECOVERY_GAMMA[1], C.SYM_RECOVERY_GAMMA[0] * C.SYM_RECOVERY_GAMMA[1] ** 2),
            (C.ASYM_RECOVERY_GAMMA[0] * C.ASYM_RECOVERY_GAMMA[1], C.ASYM_RECOVERY_GAMMA[0] * C.ASYM_RECOVERY_GAMMA[1] ** 2)]):
        s = ora.sample_quantile(dwell[row], x) / day
        assert abs(s.mean() - mean) < 0.02 * mean, (row, s.mean(), mean)
        assert abs(s.var() - var) < 0.05 * var, (row, s.var(), var)
        assert s.min() >= 0
    z = ora.sample_quantile(tables.znorm_table(), x) / 65536.0           # half-normal magnitudes
    assert abs(z.mean() - np.sqrt(2 / np.pi)) < 0.01 and abs((z ** 2).mean() - 1.0) < 0.02


def test_move_classes_keep_the_float_product():
    """Class thresholds use the float64 product p * mild exactly as base_model.py:318 forms it."""
    spec = synth.make_spec(n_districts=3, seed=4)
    pop = synth.make_population(3000, spec, seed=4)
    rng = np.random.default_rng(4)
    pop.move_prob_weekday = pop.move_prob_weekday * rng.choice([1.0, 0.67, 0.4123], pop.n)
    cls, thr, uniq = tables.move_classes(pop, 0.05)
    assert np.array_equal(uniq[cls, 0], pop.move_prob_weekday) and np.array_equal(uniq[cls, 1], pop.move_prob_other)
    assert np.array_equal(thr[cls, 1], tables.bern_threshold(pop.move_prob_weekday * 0.05))
    assert np.array_equal(thr[cls, 2], tables.bern_threshold(pop.move_prob_other))
    assert np.all(np.diff(uniq[:, 0]) >= 0)


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(min_value=1, max_value=9), min_size=1, max_size=400),
       st.lists(st.integers(min_value=33, max_value=300), min_size=0, max_size=2),
       st.integers(min_value=0, max_value=2 ** 40), st.integers(min_value=1, max_value=4), st.integers(min_value=0, max_value=10 ** 6))
def test_slot_map_properties(small_sizes, big_sizes, id_base, n_districts, seed):
    """Any mix of household sizes (camp-sized ones included), any id base: every agent gets exactly one slot, in
    order; small households stay inside one sub-tile; a sub-tile has one home district; slot offset <-> id."""
    rng = np.random.default_rng(seed)
    sizes = np.array(small_sizes + big_sizes)
    rng.shuffle(sizes)
    household = np.repeat(np.arange(sizes.shape[0], dtype=np.int64), sizes)
    hh_district = np.sort(rng.integers(0, n_districts, sizes.shape[0]))
    district = np.repeat(hh_district, sizes).astype(np.int32)
    sm = layout.build_slot_map(household, district, agent_id_base=id_base)
    n, S = household.shape[0], C.SUBTILE
    assert sm.n_slots % C.TILE == 0 and np.all(np.diff(sm.slot_of) > 0) and sm.slot_of[-1] < sm.n_slots
    sub = sm.slot_of // S
    first = np.r_[True, household[1:] != household[:-1]]
    head_sub = sub[first][household]
    small = sizes[household] <= C.BIG_HOUSEHOLD
    assert np.all(sub[small] == head_sub[small])
    assert np.array_equal(sm.sub_home[sub], district)
    base = (sm.sub_info & np.uint64((1 << 48) - 1)).astype(np.int64)
    assert np.all(base % 4 == 0)
    assert np.array_equal(base[sub] + sm.slot_of % S, np.arange(n) + id_base)
    assert sm.big_size.tolist() == sizes[sizes > C.BIG_HOUSEHOLD].tolist()
    flags, aux = layout.pack_flags(_Pop(household, district), sm)
    idx = layout.household_index(flags).reshape(-1, S)
    assert np.all(np.diff(idx.astype(np.int64), axis=1) >= -255)          # non-decreasing within a sub-tile (255 wraps to 0)


class _Pop:
    """The fields pack_flags reads."""
    def __init__(self, household, district):
        n = household.shape[0]
        self.household, self.district, self.n = household, district, n
        self.age = np.full(n, 30, np.int32)
        self.status = np.zeros(n, np.int32)
        self.econ_kind = np.ones(n, np.int8)
        self.district_mover = np.ones(n, np.int8)
