"""CUDA path vs oracle / reference fixtures -- the parity tests proper (run on the B200 box).

Everything goes through the C ABI (abm_b200.engine.Simulation -> libabm_b200.so).  Bar: bit-exact for
every state, location and tick vector and for every log counter.
"""
import numpy as np
import pytest

from abm_b200 import synth, tables
from oracle.abm_oracle import OracleModel
from util import assert_state_equal, load_fixture, oracle_params, small_case

pytestmark = pytest.mark.gpu


def _sim(spec, pop, seed, tab=None, **kw):
    from abm_b200.engine import Simulation
    return Simulation(spec, pop, seed=seed, device=0, tables=tab, **kw)


@pytest.mark.parametrize('name', ['unmitigated', 'isolate_symptomatic', 'lockdown'])
def test_cuda_matches_reference_fixture(built_library, name):
    """CUDA state == state produced by the real reference functions (golden fixture), at every snapshot."""
    fx = load_fixture(name)
    spec, pop = fx['spec'], fx['pop']
    sim = _sim(spec, pop, fx['philox_seed'])
    sim.seed_infections(fx['seeds'])
    done = 0
    for k, snap in fx['snaps']:
        sim.step(k - done)
        done = k
        assert_state_equal(sim.state(), snap, f'{name} after sub-step {k}')
    log = sim.log_dicts()
    assert len(log) == len(fx['ref_log'])
    for got, want in zip(log, fx['ref_log']):
        for key, val in want.items():
            if key != 'date':
                assert got[key] == val, (name, key, want['date'], got[key], val)
        assert got['date'] == want['date']
    sim.close()


def test_cuda_matches_oracle_every_substep(built_library):
    """Lock-step with the oracle for 12 simulated days, flushing after every sub-step."""
    spec, pop, seeds, tab = small_case()
    sim = _sim(spec, pop, 77, tab)
    ora = OracleModel(pop, oracle_params(spec), tab, 77)
    sim.seed_infections(seeds)
    ora.set_epidemic_status(seeds)
    assert_state_equal(sim.state(), ora.state_dict(), 'after seeding')
    for k in range(6 * 12):
        sim.step(1)
        ora.step()
        assert_state_equal(sim.state(), ora.state_dict(), f'sub-step {k}')
        r, n, t = sim.debug_tables()
        assert np.array_equal(r, ora.last_tables['rsum']), f'pressure table, sub-step {k}'
        assert np.array_equal(n, ora.last_tables['ncnt']), f'headcount table, sub-step {k}'
        assert np.array_equal(t, ora.last_tables['thr_out']), f'threshold table, sub-step {k}'
    sim.close()


def test_fused_equals_flushed(built_library):
    """Results do not depend on whether the infection draw is fused into the next launch or flushed."""
    spec, pop, seeds, tab = small_case(n_agents=20_000, seed=9)
    a, b = _sim(spec, pop, 5, tab), _sim(spec, pop, 5, tab)
    a.seed_infections(seeds)
    b.seed_infections(seeds)
    a.step(6 * 10)
    for _ in range(6 * 10):
        b.step(1)
        b.flush()
    assert_state_equal(a.state(), b.state(), 'fused vs flushed')
    assert np.array_equal(a.read_log(), b.read_log())
    a.close(); b.close()


def test_long_run_with_big_households_and_mild_symptoms(built_library):
    """60 days incl. deaths; mining-camp sized households (atomic path) and symptomatic isolation."""
    spec = synth.make_spec(n_districts=6, R0=2.8, seed=21, mild_symptom_move_prob=0.05)
    pop = synth.make_population(25_000, spec, seed=21)
    # merge runs of households into three big "camps" (scenario_models.py:232-236 does this for miners)
    hh = pop.household.copy()
    for lo in (1000, 9000, 17000):
        hh[lo:lo + 400] = hh[lo]
    _, pop.household = np.unique(hh, return_inverse=True)
    pop.household = pop.household.astype(np.int64)
    seeds = np.unique(np.concatenate([synth.pick_seeds(pop, spec, infect_num=250, seed=21), [1005, 9010, 17020]]))
    tab = tables.build_tables(spec)
    sim = _sim(spec, pop, 99, tab)
    ora = OracleModel(pop, oracle_params(spec), tab, 99)
    sim.seed_infections(seeds)
    ora.set_epidemic_status(seeds)
    for day in range(60):
        sim.step(6)
        for _ in range(6):
            ora.step()
        if day % 10 == 9:
            assert_state_equal(sim.state(), ora.state_dict(), f'day {day}')
    got = sim.log_dicts()
    assert len(got) == len(ora.log_rows)
    for g, w in zip(got, ora.log_rows):
        for key, val in w.items():
            assert g[key] == val, (key, g['date'])
    assert got[-1]['died_count'] > 0 and got[-1]['recovered_count'] > 1000
    sim.close()


def test_handwashing_risk_and_weekend(built_library):
    """hw_risk multiplier per district (base_model.py:189-191) and a start on a Saturday."""
    from datetime import datetime
    spec = synth.make_spec(n_districts=7, R0=2.4, seed=31, start_datetime=datetime(2020, 9, 5, 8))
    spec.hw_risk = np.random.default_rng(1).uniform(0.9, 1.1, spec.n_districts)
    pop = synth.make_population(15_000, spec, seed=31)
    seeds = synth.pick_seeds(pop, spec, infect_num=200, seed=31)
    tab = tables.build_tables(spec)
    sim = _sim(spec, pop, 3, tab)
    ora = OracleModel(pop, oracle_params(spec), tab, 3)
    sim.seed_infections(seeds)
    ora.set_epidemic_status(seeds)
    sim.step(6 * 15)
    for _ in range(6 * 15):
        ora.step()
    assert_state_equal(sim.state(), ora.state_dict(), 'hw risk, 15 days')
    got = sim.log_dicts()
    for g, w in zip(got, ora.log_rows):
        for key, val in w.items():
            assert g[key] == val, (key, g['date'])
    sim.close()


def test_counters_and_district_statistics(built_library):
    spec, pop, seeds, tab = small_case(n_agents=20_000, seed=13)
    sim = _sim(spec, pop, 8, tab)
    ora = OracleModel(pop, oracle_params(spec), tab, 8)
    sim.seed_infections(seeds)
    ora.set_epidemic_status(seeds)
    sim.step(6 * 9)
    for _ in range(6 * 9):
        ora.step()
    c = sim.counters()
    assert c['current_exposed_cases'] == int((ora.epidemic_state == 1).sum())
    assert c['current_contagious_cases'] == int((ora.epidemic_state == 2).sum())
    assert c['asymptomatic_count'] == int((ora.infected_symptomatic_status == 0).sum())
    assert c['symptomatic_count'] == int((ora.infected_symptomatic_status == 1).sum())
    assert c['infected_count'] == int((ora.date_infected < ora.now()).sum())
    sym, dpop = sim.district_symptomatic()
    want_sym = np.bincount(pop.district[np.isin(ora.epidemic_state, [1, 2]) & (ora.infected_symptomatic_status == 1)],
                           minlength=spec.n_districts)
    assert np.array_equal(sym, want_sym)
    assert np.array_equal(dpop, np.bincount(pop.district, minlength=spec.n_districts))
    sim.close()


def test_graph_replay_equals_plain_launches(built_library):
    """Whole days replayed as one CUDA graph (device-resident clock) give the same trajectory as
    sub-step-by-sub-step launches, including the day log."""
    spec, pop, seeds, tab = small_case(n_agents=20_000, seed=17)
    a, b = _sim(spec, pop, 6, tab, use_graph=True), _sim(spec, pop, 6, tab, use_graph=False)
    for s in (a, b):
        s.seed_infections(seeds)
    a.step(6 * 21 + 2)                   # graph path for the whole days in the middle, plain launches at both ends
    for n in (5, 6 * 20, 3):
        b.step(n)
    assert_state_equal(a.state(), b.state(), 'graph vs plain')
    assert np.array_equal(a.read_log(), b.read_log())
    assert a.launch_count == b.launch_count
    a.close(); b.close()


def test_checkpoint_restore_resumes_identically(built_library):
    """checkpoint() -> fresh handle -> restore() continues exactly like the uninterrupted run
    (the reference's pickle / unpickle of the model, scenario_models.py:497-500)."""
    spec, pop, seeds, tab = small_case(n_agents=20_000, seed=19)
    a = _sim(spec, pop, 11, tab)
    a.seed_infections(seeds)
    a.step(6 * 7 + 3)                    # stop in the middle of a day, while agents are out
    ck = a.checkpoint()
    b = _sim(spec, pop, 11, tab, upload=False)
    b.restore(ck)
    assert_state_equal(a.state(), b.state(), 'restored state')
    a.step(6 * 6)
    b.step(6 * 6)
    assert_state_equal(a.state(), b.state(), 'after resuming')
    la, lb = a.log_dicts(), b.log_dicts()
    assert la[-len(lb):] == lb and len(lb) == 6
    assert a.counters() == b.counters()
    a.close(); b.close()


def test_five_hundred_districts(built_library):
    """BASELINE.json configs[4] shape: 500 districts (OD rows of 500 columns, 500-row tables), small population."""
    spec = synth.make_spec(n_districts=500, R0=2.6, seed=51)
    pop = synth.make_population(60_000, spec, seed=51)
    seeds = synth.pick_seeds(pop, spec, infect_num=400, seed=51, cases=synth.seed_case_counts(pop, spec, seed=51, n_seed_districts=40))
    tab = tables.build_tables(spec)
    sim = _sim(spec, pop, 15, tab)
    ora = OracleModel(pop, oracle_params(spec), tab, 15)
    sim.seed_infections(seeds)
    ora.set_epidemic_status(seeds)
    sim.step(6 * 12)
    for _ in range(6 * 12):
        ora.step()
    assert_state_equal(sim.state(), ora.state_dict(), '500 districts, 12 days')
    r, n, t = sim.debug_tables()
    assert np.array_equal(r, ora.last_tables['rsum']) and np.array_equal(n, ora.last_tables['ncnt'])
    assert np.array_equal(t, ora.last_tables['thr_out'])
    assert (ora.current_district_ids != ora.district_ids).sum() > 100       # travellers are under way
    sim.close()


def test_population_conservation_and_monotone_counters_at_scale(built_library):
    """Size-independent properties on a 2 M-agent population (too big for the oracle to follow in seconds): nobody is
    lost or duplicated, cumulative counters never decrease, the state machine only moves forward, and the running
    counters agree with an independent recount of the agent store."""
    spec = synth.make_spec(n_districts=60, R0=2.4, seed=61)
    pop = synth.make_population(2_000_000, spec, seed=61)
    seeds = synth.pick_seeds(pop, spec, infect_num=4000, seed=61)
    sim = _sim(spec, pop, 21)
    sim.seed_infections(seeds)
    prev = sim.state()['epidemic_state'].astype(np.int64)
    for _ in range(4):
        sim.step(6 * 10)
        st = sim.state()
        epi = st['epidemic_state'].astype(np.int64)
        assert epi.shape[0] == pop.n and epi.min() >= 0 and epi.max() <= 4
        assert np.all(epi >= prev)                                   # S -> I -> C -> R / D never goes back
        assert np.all((st['current_location_ids'] >= -spec.n_districts))
        dead = epi == 4
        assert np.all(st['current_location_ids'][dead] == -1) and np.all(st['clinical_state'][dead] == 3)
        infected = st['date_infected'] != 2 ** 31 - 1
        assert np.array_equal(infected, epi > 0)
        c = sim.counters()
        assert c['current_exposed_cases'] == int((epi == 1).sum()) and c['current_contagious_cases'] == int((epi == 2).sum())
        assert c['asymptomatic_count'] + c['symptomatic_count'] == int(infected.sum())
        prev = epi
    log = sim.read_log()
    for col in (0, 1, 2, 3, 4, 5, 10, 11):                          # cumulative columns of the day log
        assert np.all(np.diff(log[:, col]) >= 0), col
    assert log[-1, 1] > 100_000
    sim.close()


def test_ensemble_replicas_equal_individual_runs(built_library):
    """Members run concurrently on their own streams (abm_b200.ensemble) give exactly the curves of one-at-a-time runs."""
    from abm_b200 import ensemble
    from abm_b200.engine import COUNTER_NAMES
    spec, pop, seeds, tab = small_case(n_agents=30_000, seed=27)
    keys = [101, 202, 303, 404, 505]
    seed_sets = [synth.pick_seeds(pop, spec, infect_num=200, seed=s) for s in range(5)]
    curves, dates = ensemble.run_ensemble(spec, pop, seed_sets, keys, days=12, max_concurrent=3, tables=tab)
    assert curves.shape == (5, 12, 12) and len(dates) == 12
    for m in (0, 3, 4):
        sim = _sim(spec, pop, keys[m], tab)
        sim.seed_infections(seed_sets[m])
        sim.step(6 * 12)
        assert np.array_equal(sim.read_log()[:, :12], curves[m]), m
        sim.close()
    assert not np.array_equal(curves[0], curves[1])
    band = ensemble.summarize(curves)
    assert band['mean'].shape == (12, 12) and np.all(band['lo'] <= band['hi'])
    assert curves[:, -1, COUNTER_NAMES.index('infected_count')].min() > 200


@pytest.mark.parametrize('step_minutes,start', [(120, (2020, 9, 1, 8)), (480, (2020, 9, 1, 0)), (240, (2020, 9, 1, 8)),
                                                (60, (2020, 9, 4, 22))])
def test_other_time_steps_and_start_hours(built_library, step_minutes, start):
    """The sub-step need not be 4 h and the run need not start at midnight (the reference's default start is 08:00,
    params.py:369): 2 h (12 sub-steps a day), 8 h (an odd number: 3), 1 h starting on a Friday night."""
    from datetime import datetime
    spec = synth.make_spec(n_districts=5, R0=3.0, seed=71, start_datetime=datetime(*start), step_minutes=step_minutes)
    pop = synth.make_population(12_000, spec, seed=71)
    seeds = synth.pick_seeds(pop, spec, infect_num=200, seed=71)
    tab = tables.build_tables(spec)
    sim = _sim(spec, pop, 33, tab)
    ora = OracleModel(pop, oracle_params(spec), tab, 33)
    sim.seed_infections(seeds)
    ora.set_epidemic_status(seeds)
    per_day = 1440 // step_minutes
    n = per_day * 8 + 5
    sim.step(n)
    for _ in range(n):
        ora.step()
    assert_state_equal(sim.state(), ora.state_dict(), f'{step_minutes} min steps from {start}')
    got = sim.log_dicts()
    assert len(got) == len(ora.log_rows) and len(got) >= 8
    for g, w in zip(got, ora.log_rows):
        for key, val in w.items():
            assert g[key] == val, (key, g['date'])
    sim.close()


@pytest.mark.parametrize('n_agents,n_districts', [(300, 1), (1100, 2), (4099, 3)])
def test_tiny_and_ragged_populations(built_library, n_agents, n_districts):
    """One district, populations smaller than a tile, agent counts that leave ragged last sub-tiles."""
    spec = synth.make_spec(n_districts=n_districts, R0=4.0, seed=81)
    pop = synth.make_population(n_agents, spec, seed=81)
    rng = np.random.default_rng(81)
    eligible = np.nonzero((pop.age > 20) & (pop.age < 60))[0]
    seeds = np.sort(rng.choice(eligible, size=min(25, eligible.shape[0]), replace=False)).astype(np.int64)
    tab = tables.build_tables(spec)
    sim = _sim(spec, pop, 44, tab)
    ora = OracleModel(pop, oracle_params(spec), tab, 44)
    sim.seed_infections(seeds)
    ora.set_epidemic_status(seeds)
    for _ in range(5):
        sim.step(6 * 4)
        for _ in range(6 * 4):
            ora.step()
        assert_state_equal(sim.state(), ora.state_dict(), f'{n_agents} agents, {n_districts} district(s)')
    assert sim.read_log().shape[0] == 20
    sim.close()


def test_no_infection_and_empty_seed_list(built_library):
    """Nothing seeded: nobody is ever infected, people still move; seeding an empty list is a no-op."""
    spec, pop, _, tab = small_case(n_agents=9000, seed=91)
    sim = _sim(spec, pop, 1, tab)
    sim.seed_infections(np.zeros(0, dtype=np.int64))
    sim.step(6 * 3 + 3)                               # stop at noon: people are out
    st = sim.state()
    assert not st['epidemic_state'].any() and np.all(st['date_infected'] == 2 ** 31 - 1)
    assert (st['current_location_ids'] < spec.n_districts).sum() > 5000
    log = sim.log_dicts()
    assert len(log) == 4 and all(r['infected_count'] == 0 and r['current_contagious_cases'] == 0 for r in log)
    sim.close()
