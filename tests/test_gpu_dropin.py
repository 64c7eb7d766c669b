"""The drop-in `covid19_abm` package on the real CUDA engine: the reference's driver sequence (ParamsConfig ->
scenario class -> load_agents with seeding -> step loop -> log lines -> pickle) must give, person by person, the
state an independent oracle run produces from the same inputs and draws."""
import os
from datetime import datetime

import numpy as np
import pytest

import dropin_flow

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _library(built_library):
    old = os.environ.get('COVID19_ABM_DATA_DIR')
    yield
    if old is None:
        os.environ.pop('COVID19_ABM_DATA_DIR', None)
    else:
        os.environ['COVID19_ABM_DATA_DIR'] = old


def test_unmitigated_with_pickle_resume(tmp_path):
    m, ora, rows = dropin_flow.run_flow(tmp_path, 'UnmitigatedScenario', days=14, n_agents=20_000, pickle_at=6 * 6 + 2)
    assert rows[-1]['infected_count'] > 500
    from abm_b200.engine import Simulation
    assert isinstance(m._sim, Simulation)
    m.close()


@pytest.mark.parametrize('scenario', ['LockdownGreatestMobilityScenario', 'IsolateSymptomaticScenario',
                                      'HandWashingRiskScenario', 'IsolateVulnerableScenario', 'OpenSchoolsScenario'])
def test_scenarios_match_oracle(tmp_path, scenario):
    m, ora, rows = dropin_flow.run_flow(tmp_path, scenario, days=10, n_agents=20_000)
    assert rows[-1]['infected_count'] > 200
    m.close()


def test_school_clock_hooks_refresh_the_device(tmp_path):
    m, ora, rows = dropin_flow.run_flow(tmp_path, 'AcceleratedGovernmentOpenSchoolsScenario', days=5,
                                        start=datetime(2020, 9, 29, 0), n_agents=20_000, school_columns=True)
    assert m.active_school_phases[-1] == m.max_school_phase
    safe, unsafe = m.get_safe_and_unsafe_districts()
    assert safe.shape[0] + unsafe.shape[0] == 60
    m.close()


def test_run_days_equals_single_steps(tmp_path):
    root = str(tmp_path / 'data')
    os.makedirs(root)
    dropin_flow.build_world(root, n_agents=20_000)
    a = dropin_flow.make_model('UnmitigatedScenario', datetime(2020, 9, 1, 0), log_file=str(tmp_path / 'a.log'))
    b = dropin_flow.make_model('UnmitigatedScenario', datetime(2020, 9, 1, 0), log_file=str(tmp_path / 'b.log'))
    a.run_days(9)
    for _ in range(6 * 9):
        b.step()
    assert np.array_equal(a.epidemic_state, b.epidemic_state)
    assert np.array_equal(a.current_location_ids, b.current_location_ids)
    assert np.array_equal(a.event_ticks('date_recovered'), b.event_ticks('date_recovered'))
    assert [r for r in dropin_flow.read_log(str(tmp_path / 'a.log'))] == [r for r in dropin_flow.read_log(str(tmp_path / 'b.log'))]
    assert a.scheduler.real_time == b.scheduler.real_time
    a.close(); b.close()


def test_run_scenario_driver(tmp_path):
    """The reference's own driver function (scenario_models.py:443-500): parameters, model, load + seed, step loop to
    2021-06-01, log file, pickled model -- and the same trajectory as stepping the model by hand."""
    import glob
    import pickle
    from datetime import timedelta
    from covid19_abm import scenario_models
    root = str(tmp_path / 'data')
    os.makedirs(root)
    dropin_flow.build_world(root, n_agents=20_000)
    start = datetime(2021, 5, 22, 0)
    scenario_models.UnmitigatedScenario.philox_seed = 4711
    try:
        np.random.seed(5)
        model = scenario_models.run_scenario(scenario_models.UnmitigatedScenario, 'Unmitigated', '3', 2.6, 100, 150,
                                             start_date=start, timestep=timedelta(hours=4))
        assert model.scheduler.real_time == datetime(2021, 6, 1, 4)          # first step past the end date
        logs = glob.glob(os.path.join(root, 'logs', 'model_log_file_Unmitigated_03_R2.6_samp100_seed150.*.log'))
        dumps = glob.glob(os.path.join(root, 'logs', 'model_dump_file_Unmitigated_03_R2.6_samp100_seed150.*.pickle'))
        assert len(logs) == 1 and len(dumps) == 1
        rows = dropin_flow.read_log(logs[0])
        assert len(rows) == 10 and rows[0]['date'] == '2021-05-22T08:00:00' and rows[-1]['infected_count'] > 300
        from covid19_abm import data_plotting                      # the notebooks' reader of that file (§8 f1)
        df = data_plotting.load_model_logs_df(logs[0])
        assert len(df) == 10 and int(df['new_cases'].sum()) == rows[-1]['infected_count']
        with open(dumps[0], 'rb') as fl:
            again = pickle.load(fl)
        assert again._sim is None and np.array_equal(again.epidemic_state, model.epidemic_state)
        # by hand: same numpy seed, same Philox key, same number of sub-steps
        np.random.seed(5)
        p = dropin_flow.make_params(R0=2.6, seed_infected=150)
        by_hand = scenario_models.UnmitigatedScenario(p, model_log_file=None)
        p.SIMULATION_START_DATE = start
        by_hand.scheduler.real_time = start
        by_hand.load_agents(p.data_file_name, size=None, infect_num=p.SEED_INFECT_NUM)
        for _ in range(model.scheduler.steps):
            by_hand.step()
        assert np.array_equal(by_hand.epidemic_state, model.epidemic_state)
        assert np.array_equal(by_hand.event_ticks('date_recovered'), model.event_ticks('date_recovered'))
        by_hand.close(); model.close()
    finally:
        scenario_models.UnmitigatedScenario.philox_seed = None


def test_dropin_at_scale_equals_engine_level_run(tmp_path):
    """1 M agents through the drop-in (shuffled census, lockdown hooks, seeding, 20 days in one call) give the day log
    and the per-person state of an engine-level Simulation fed the same canonical population, spec, seeds and key."""
    from abm_b200 import tables
    from abm_b200.engine import COUNTER_NAMES, Simulation
    root = str(tmp_path / 'data')
    os.makedirs(root)
    dropin_flow.build_world(root, n_agents=1_000_000)
    m = dropin_flow.make_model('ContinuedLockdownScenario', datetime(2020, 9, 1, 0), log_file=str(tmp_path / 'm.log'))
    seeded = np.sort(m._canon_of_person[np.nonzero(m.epidemic_state > 0)[0]])
    m.run_days(20)
    sim = Simulation(m._spec_obj, m._pop, seed=m._philox_seed, tables=tables.build_tables(m._spec_obj), max_days=30)
    sim.seed_infections(seeded)
    sim.step(6 * 20)
    rows = dropin_flow.read_log(str(tmp_path / 'm.log'))
    want = sim.log_dicts()
    assert len(rows) == len(want) == 20
    for g, w in zip(rows, want):
        assert g['date'] == w['date']
        for name in COUNTER_NAMES:
            assert g[name] == w[name], (name, g['date'])
    st = sim.state()
    cp = m._canon_of_person
    assert np.array_equal(m.epidemic_state.astype(np.int64), st['epidemic_state'].astype(np.int64)[cp])
    assert np.array_equal(m.current_district_ids, st['current_district_ids'][cp])
    assert np.array_equal(m.event_ticks('date_recovered'), st['date_recovered'][cp])
    n = m.person_ids.shape[0]
    assert n == 1_000_000 and rows[-1]['infected_count'] > 300
    exact = int(((st['epidemic_state'] >= 1) & (st['epidemic_state'] < 3)).sum())
    assert 0 < m.active_infections() <= exact          # running counters: the last draw's infections come with the next sub-step
    sim.close(); m.close()
