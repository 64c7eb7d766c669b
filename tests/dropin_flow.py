"""Shared body of the drop-in flow tests: the reference's own driver sequence (scenario_models.py:471-495) through
the mirror package, checked against an independent oracle run.  Used with the real CUDA engine (-m gpu) and with the
oracle-backed test double (CPU)."""
import json
import os
import pickle
from datetime import datetime, timedelta

import numpy as np

from abm_b200 import synth, tables
from oracle.abm_oracle import OracleModel
from util import STATE_KEYS

T_NEVER = 2 ** 31 - 1


def build_world(root, n_agents=20_000, n_districts=60, seed=7, R0=2.6, school_columns=False, shuffle=True,
                mining_columns=False, mining_share=0.04):
    spec = synth.make_spec(n_districts=n_districts, R0=R0, seed=seed, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(n_agents, spec, seed=seed, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    census = synth.write_data_dir(root, spec, pop, seed=seed, school_columns=school_columns, mining_columns=mining_columns,
                                  mining_share=mining_share)
    if shuffle:      # person order != canonical (district, household) order, as in a real census extract
        import pandas as pd
        df = pd.read_pickle(census)
        df = df.sample(frac=1.0, random_state=seed).reset_index(drop=True)
        df['person_id'] = np.arange(df.shape[0])
        df.to_pickle(census)
    os.environ['COVID19_ABM_DATA_DIR'] = root
    return spec, pop


def make_params(R0=2.6, seed_infected=150):
    from covid19_abm.dir_manager import get_data_dir
    from covid19_abm.params import ParamsConfig
    p = ParamsConfig(
        district='new', data_sample_size=100, R0=R0,
        normal_interaction_matrix_file=get_data_dir('raw', 'final_close_interaction_matrix_normal.xlsx'),
        lockdown_interaction_matrix_file=get_data_dir('raw', 'final_close_interaction_matrix_lockdown.xlsx'),
        stay_duration_file=get_data_dir('preprocessed', 'mobility', 'weekday_mobility_duration_count_df-new-district.pickle'),
        transition_probability_file=get_data_dir('preprocessed', 'mobility',
                                                 'daily_region_transition_probability-new-district-pre-lockdown.csv'),
        timestep=timedelta(hours=4))
    p.set_new_district_seed(seed_infected=seed_infected)
    return p


def make_model(scenario_name, start, log_file=None, np_seed=5, philox_seed=99, **params_kw):
    from covid19_abm import scenario_models
    np.random.seed(np_seed)
    p = make_params(**params_kw)
    cls = getattr(scenario_models, scenario_name)
    m = cls(p, model_log_file=log_file, individual_log_file=None)
    m.philox_seed = philox_seed
    p.SIMULATION_START_DATE = start
    m.scheduler.real_time = start
    m.load_agents(p.data_file_name, size=None, infect_num=p.SEED_INFECT_NUM)
    return m


def to_ticks(dates, start):
    out = np.full(len(dates), T_NEVER, dtype=np.int64)
    for i, d in enumerate(dates):
        if d != datetime.max:
            out[i] = (d - start) // timedelta(minutes=1)
    return out


def twin_oracle(m):
    """Independent oracle run on the population / spec / seeds the model handed to its engine."""
    spec, pop = m._spec_obj, m._pop
    ora = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob), tables.build_tables(spec),
                      m._philox_seed)
    return ora


def sync_oracle_inputs(ora, m):
    """Mirror a refresh_device(): the host vectors scenario hooks changed, in canonical order."""
    order = m._order
    D, H = ora.D, ora.H
    ora.district_mover = m.district_mover[order].astype(np.int64)
    ora.economic_status_weekday_movement_probability = m.economic_status_weekday_movement_probability[order].astype(np.float64)
    ora.economic_status_other_day_movement_probability = m.economic_status_other_day_movement_probability[order].astype(np.float64)
    kind, pool = m._econ_encoding(order)
    econ = np.where(kind == 0, ora.household_ids, ora.district_ids)
    if pool is not None:
        econ = np.where(kind == 2, D + H + (pool.astype(np.int64) - D), econ)
    ora.economic_activity_location_ids = econ


def assert_model_matches_oracle(m, ora, where=''):
    start = m._spec_obj.start_datetime
    cp = m._canon_of_person
    want = ora.state_dict()
    D, H = ora.D, ora.H
    hh_person = m.household_ids
    for key in STATE_KEYS:
        w = want[key].astype(np.int64)[cp]
        g = getattr(m, key)
        g = to_ticks(g, start) if g.dtype == object else np.asarray(g).astype(np.int64)
        if key in ('current_location_ids', 'infected_at_location_ids'):
            # oracle numbers households D + dense index and schools after them; the model uses the reference's ids
            own_home = (w >= D) & (w < D + H)
            w = np.where(own_home, hh_person, w)
            if m.is_school_scenario:
                first_school = max(m.params.HOUSEHOLD_ID_TO_NAME) + 1
                w = np.where(want[key].astype(np.int64)[cp] >= D + H, want[key].astype(np.int64)[cp] - (D + H) + first_school, w)
        if not np.array_equal(g, w):
            bad = np.nonzero(g != w)[0]
            raise AssertionError(f'{where}: {key} differs for {bad.size} persons, first {bad[:6]}: got {g[bad[:6]]} want {w[bad[:6]]}')


def read_log(path):
    rows = []
    with open(path) as fl:
        for line in fl:
            stamp, payload = line.split(' $$ ', 1)
            rows.append(json.loads(payload))
    return rows


def run_flow(tmp_path, scenario_name, days, start=datetime(2020, 9, 1, 0), check_every=6, pickle_at=None, **world_kw):
    """Drive a scenario like run_scenario does and compare with the twin oracle; returns (model, oracle, log rows)."""
    root = str(tmp_path / 'data')
    os.makedirs(root, exist_ok=True)
    build_world(root, **world_kw)
    log_file = str(tmp_path / 'model.log')
    m = make_model(scenario_name, start, log_file=log_file)
    assert m._sim is not None, 'seeding creates the device store'
    ora = twin_oracle(m)
    seeded = np.nonzero(m.epidemic_state > 0)[0]
    ora.set_epidemic_status(np.sort(m._canon_of_person[seeded]))
    assert_model_matches_oracle(m, ora, 'after seeding')
    n_steps = 6 * days
    for k in range(n_steps):
        m.step()
        # scenario clock hooks run inside m.step() before the device step: replay their effect on the oracle inputs
        if scenario_name.endswith('GovernmentOpenSchoolsScenario'):
            sync_oracle_inputs(ora, m)
        ora.step()
        if (k + 1) % check_every == 0:
            assert_model_matches_oracle(m, ora, f'{scenario_name} sub-step {k + 1}')
        if pickle_at is not None and k + 1 == pickle_at:
            blob = pickle.dumps(m)
            m.close()
            m = pickle.loads(blob)
            assert m._sim is None
    assert_model_matches_oracle(m, ora, f'{scenario_name} end')
    m.log_model_output()
    rows = read_log(log_file)
    assert len(rows) == len(ora.log_rows)
    keys = ['current_contagious_count', 'infected_count', 'current_exposed_cases', 'current_contagious_cases',
            'current_hospitalized_cases', 'current_critical_cases', 'asymptomatic_count', 'symptomatic_count',
            'hospitalized_count', 'critical_count', 'died_count', 'recovered_count']
    for got, want in zip(rows, ora.log_rows):
        assert list(got.keys()) == ['date'] + keys                     # key order of base_model.py:1086-1104
        assert got['date'] == (start + timedelta(minutes=int(want['tick']))).isoformat()
        for key in keys:
            assert got[key] == want[key], (key, got['date'], got[key], want[key])
    # the stop test of run_scenario: exposed + contagious from the running device counters -- exact but for the infections
    # of the last sub-step's draw (applied with the next sub-step), zero exactly when the reference's test is false
    exact = int(((ora.epidemic_state >= 1) & (ora.epidemic_state < 3)).sum())
    assert m.active_infections() <= exact and (m.active_infections() == 0) == (exact == 0)
    return m, ora, rows
