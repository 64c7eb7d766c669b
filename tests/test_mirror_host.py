"""Host side of the drop-in `covid19_abm` package against the REAL reference
(tests/golden/ref_mirror_host.npz, generator tests/golden/make_ref_mirror_fixtures.py): ParamsConfig's derived
tables and, for ten scenario classes, the per-agent vectors `load_agents` + the scenario hooks leave on the host
under the same numpy seed.  No GPU needed: the device store is only created by seeding / stepping."""
import os
import pickle
from datetime import datetime, timedelta

import numpy as np
import pytest

from abm_b200 import synth
from util import GOLDEN

N_AGENTS, N_DISTRICTS, POP_SEED, NP_SEED = 6000, 60, 77, 1234
SCENARIOS = ['UnmitigatedScenario', 'ContinuedLockdownScenario', 'EasedLockdownScenario', 'LockdownGreatestMobilityScenario',
             'BlockGreatestMobilityScenario', 'IsolateSymptomaticScenario', 'IsolateVulnerableScenario',
             'HandWashingRiskScenario', 'OpenSchoolsScenario', 'OpenManufacturingScenario']


@pytest.fixture(scope='module')
def world(tmp_path_factory):
    root = str(tmp_path_factory.mktemp('abm_data'))
    spec = synth.make_spec(n_districts=N_DISTRICTS, R0=1.9, seed=POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(N_AGENTS, spec, seed=POP_SEED, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    cases = synth.seed_case_counts(pop, spec, seed=POP_SEED)
    synth.write_data_dir(root, spec, pop, seed_cases=cases, seed=POP_SEED)
    old = os.environ.get('COVID19_ABM_DATA_DIR')
    os.environ['COVID19_ABM_DATA_DIR'] = root
    yield dict(root=root, spec=spec, pop=pop, ref=np.load(os.path.join(GOLDEN, 'ref_mirror_host.npz')))
    if old is None:
        os.environ.pop('COVID19_ABM_DATA_DIR', None)
    else:
        os.environ['COVID19_ABM_DATA_DIR'] = old


def make_params(seed_infected=25):
    from covid19_abm.dir_manager import get_data_dir
    from covid19_abm.params import ParamsConfig
    p = ParamsConfig(
        district='new', data_sample_size=100, R0=1.9,
        normal_interaction_matrix_file=get_data_dir('raw', 'final_close_interaction_matrix_normal.xlsx'),
        # the fixture's reference run read the same workbook for both (the lockdown one is not shipped)
        lockdown_interaction_matrix_file=get_data_dir('raw', 'final_close_interaction_matrix_normal.xlsx'),
        stay_duration_file=get_data_dir('preprocessed', 'mobility', 'weekday_mobility_duration_count_df-new-district.pickle'),
        transition_probability_file=get_data_dir('preprocessed', 'mobility',
                                                 'daily_region_transition_probability-new-district-pre-lockdown.csv'),
        timestep=timedelta(hours=4))
    p.set_new_district_seed(seed_infected=seed_infected)
    return p


def test_params_derived_tables_match_reference(world):
    ref = world['ref']
    np.random.seed(NP_SEED)
    p = make_params()
    ages = list(range(100))
    assert np.array_equal(p.AGE_ASYMPTOMATIC_INFECTION_RATE_VALUES, ref['params/r_asym'])
    assert np.array_equal(p.AGE_SYMPTOMATIC_INFECTION_RATE_VALUES, ref['params/r_sym'])
    assert np.array_equal([p.AGE_HOSPITALIZATION_PROBABILITY[a] for a in ages], ref['params/p_hosp'])
    assert np.array_equal([p.AGE_CRITICAL_CARE_PROBABILITY[a] for a in ages], ref['params/p_crit'])
    assert np.array_equal([p.AGE_HOSPITALIZATION_FATALITY_PROBABILITY[a] for a in ages], ref['params/p_hfat'])
    assert np.array_equal([p.AGE_SUSCEPTIBILITY_PROBABILITY[a] for a in ages], ref['params/p_susc'])
    assert np.array_equal([p.ASYMPTOMATIC_INFECTION_RATE, p.SYMPTOMATIC_INFECTION_RATE, p.MEAN_HOSPITALIZATION_RATE],
                          ref['params/scalars'])
    assert list(p.DISTRICT_IDS) == list(ref['params/district_ids'])
    assert np.array_equal(p.DAILY_DISTRICT_TRANSITION_PROBABILITY.values, ref['params/od'])
    assert np.array_equal(p.SEED_INFECT_DISTRICT_IDS, ref['params/seed_districts'])
    assert np.array_equal([p.DISTRICT_ID_INFECTED_PROB[d] for d in p.SEED_INFECT_DISTRICT_IDS], ref['params/seed_prob'])
    assert sorted(p.DISTRICT_MOVING_ECONOMIC_STATUS) == list(ref['params/moving_status'])
    assert np.array_equal(p.get_gamma_shape_scale(30.0, 12.0), ref['params/gamma'])
    # analytic known answers (params.py:236-242): R0 / (contagious days * 6 steps per day * contact rate of the age band)
    assert np.isclose(p.AGE_ASYMPTOMATIC_INFECTION_RATE_VALUES.min(), 1.9 / (6.5 * 6 * 13.2094), rtol=1e-12)
    assert np.isclose(p.AGE_SYMPTOMATIC_INFECTION_RATE_VALUES.max(), 1.9 / (7 * 6 * 7.4709), rtol=1e-12)
    assert np.isclose(p.AGE_HOSPITALIZATION_PROBABILITY[50], 0.058 / 0.6, rtol=1e-12)
    assert p.SIMULATION_START_DATE == datetime(2020, 9, 1, 8) and p.SCENARIO == 'UNMITIGATED'
    assert not p.lockdown and not p.blocked


def test_every_reference_scenario_mutator_exists(world):
    np.random.seed(0)
    p = make_params()
    names = [n for n in dir(p) if n.startswith('scenario_')]
    assert len(names) == 29, names          # params.py:379-574
    with pytest.raises(ValueError):
        p.set_lockdown_parameters('nonsense')
    with pytest.raises(ValueError):
        p.set_blocked_parameters('nonsense')
    p.scenario_isolate_symptomatic_population_50pct()
    assert p.MILD_SYMPTOM_MOVEMENT_PROBABILITY == 0.5 and p.SCENARIO == 'ISOLATE_SYMPTOMATIC_50pct'
    p.scenario_continued_all_lockdown_open_schools_seed_kids()
    assert (p.SEED_INFECT_AGE_MIN, p.SEED_INFECT_AGE_MAX, p.lockdown) == (8, 18, True)
    assert p.LOCKDOWN_ALLOWED_PROBABILITY[0] == 0.67 and len(p.LOCKDOWN_DISTRICTS_IDS) == N_DISTRICTS
    p.scenario_eased_all_lockdown()
    assert p.LOCKDOWN_ALLOWED_PROBABILITY[3] == 0.86
    p.scenario_isolate_vulnerable_groups()
    assert p.VULNERABLE_LOCATION == -N_DISTRICTS


def load_scenario(name, np_seed):
    from covid19_abm import scenario_models
    np.random.seed(np_seed)
    p = make_params()
    m = getattr(scenario_models, name)(p, model_log_file=None, individual_log_file=None)
    p.SIMULATION_START_DATE = datetime(2020, 9, 1, 0)
    m.scheduler.real_time = p.SIMULATION_START_DATE
    m.load_agents(p.data_file_name, size=None, infect_num=None)
    return m, p


def check_host_vectors(ref, name, m, p):
    pre = name + '/'
    assert np.array_equal(m.district_mover, ref[pre + 'district_mover'])
    assert np.array_equal(m.economic_status_weekday_movement_probability, ref[pre + 'move_weekday'])
    assert np.array_equal(m.economic_status_other_day_movement_probability, ref[pre + 'move_other'])
    assert np.array_equal(m.economic_activity_location_ids, ref[pre + 'econ_loc'])
    assert np.array_equal(m.household_ids, ref[pre + 'household_ids'])
    assert np.array_equal(m.district_ids, ref[pre + 'district_ids'])
    assert np.array_equal(m.economic_status_ids, ref[pre + 'economic_status_ids'])
    assert [int(m.lockdown), int(m.blocked), int(p.lockdown), int(p.blocked)] == list(ref[pre + 'lockdown_flags'])   # quirk Q1
    assert p.SCENARIO == str(ref[pre + 'scenario']) and p.MILD_SYMPTOM_MOVEMENT_PROBABILITY == float(ref[pre + 'mild'])
    assert np.array_equal(m.severe_disease_risk, ref[pre + 'severe_disease_risk'])                              # quirk Q2
    if pre + 'hw_risk' in ref:
        assert np.allclose(m.hw_risk, ref[pre + 'hw_risk'], rtol=0, atol=0)
    assert np.array_equal(p.ECONOMIC_STATUS_INTERACTION_MATRIX_VALUES, ref[pre + 'interaction_matrix'])
    # initial dynamic vectors as the reference declares them (base_model.py:546-570, 621-625)
    assert m.epidemic_state.dtype == np.float64 and not m.epidemic_state.any()
    assert m.date_infected[0] == datetime.max and m.return_district_at[-1] == datetime.max
    assert np.array_equal(m.current_location_ids, m.household_ids)
    assert m.STATE_DEAD == 4 and m.CLINICAL_RELEASED_OR_DEAD == 3 and m.SYMPTOM_NONE == -1 and m.DISTRICT_MOVER_TRUE == 1
    # the model pickles before any device exists (scenario_models.py:497-500)
    again = pickle.loads(pickle.dumps(m))
    assert np.array_equal(again.district_mover, m.district_mover) and again.params.SCENARIO == p.SCENARIO


@pytest.mark.parametrize('name', SCENARIOS)
def test_load_agents_host_vectors_match_reference(world, name):
    m, p = load_scenario(name, NP_SEED)
    check_host_vectors(world['ref'], name, m, p)


def test_step_without_gpu_fails_loudly(world):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from covid19_abm import scenario_models
    np.random.seed(1)
    p = make_params()
    m = scenario_models.UnmitigatedScenario(p)
    m.scheduler.real_time = datetime(2020, 9, 1, 0)
    m.load_agents(p.data_file_name, size=800, infect_num=None)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.step()
    with pytest.raises(NotImplementedError):
        m.scheduler.go_out()


def test_log_to_file_format(tmp_path):
    from covid19_abm.params import log_to_file
    f = tmp_path / 'x.log'
    log_to_file(str(f), '{"a": 1}', as_log=True, verbose=False)
    log_to_file(str(f), 'plain', as_log=False, verbose=False)
    lines = f.read_text().splitlines()
    assert lines[0].endswith(' $$ {"a": 1}') and lines[1] == 'plain'
