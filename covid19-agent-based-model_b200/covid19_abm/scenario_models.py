"""The 20 scenario classes and `run_scenario` of the reference
(/root/reference/src/covid19_abm/scenario_models.py:9-500), on the GPU-backed `Country`.

A scenario = a `ParamsConfig.scenario_*()` mutator applied in the constructor + two hooks around
`Country.initialize_agent_vectors` in `load_agents` (base_model.py:773-775):
    scenario_data_preprocessing(df)    edits the census DataFrame before the vectors are built
    scenario_data_postprocessing(df)   edits the host vectors (district_mover, movement probabilities,
                                       economic_activity_location_ids, hw_risk ...) before the device store exists
Each class below is declared by those three ingredients; names, constructor signature and hook names
are the reference's, so `scenarioClass(params, model_log_file=..., individual_log_file=...)` and
subclassing keep working.
"""
from datetime import datetime, timedelta

import numpy as np

from covid19_abm.base_model import Country


def _declare(name, params_method, expect, post=None, pre=None, school=False, doc=''):
    """Build one scenario class.  `post(model, df)` / `pre(model, df)` are the hook bodies."""

    def __init__(self, params, model_log_file=None, individual_log_file=None):
        Country.__init__(self, params, model_log_file=model_log_file, individual_log_file=individual_log_file)
        if school:
            self.is_school_scenario = True
        if params_method is not None:
            getattr(self.params, params_method)()

    def scenario_data_preprocessing(self, df):
        if pre is not None:
            pre(self, df)

    def scenario_data_postprocessing(self, df):
        assert self.params.SCENARIO == expect, (self.params.SCENARIO, expect)
        if post is not None:
            post(self, df)

    cls = type(name, (Country,), dict(__init__=__init__, scenario_data_preprocessing=scenario_data_preprocessing,
                                      scenario_data_postprocessing=scenario_data_postprocessing, __doc__=doc,
                                      __module__=__name__))
    return cls


# ---------------------------------------------------------------------------------------------------- hook bodies
def _district_risk_vectors(model, df):
    """Per-district hand-washing and severe-disease multipliers in district-id order (scenario_models.py:35-43).
    `severe_disease_risk` is overwritten with ones later in load_agents (quirk Q2); `hw_risk` survives."""
    p = model.params
    model.hw_risk = np.array([p.DISTRICT_HW_RISK.get(d) for d in p.DISTRICT_IDS])
    model.severe_disease_risk = np.array([p.DISTRICT_SEVERE_DISEASE_RISK.get(d) for d in p.DISTRICT_IDS])


def _check_isolation(model, df):
    assert model.params.MILD_SYMPTOM_MOVEMENT_PROBABILITY < 1


def _shield_vulnerable(model, df):
    """People of VULNERABLE_AGE and above neither travel nor leave the house (scenario_models.py:138-149)."""
    p = model.params
    old = model.age >= p.VULNERABLE_AGE
    model.district_mover = model.DISTRICT_MOVER_TRUE * model._may_travel((model.age >= p.DISTRICT_MOVEMENT_ALLOWED_AGE) & ~old)
    model.economic_activity_location_ids[old] = model.current_location_ids[old]


def _block(model, df):
    model.set_blocked_movers_and_movement_probabilities()


def _lockdown_everyone(model, df):
    model.set_lockdown_movers_and_movement_probabilities(unrestricted_ids=None)


def _miners_to_camps(model, df):
    """Miners live and work in their mining camp: the camp id replaces household and work place (:234-235)."""
    miner = df['mining_district_id'] != ''
    df.loc[miner, 'economic_activity_location_id'] = df.loc[miner, 'mining_district_id']
    df.loc[miner, 'household_id'] = df.loc[miner, 'mining_district_id']


def _reopen_mining(model, df):
    mining_ids = df[~df['household_id'].str.startswith('h_')]['person_id'].values
    model.set_lockdown_movers_and_movement_probabilities(unrestricted_ids=mining_ids)
    model.district_mover[mining_ids] = model.DISTRICT_MOVER_FALSE


def _school_goers(df, strict=False):
    if 'school_id_district' in df:
        return df[df['school_id_district'] != '']['person_id'].values
    if not strict and 'school_goers' in df:
        return df[df['school_goers'] == 1]['person_id'].values
    if strict:
        raise KeyError('school_id_district')
    raise ValueError('Column `school_id_district` or `school_goers` not found!')


def _reopen(select):
    def post(model, df):
        model.setup_selective_movement_restriction_scenarios(unrestricted_ids=select(df), set_lockdown=True)
    return post


def _manufacturing(df):
    return df[df['manufacturing_workers'].notnull()]['person_id'].values


def _manufacturing_and_schools(df):
    return np.union1d(_manufacturing(df), _school_goers(df, strict=True))


def _phase(test):
    def post(model, df):
        if 'phase' not in df:
            raise ValueError('Column `phase` not found!')
        ids = df[test(df['phase'])]['person_id'].values
        model.setup_selective_movement_restriction_scenarios(unrestricted_ids=ids, set_lockdown=True)
        model.set_school_params(df)
    return post


# ---------------------------------------------------------------------------------------------------- the classes
UnmitigatedScenario = _declare('UnmitigatedScenario', None, 'UNMITIGATED', doc='No intervention (:9-19).')
HandWashingRiskScenario = _declare('HandWashingRiskScenario', 'scenario_handwashing_risk', 'HANDWASHING_RISK',
                                   post=_district_risk_vectors, doc='WASH risk multipliers per district (:22-43).')
HandWashingRiskImproved1Scenario = _declare('HandWashingRiskImproved1Scenario', 'scenario_improved_handwashing_risk_1',
                                            'HANDWASHING_RISK_1', post=_district_risk_vectors, doc=':46-67')
HandWashingRiskImproved2Scenario = _declare('HandWashingRiskImproved2Scenario', 'scenario_improved_handwashing_risk_2',
                                            'HANDWASHING_RISK_2', post=_district_risk_vectors, doc=':70-91')
InteractionSensitivityScenario = _declare('InteractionSensitivityScenario', 'scenario_test_interaction_matrix_sensitivity',
                                          'INTERACTION_MATRIX_SENSITIVITY', doc=':94-106')
IsolateSymptomaticScenario = _declare('IsolateSymptomaticScenario', 'scenario_isolate_symptomatic_population',
                                      'ISOLATE_SYMPTOMATIC', post=_check_isolation, doc=':109-123')
IsolateVulnerableScenario = _declare('IsolateVulnerableScenario', 'scenario_isolate_vulnerable_groups_in_house',
                                     'ISOLATE_VULNERABLE_HOUSE', post=_shield_vulnerable, doc=':126-149')
BlockGreatestMobilityScenario = _declare('BlockGreatestMobilityScenario', 'scenario_block_new_district_greatest_movement',
                                         'BLOCK_GREATEST_NEW_DIST', post=_block, doc=':152-167')
LockdownGreatestMobilityScenario = _declare('LockdownGreatestMobilityScenario',
                                            'scenario_lockdown_new_district_greatest_movement', 'LOCKDOWN_GREATEST_NEW_DIST',
                                            post=_lockdown_everyone, doc='Selective lockdown of the high-mobility districts (:170-185).')
ContinuedLockdownScenario = _declare('ContinuedLockdownScenario', 'scenario_continued_all_lockdown', 'CONTINUED_ALL_LOCKDOWN',
                                     post=_lockdown_everyone, doc=':188-203')
EasedLockdownScenario = _declare('EasedLockdownScenario', 'scenario_eased_all_lockdown', 'EASED_ALL_LOCKDOWN',
                                 post=_lockdown_everyone, doc=':206-221')
OpenMiningScenario = _declare('OpenMiningScenario', 'scenario_continued_all_lockdown_open_mining', 'CONTINUED_ALL_LOCKDOWN_MINING',
                              pre=_miners_to_camps, post=_reopen_mining, doc=':224-246')
OpenSchoolsScenario = _declare('OpenSchoolsScenario', 'scenario_continued_all_lockdown_open_schools',
                               'CONTINUED_ALL_LOCKDOWN_SCHOOLS', post=_reopen(_school_goers), doc=':249-270')
EasedOpenSchoolsScenario = _declare('EasedOpenSchoolsScenario', 'scenario_eased_all_lockdown_open_schools',
                                    'EASED_ALL_LOCKDOWN_SCHOOLS', post=_reopen(_school_goers), doc=':273-294')
OpenSchoolsSeedKidsScenario = _declare('OpenSchoolsSeedKidsScenario', 'scenario_continued_all_lockdown_open_schools_seed_kids',
                                       'CONTINUED_ALL_LOCKDOWN_SCHOOLS_SEED_KIDS',
                                       post=_reopen(lambda df: _school_goers(df, strict=True)), doc=':296-312')
OpenManufacturingScenario = _declare('OpenManufacturingScenario', 'scenario_continued_all_lockdown_open_manufacturing',
                                     'CONTINUED_ALL_LOCKDOWN_MANUFACTURING', post=_reopen(_manufacturing), doc=':315-331')
OpenManufacturingAndSchoolsScenario = _declare('OpenManufacturingAndSchoolsScenario',
                                               'scenario_continued_all_lockdown_open_manufacturing_schools',
                                               'CONTINUED_ALL_LOCKDOWN_MANUFACTURING_SCHOOLS',
                                               post=_reopen(_manufacturing_and_schools), doc=':334-353')
Phase1GovernmentOpenSchoolsScenario = _declare('Phase1GovernmentOpenSchoolsScenario', 'scenario_phase1_government_open_schools',
                                               'PHASE1_GOVERNMENT_OPEN_SCHOOLS', post=_phase(lambda ph: ph == 1), school=True,
                                               doc=':356-382')
DynamicPhase1GovernmentOpenSchoolsScenario = _declare('DynamicPhase1GovernmentOpenSchoolsScenario',
                                                      'scenario_dynamic_phase1_government_open_schools',
                                                      'DYNAMIC_PHASE1_GOVERNMENT_OPEN_SCHOOLS', post=_phase(lambda ph: ph == 1),
                                                      school=True, doc=':385-415')
AcceleratedGovernmentOpenSchoolsScenario = _declare('AcceleratedGovernmentOpenSchoolsScenario',
                                                    'scenario_accelerated_government_open_schools',
                                                    'ACCELERATED_GOVERNMENT_OPEN_SCHOOLS', post=_phase(lambda ph: ph > 0),
                                                    school=True, doc=':418-440')


def run_scenario(scenarioClass, scenario_name, sim_fname, R0, sample_size, seed_num, start_date=None,
                 timestep=timedelta(hours=4), scaled_mobility=False):
    """One full run, driven like the reference's (scenario_models.py:443-500): build the parameters, the model,
    load and seed the agents, step until 2021-06-01 or until nobody is exposed or contagious, pickle the model."""
    import pickle
    from covid19_abm.dir_manager import get_data_dir
    from covid19_abm.params import ParamsConfig

    sim_fname = sim_fname.zfill(2)
    tag = 'scaled_' if scaled_mobility else ''
    sim_fname = f'{tag}{scenario_name}_{sim_fname}_R{R0}_samp{sample_size}_seed{seed_num}'
    suffix = '-scaled' if scaled_mobility else ''
    stay_duration_file = f'weekday_mobility_duration_count_df-new-district{suffix}.pickle'
    transition_probability_file = ('daily_region_transition_probability-new-district-scaled.csv' if scaled_mobility
                                   else 'daily_region_transition_probability-new-district-pre-lockdown.csv')
    now = datetime.now().isoformat()
    model_log_file = get_data_dir('logs', f'model_log_file_{sim_fname}.{now}.log')
    individual_log_file = get_data_dir('logs', f'individual_log_file_{sim_fname}.{now}.log')

    params = ParamsConfig(
        district='new', data_sample_size=sample_size, R0=R0,
        normal_interaction_matrix_file=get_data_dir('raw', 'final_close_interaction_matrix_normal.xlsx'),
        lockdown_interaction_matrix_file=get_data_dir('raw', 'final_close_interaction_matrix_lockdown.xlsx'),
        stay_duration_file=get_data_dir('preprocessed', 'mobility', stay_duration_file),
        transition_probability_file=get_data_dir('preprocessed', 'mobility', transition_probability_file),
        timestep=timestep)
    params.set_new_district_seed(seed_infected=seed_num)
    model = scenarioClass(params, model_log_file=model_log_file, individual_log_file=individual_log_file)
    if start_date is not None:
        params.SIMULATION_START_DATE = start_date
        model.scheduler.real_time = params.SIMULATION_START_DATE
    model.load_agents(params.data_file_name, size=None, infect_num=params.SEED_INFECT_NUM)

    end_date = datetime(2021, 6, 1)
    while model.scheduler.real_time <= end_date:
        model.step()
        # the reference tests the epidemic_state vector after every sub-step; the same count comes from the device
        # counters without pulling the state of every agent over PCIe
        if model.active_infections() == 0:
            break

    model_dump_file = get_data_dir('logs', f'model_dump_file_{sim_fname}.{now}.pickle')
    # several ranks (torchrun): gathering the vectors for the pickle is a collective, so every rank serialises the model;
    # rank 0 writes the one file
    payload = pickle.dumps(model)
    if getattr(model._sim, 'rank', 0) == 0:
        with open(model_dump_file, 'wb') as fl:
            fl.write(payload)
    return model
