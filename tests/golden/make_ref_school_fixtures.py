"""Generate tests/golden/ref_mirror_school.npz -- host-side behaviour of the REAL reference's three school-phase
scenarios (scenario_models.py:356-440; base_model.py:681-758, 801-805), for the drop-in to be compared against
(tests/test_mirror_school_host.py):

  * what `load_agents` + `set_school_params` leave on the host (school ids, phases, 08:00 destinations, movement
    probabilities, travellers),
  * `get_safe_and_unsafe_districts` on a prescribed epidemic state,
  * `open_schools` and `lockdown_schools` (the latter under a fixed numpy seed).

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_school_fixtures.py
"""
import os
import sys
import tempfile
from datetime import datetime

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

SCENARIOS = ['Phase1GovernmentOpenSchoolsScenario', 'DynamicPhase1GovernmentOpenSchoolsScenario',
             'AcceleratedGovernmentOpenSchoolsScenario']
N_AGENTS, N_DISTRICTS, POP_SEED, NP_SEED = 6000, 60, 78, 4321
LOCKED_DISTRICTS, LOCKED_PHASES, OPEN_PHASES = [3, 7, 11, 40], [1, 2], [1, 2, 3]


def prescribed_sick(n):
    """Who is symptomatic and infected / contagious for the district-safety statistic (same on both sides)."""
    rng = np.random.default_rng(99)
    return rng.random(n) < 0.03, rng.integers(1, 3, n)


def exercise(m, out, pre):
    """The calls both sides make after load_agents; `out[pre + ...]` receives what they leave behind."""
    g = lambda a: np.asarray(a)
    nan_to = lambda a: np.where(np.isnan(g(a).astype(np.float64)), -1, g(a).astype(np.float64)).astype(np.int64)
    out[pre + 'school_phase'] = g(m.school_phase).astype(np.int64)
    out[pre + 'max_school_phase'] = np.array(int(m.max_school_phase))
    out[pre + 'active_school_phases'] = np.array(list(m.active_school_phases))
    out[pre + 'school_ids'] = nan_to(m.school_ids)
    out[pre + 'econ_loc'] = nan_to(m.economic_activity_location_ids)
    out[pre + 'move_weekday'] = g(m.economic_status_weekday_movement_probability).astype(np.float64).copy()
    out[pre + 'move_other'] = g(m.economic_status_other_day_movement_probability).astype(np.float64).copy()
    out[pre + 'district_mover'] = g(m.district_mover).astype(np.int8).copy()
    out[pre + 'n_schools'] = np.array(len(m.params.SCHOOL_ID_TO_NAME))
    out[pre + 'first_school_id'] = np.array(min(m.params.SCHOOL_ID_TO_NAME))
    out[pre + 'school_names'] = np.array([m.params.SCHOOL_ID_TO_NAME[i] for i in sorted(m.params.SCHOOL_ID_TO_NAME)])
    sick, state = prescribed_sick(g(m.district_ids).shape[0])
    m.epidemic_state[sick] = state[sick]
    m.infected_symptomatic_status[sick] = m.SYMPTOM_SYMPTOMATIC
    safe, unsafe = m.get_safe_and_unsafe_districts(quantile_value=0.75)
    out[pre + 'safe'], out[pre + 'unsafe'] = g(safe).astype(np.float64), g(unsafe).astype(np.float64)
    m.open_schools(np.unique(m.district_ids), OPEN_PHASES)
    out[pre + 'open/econ_loc'] = nan_to(m.economic_activity_location_ids)
    out[pre + 'open/move_weekday'] = g(m.economic_status_weekday_movement_probability).astype(np.float64).copy()
    np.random.seed(NP_SEED + 1)
    m.lockdown_schools(LOCKED_DISTRICTS, LOCKED_PHASES)
    out[pre + 'locked/econ_loc'] = nan_to(m.economic_activity_location_ids)
    out[pre + 'locked/move_weekday'] = g(m.economic_status_weekday_movement_probability).astype(np.float64).copy()
    out[pre + 'locked/district_mover'] = g(m.district_mover).astype(np.int8).copy()


def main():
    import ref_harness as rh
    from abm_b200 import synth
    spec = synth.make_spec(n_districts=N_DISTRICTS, R0=1.9, seed=POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(N_AGENTS, spec, seed=POP_SEED, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    cases = synth.seed_case_counts(pop, spec, seed=POP_SEED)
    out = {}
    for name in SCENARIOS:
        w = rh.RefWorld(spec, pop, seed_cases=cases)
        try:
            tmp = tempfile.mkdtemp()
            census = synth.write_data_dir(tmp, spec, pop, seed_cases=cases, seed=POP_SEED, school_columns=True)
            os.replace(census, w.census_file)          # the very census (with school columns) the drop-in's test reads
            for rel in ('preprocessed/risk/hw_and_severe_disease_risk.csv',
                        'preprocessed/mobility/intra_district_decreased_mobility_rates.csv'):
                os.replace(os.path.join(tmp, rel), os.path.join(w.tmp, rel))
            np.random.seed(NP_SEED)
            m = w.make_model(name, seed_infected=25, load=True, infect=False)
            assert m.is_school_scenario
            exercise(m, out, name + '/')
            print(name, 'schools', int(out[name + '/n_schools']), 'pupils at school after open',
                  int((out[name + '/open/econ_loc'] >= out[name + '/first_school_id']).sum()),
                  'after lockdown', int((out[name + '/locked/econ_loc'] >= out[name + '/first_school_id']).sum()),
                  'safe/unsafe', out[name + '/safe'].shape[0], out[name + '/unsafe'].shape[0])
        finally:
            w.close()
    np.savez_compressed(os.path.join(HERE, 'ref_mirror_school.npz'), **out)


if __name__ == '__main__':
    main()
