"""Generate tests/golden/ref_mirror_host_more.npz -- the scenario classes make_ref_mirror_fixtures.py does not cover
(the two improved hand-washing variants, the interaction-matrix sensitivity run, open mining, eased lockdown with
schools, open schools seeded among children, open manufacturing + schools), on a census that carries the school and
mining-camp columns.  Same content per scenario as ref_mirror_host.npz: what the REAL reference's
ParamsConfig / load_agents / scenario hooks leave on the host under a fixed numpy seed.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_mirror_fixtures_more.py
"""
import os
import sys
import tempfile
from datetime import datetime

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

SCENARIOS = ['HandWashingRiskImproved1Scenario', 'HandWashingRiskImproved2Scenario', 'InteractionSensitivityScenario',
             'OpenMiningScenario', 'EasedOpenSchoolsScenario', 'OpenSchoolsSeedKidsScenario',
             'OpenManufacturingAndSchoolsScenario']
N_AGENTS, N_DISTRICTS, POP_SEED, NP_SEED = 6000, 60, 79, 2468


def snapshot(m, out, pre):
    p = m.params
    g = np.asarray
    out[pre + 'district_mover'] = g(m.district_mover).astype(np.int8)
    out[pre + 'move_weekday'] = g(m.economic_status_weekday_movement_probability, dtype=np.float64)
    out[pre + 'move_other'] = g(m.economic_status_other_day_movement_probability, dtype=np.float64)
    out[pre + 'econ_loc'] = g(m.economic_activity_location_ids).astype(np.int64)
    out[pre + 'household_ids'] = g(m.household_ids).astype(np.int64)
    out[pre + 'district_ids'] = g(m.district_ids).astype(np.int64)
    out[pre + 'economic_status_ids'] = g(m.economic_status_ids).astype(np.int64)
    out[pre + 'lockdown_flags'] = np.array([int(m.lockdown), int(m.blocked), int(p.lockdown), int(p.blocked)])
    out[pre + 'scenario'] = np.array(p.SCENARIO)
    out[pre + 'mild'] = np.array(float(p.MILD_SYMPTOM_MOVEMENT_PROBABILITY))
    out[pre + 'severe_disease_risk'] = g(m.severe_disease_risk, dtype=np.float64)
    if hasattr(m, 'hw_risk'):
        out[pre + 'hw_risk'] = g(m.hw_risk, dtype=np.float64)
    out[pre + 'interaction_matrix'] = g(p.ECONOMIC_STATUS_INTERACTION_MATRIX_VALUES, dtype=np.float64)
    out[pre + 'interaction_sizes'] = np.array([p.ECONOMIC_STATUS_INTERACTION_SIZE_MAP[p.ECON_STAT_ID_TO_NAME[i]]
                                              for i in range(len(p.ECON_STAT_ID_TO_NAME))], dtype=np.float64)
    out[pre + 'seed_age_window'] = np.array([p.SEED_INFECT_AGE_MIN, p.SEED_INFECT_AGE_MAX])
    out[pre + 'n_households'] = np.array(len(p.HOUSEHOLD_ID_TO_NAME))
    out[pre + 'first_household_names'] = np.array([p.HOUSEHOLD_ID_TO_NAME[i] for i in sorted(p.HOUSEHOLD_ID_TO_NAME)[:20]])


def main():
    import ref_harness as rh
    from abm_b200 import synth
    spec = synth.make_spec(n_districts=N_DISTRICTS, R0=1.9, seed=POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(N_AGENTS, spec, seed=POP_SEED, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    cases = synth.seed_case_counts(pop, spec, seed=POP_SEED)
    A = spec.n_status
    sens = np.zeros((A, A))                       # the workbook abm_b200.synth.write_data_dir writes for this scenario
    sens[:A - 1, :A - 1] = 1.0 / (A - 1)
    sens[A - 1, A - 1] = 1.0
    normal = rh._fake_read_excel(spec.status_names, spec.interaction_matrix, spec.interaction_sizes)
    sensitivity = rh._fake_read_excel(spec.status_names, sens, np.full(A, 11))
    out = {}
    for name in SCENARIOS:
        w = rh.RefWorld(spec, pop, seed_cases=cases)
        try:
            pd.read_excel = lambda fname, **kw: (sensitivity if 'sensitivity' in str(fname) else normal)(fname, **kw)
            tmp = tempfile.mkdtemp()
            census = synth.write_data_dir(tmp, spec, pop, seed_cases=cases, seed=POP_SEED, school_columns=True,
                                          mining_columns=True)
            os.replace(census, w.census_file)
            for rel in ('preprocessed/risk/hw_and_severe_disease_risk.csv',
                        'preprocessed/mobility/intra_district_decreased_mobility_rates.csv'):
                os.replace(os.path.join(tmp, rel), os.path.join(w.tmp, rel))
            np.random.seed(NP_SEED)
            m = w.make_model(name, seed_infected=25, load=True, infect=False)
            snapshot(m, out, name + '/')
            print(name, str(out[name + '/scenario']), 'movers', int(out[name + '/district_mover'].sum()),
                  'mean weekday prob %.4f' % out[name + '/move_weekday'].mean(), 'households', int(out[name + '/n_households']))
        finally:
            w.close()
    np.savez_compressed(os.path.join(HERE, 'ref_mirror_host_more.npz'), **out)


if __name__ == '__main__':
    main()
