"""Generate tests/golden/ref_mirror_host.npz -- what the REAL reference's ParamsConfig / Country.load_agents build
on the host (derived tables, id maps, per-agent vectors after the scenario hooks), for the drop-in mirror
(covid19-agent-based-model_b200/covid19_abm) to be compared against.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_mirror_fixtures.py
For every scenario the global numpy RNG is seeded, the reference's params + scenario class are built and
`load_agents(..., infect_num=None)` is run (no seeding: the disease course is sampled on the device in the mirror).
"""
import os
import sys
from datetime import datetime

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

SCENARIOS = ['UnmitigatedScenario', 'ContinuedLockdownScenario', 'EasedLockdownScenario', 'LockdownGreatestMobilityScenario',
             'BlockGreatestMobilityScenario', 'IsolateSymptomaticScenario', 'IsolateVulnerableScenario',
             'HandWashingRiskScenario', 'OpenSchoolsScenario', 'OpenManufacturingScenario']
N_AGENTS, N_DISTRICTS, POP_SEED, NP_SEED = 6000, 60, 77, 1234


def world():
    import ref_harness as rh
    from abm_b200 import synth
    spec = synth.make_spec(n_districts=N_DISTRICTS, R0=1.9, seed=POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(N_AGENTS, spec, seed=POP_SEED, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    cases = synth.seed_case_counts(pop, spec, seed=POP_SEED)
    return rh, synth, spec, pop, cases


def main():
    rh, synth, spec, pop, cases = world()
    out = {}
    for name in SCENARIOS:
        w = rh.RefWorld(spec, pop, seed_cases=cases)
        try:
            # same synthetic risk table the mirror test will use (the shipped one stays in /root/reference)
            import tempfile
            tmp = tempfile.mkdtemp()
            synth.write_data_dir(tmp, spec, pop, seed_cases=cases, seed=POP_SEED)
            os.replace(os.path.join(tmp, 'preprocessed/risk/hw_and_severe_disease_risk.csv'),
                       os.path.join(w.tmp, 'preprocessed/risk/hw_and_severe_disease_risk.csv'))
            os.replace(os.path.join(tmp, 'preprocessed/mobility/intra_district_decreased_mobility_rates.csv'),
                       os.path.join(w.tmp, 'preprocessed/mobility/intra_district_decreased_mobility_rates.csv'))
            np.random.seed(NP_SEED)
            m = w.make_model(name, seed_infected=25, load=True, infect=False)
            p = m.params
            pre = name + '/'
            out[pre + 'district_mover'] = np.asarray(m.district_mover).astype(np.int8)
            out[pre + 'move_weekday'] = np.asarray(m.economic_status_weekday_movement_probability, dtype=np.float64)
            out[pre + 'move_other'] = np.asarray(m.economic_status_other_day_movement_probability, dtype=np.float64)
            out[pre + 'econ_loc'] = np.asarray(m.economic_activity_location_ids).astype(np.int64)
            out[pre + 'household_ids'] = np.asarray(m.household_ids).astype(np.int64)
            out[pre + 'district_ids'] = np.asarray(m.district_ids).astype(np.int64)
            out[pre + 'economic_status_ids'] = np.asarray(m.economic_status_ids).astype(np.int64)
            out[pre + 'lockdown_flags'] = np.array([int(m.lockdown), int(m.blocked), int(p.lockdown), int(p.blocked)])
            out[pre + 'scenario'] = np.array(p.SCENARIO)
            out[pre + 'mild'] = np.array(float(p.MILD_SYMPTOM_MOVEMENT_PROBABILITY))
            out[pre + 'severe_disease_risk'] = np.asarray(m.severe_disease_risk, dtype=np.float64)
            if hasattr(m, 'hw_risk'):
                out[pre + 'hw_risk'] = np.asarray(m.hw_risk, dtype=np.float64)
            out[pre + 'interaction_matrix'] = np.asarray(p.ECONOMIC_STATUS_INTERACTION_MATRIX_VALUES, dtype=np.float64)
            if name == 'UnmitigatedScenario':
                ages = list(range(100))
                out['params/r_asym'] = np.asarray(p.AGE_ASYMPTOMATIC_INFECTION_RATE_VALUES, dtype=np.float64)
                out['params/r_sym'] = np.asarray(p.AGE_SYMPTOMATIC_INFECTION_RATE_VALUES, dtype=np.float64)
                out['params/p_hosp'] = np.array([p.AGE_HOSPITALIZATION_PROBABILITY[a] for a in ages])
                out['params/p_crit'] = np.array([p.AGE_CRITICAL_CARE_PROBABILITY[a] for a in ages])
                out['params/p_hfat'] = np.array([p.AGE_HOSPITALIZATION_FATALITY_PROBABILITY[a] for a in ages])
                out['params/p_susc'] = np.array([p.AGE_SUSCEPTIBILITY_PROBABILITY[a] for a in ages])
                out['params/scalars'] = np.array([p.ASYMPTOMATIC_INFECTION_RATE, p.SYMPTOMATIC_INFECTION_RATE,
                                                  p.MEAN_HOSPITALIZATION_RATE])
                out['params/district_ids'] = np.array(p.DISTRICT_IDS)
                out['params/od'] = p.DAILY_DISTRICT_TRANSITION_PROBABILITY.values.astype(np.float64)
                out['params/seed_districts'] = np.asarray(p.SEED_INFECT_DISTRICT_IDS).astype(np.int64)
                out['params/seed_prob'] = np.array([p.DISTRICT_ID_INFECTED_PROB[d] for d in p.SEED_INFECT_DISTRICT_IDS])
                out['params/moving_status'] = np.array(sorted(p.DISTRICT_MOVING_ECONOMIC_STATUS))
                out['params/gamma'] = np.array(p.get_gamma_shape_scale(30.0, 12.0))
            print(name, 'movers', int(out[pre + 'district_mover'].sum()), 'mean weekday prob', out[pre + 'move_weekday'].mean())
        finally:
            w.close()
    np.savez_compressed(os.path.join(HERE, 'ref_mirror_host.npz'), **out)


if __name__ == '__main__':
    main()
