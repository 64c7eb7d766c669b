"""Shared helpers for the parity tests."""
import json
import os
from datetime import datetime

import numpy as np

from abm_b200 import synth, tables
from abm_b200.spec import ModelSpec, Population

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STATE_KEYS = ['epidemic_state', 'clinical_state', 'infected_symptomatic_status', 'current_location_ids',
              'current_district_ids', 'return_district_at', 'date_infected', 'date_start_contagious',
              'date_start_symptomatic', 'date_recovered', 'date_start_hospitalized', 'date_end_hospitalized',
              'date_start_critical', 'date_end_critical', 'date_died', 'infected_at_district_ids',
              'infected_at_location_ids']


def load_fixture(name):
    """Inputs + reference snapshots of tests/golden/ref_functions_<name>.npz."""
    z = np.load(os.path.join(GOLDEN, f'ref_functions_{name}.npz'))
    start = datetime(*[int(v) for v in z['start']])
    hw = z['spec_hw_risk']
    spec = ModelSpec(n_districts=int(z['n_districts']), R0=float(z['spec_R0']), step_minutes=int(z['spec_step_minutes']),
                     start_datetime=start, interaction_matrix=z['spec_interaction_matrix'],
                     interaction_sizes=z['spec_interaction_sizes'], od_cum=z['spec_od_cum'],
                     stay_mean=z['spec_stay_mean'], stay_std=z['spec_stay_std'],
                     hw_risk=hw if hw.size else None, mild_symptom_move_prob=float(z['spec_mild']),
                     n_extra_pools=int(z['spec_n_extra_pools']) if 'spec_n_extra_pools' in z.files else 0)
    pop = Population(district=z['pop_district'], household=z['pop_household'], age=z['pop_age'], status=z['pop_status'],
                     econ_kind=z['pop_econ_kind'], district_mover=z['pop_district_mover'],
                     move_prob_weekday=z['pop_move_prob_weekday'], move_prob_other=z['pop_move_prob_other'],
                     econ_pool=z['pop_econ_pool'] if 'pop_econ_pool' in z.files else None)      # schools as pools
    snaps = []
    for i, k in enumerate(z['snap_steps']):
        snaps.append((int(k), {key: z[f'snap{i}_{key}'] for key in STATE_KEYS}))
    return dict(spec=spec, pop=pop, seeds=z['seeds'], philox_seed=int(z['philox_seed']), n_steps=int(z['n_steps']),
                snaps=snaps, ref_log=json.loads(str(z['ref_log'])))


def oracle_params(spec):
    return dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob)


def assert_state_equal(got, want, where=''):
    for k in STATE_KEYS:
        g, w = np.asarray(got[k]).astype(np.int64), np.asarray(want[k]).astype(np.int64)
        if not np.array_equal(g, w):
            bad = np.nonzero(g != w)[0]
            raise AssertionError(f'{where}: {k} differs at {bad.size} agents, first {bad[:8]}: got {g[bad[:8]]} want {w[bad[:8]]}')


def small_case(n_agents=30_000, n_districts=8, seed=5, R0=2.6, infect_num=300, **spec_kw):
    spec = synth.make_spec(n_districts=n_districts, R0=R0, seed=seed, **spec_kw)
    pop = synth.make_population(n_agents, spec, seed=seed)
    seeds = synth.pick_seeds(pop, spec, infect_num=infect_num, seed=seed)
    return spec, pop, seeds, tables.build_tables(spec)
