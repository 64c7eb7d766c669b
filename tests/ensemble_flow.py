"""Tier-2 parity helper: run ensemble members of a scenario through the drop-in package exactly as
tests/golden/make_ref_ensemble.py drove the real reference (same synthetic world, same numpy seed per member, same
driver sequence) and compare daily curves with the stored reference ensemble."""
import json
import os
from datetime import datetime, timedelta

import numpy as np

from abm_b200 import synth
from util import GOLDEN

POP_SEED = 2020
CURVES = ('current_contagious_cases', 'current_critical_cases', 'died_count', 'infected_count')


def load_reference(case):
    with open(os.path.join(GOLDEN, f'ref_ensemble_{case}.json')) as fl:
        return json.load(fl)


def build_world(root, ref):
    spec = synth.make_spec(n_districts=ref['n_districts'], R0=ref['R0'], seed=POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(ref['n_agents'], spec, seed=POP_SEED, scenario_flags=True)
    if ref['case'] in ('selective_lockdown_100k', 'open_manufacturing_schools_100k', 'open_manufacturing_schools_100k_120d'):
        # the census columns the sector / school scenarios read, exactly as tests/golden/make_ref_ensemble.py adds them
        names = np.array(synth.district_names(ref['n_districts']))
        pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
        pop.extras['school_id_district'] = np.where(pop.extras['school_goers'] == 1, names[pop.district], '')
    if ref['case'] == 'open_mining_100k':
        pop.extras['mining_district_id'] = synth.mining_camp_column(pop, ref['n_districts'], POP_SEED, 0.15)
    if ref['case'] == 'accelerated_schools_100k':
        pop.extras.update(synth.school_census_columns(pop, ref['n_districts']))
    cases = synth.seed_case_counts(pop, spec, seed=POP_SEED)
    synth.write_data_dir(root, spec, pop, seed_cases=cases, seed=0)
    os.environ['COVID19_ABM_DATA_DIR'] = root
    return spec, pop


def run_member(ref, member, log_file, days=None):
    """One member: returns {counter: [value per logged day]}."""
    from covid19_abm import scenario_models
    from covid19_abm.dir_manager import get_data_dir
    from covid19_abm.params import ParamsConfig
    days = ref['days'] if days is None else days
    start = datetime(2020, 9, 1, 0)
    np.random.seed(1000 + member)
    normal = get_data_dir('raw', 'final_close_interaction_matrix_normal.xlsx')
    p = ParamsConfig(
        district='new', data_sample_size=100, R0=ref['R0'], normal_interaction_matrix_file=normal,
        lockdown_interaction_matrix_file=normal,      # the reference ensemble read the same workbook for both
        stay_duration_file=get_data_dir('preprocessed', 'mobility', 'weekday_mobility_duration_count_df-new-district.pickle'),
        transition_probability_file=get_data_dir('preprocessed', 'mobility',
                                                 'daily_region_transition_probability-new-district-pre-lockdown.csv'),
        timestep=timedelta(hours=4))
    p.set_new_district_seed(seed_infected=ref['infect_num'])
    if os.path.exists(log_file):
        os.remove(log_file)
    m = getattr(scenario_models, ref['scenario'])(p, model_log_file=log_file, individual_log_file=None)
    m.log_flush_days = 1000
    p.SIMULATION_START_DATE = start
    m.scheduler.real_time = start
    m.load_agents(p.data_file_name, size=None, infect_num=p.SEED_INFECT_NUM)
    end = start + timedelta(days=days, hours=8)
    n = 0
    while m.scheduler.real_time + n * m.step_timedelta <= end:
        n += 1
    m.run_days(n // 6)
    for _ in range(n % 6):
        m.step()
    m.log_model_output()
    m.close()
    rows = []
    with open(log_file) as fl:
        for line in fl:
            rows.append(json.loads(line.split(' $$ ', 1)[1]))
    return {k: [r[k] for r in rows] for k in rows[0] if k != 'date'}, [r['date'] for r in rows]


def compare(ref, members):
    """Per curve: share of days on which the mean of `members` lies inside the reference's central 95 % band, and
    the largest |difference of means| in units of the combined standard error."""
    out = {}
    for key in CURVES:
        r = np.array(ref['counters'][key], dtype=np.float64)                     # [members, days]
        g = np.array([m[key] for m in members], dtype=np.float64)
        days = min(r.shape[1], g.shape[1])
        r, g = r[:, :days], g[:, :days]
        lo, hi = np.percentile(r, 2.5, axis=0), np.percentile(r, 97.5, axis=0)
        gm, rm = g.mean(axis=0), r.mean(axis=0)
        inside = (gm >= lo) & (gm <= hi)
        se = np.sqrt(r.var(axis=0, ddof=1) / r.shape[0] + g.var(axis=0, ddof=1) / g.shape[0])
        z = np.abs(gm - rm) / np.maximum(se, 0.5)
        rel = np.abs(gm - rm) / np.maximum(rm, 1.0)
        out[key] = dict(inside=float(inside.mean()), max_z=float(z.max()), final_ref=float(rm[-1]), final_got=float(gm[-1]),
                        max_rel=float(rel[rm > 50].max()) if (rm > 50).any() else 0.0)
    return out
