"""Generate tests/golden/ref_ensemble_<case>.json -- ensemble curves of the REAL reference, own RNG.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_ensemble.py [case ...]

Each member is a full reference run driven exactly like scenario_models.run_scenario
(scenario_models.py:471-495): real ParamsConfig, real scenario class, real load_agents seeding, real
EpidemicScheduler.step with explicit contact sampling, numpy's global RNG seeded per member.  The daily
log lines (base_model.py:1082-1106) of every member are stored so tests can form any band they state.
This is the tier-2 pin: it bounds the one restated piece of the oracle / kernels (pressure-table
infection) and the table-based dwell-time sampling against the real thing.
"""
import json
import multiprocessing as mp
import os
import sys
import time
from datetime import datetime, timedelta

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

CASES = {
    # name: (scenario class, n_agents, n_districts, days, infect_num, members, R0)
    'unmitigated_100k': ('UnmitigatedScenario', 100_000, 60, 45, 200, 64, 1.9),
    'lockdown_100k': ('ContinuedLockdownScenario', 100_000, 60, 45, 200, 32, 1.9),
    # BASELINE.json configs[2]: selective lockdown of the ten high-mobility districts
    'selective_lockdown_100k': ('LockdownGreatestMobilityScenario', 100_000, 60, 45, 200, 32, 1.9),
    # BASELINE.json configs[3]: continued lockdown with schools and manufacturing re-opened
    'open_manufacturing_schools_100k': ('OpenManufacturingAndSchoolsScenario', 100_000, 60, 45, 200, 32, 1.9),
    # BASELINE.json configs[3]: continued lockdown with the mines open (camps = households of tens to hundreds of miners)
    'open_mining_100k': ('OpenMiningScenario', 100_000, 60, 45, 200, 32, 1.9),
    # school-phase scenario -- schools are extra infection pools, one more phase opens on 1 October
    'accelerated_schools_100k': ('AcceleratedGovernmentOpenSchoolsScenario', 100_000, 60, 45, 200, 32, 1.9),
    'isolate_symptomatic_100k': ('IsolateSymptomaticScenario', 100_000, 60, 45, 200, 32, 1.9),
    'handwashing_risk_100k': ('HandWashingRiskScenario', 100_000, 60, 45, 200, 32, 1.9),
    # round 2: the same two headline scenarios THROUGH THE PEAK and the decline (the 45-day windows above stop before it)
    'unmitigated_100k_120d': ('UnmitigatedScenario', 100_000, 60, 120, 200, 32, 1.9),
    'open_manufacturing_schools_100k_120d': ('OpenManufacturingAndSchoolsScenario', 100_000, 60, 120, 200, 32, 1.9),
}
MINING_SHARE = 0.15
POP_SEED = 2020


def scenario_columns(pop, n_districts):
    """Census columns the sector / school scenarios read (scenario_models.py:329, 349): manufacturing workers are the
    non-null rows, school goers carry the district of their school.  Same function on both sides of the comparison."""
    from abm_b200 import synth
    names = np.array(synth.district_names(n_districts))
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    pop.extras['school_id_district'] = np.where(pop.extras['school_goers'] == 1, names[pop.district], '')
    return pop


def member(args):
    case, m = args
    import ref_harness as rh
    from abm_b200 import synth
    scenario, n_agents, D, days, infect_num, _, R0 = CASES[case]
    start = datetime(2020, 9, 1, 0)
    spec = synth.make_spec(n_districts=D, R0=R0, seed=POP_SEED, start_datetime=start)
    pop = synth.make_population(n_agents, spec, seed=POP_SEED, scenario_flags=True)
    if case in ('selective_lockdown_100k', 'open_manufacturing_schools_100k', 'open_manufacturing_schools_100k_120d'):
        pop = scenario_columns(pop, D)
    if case == 'open_mining_100k':
        pop.extras['mining_district_id'] = synth.mining_camp_column(pop, D, POP_SEED, MINING_SHARE)
    if case == 'accelerated_schools_100k':
        pop.extras.update(synth.school_census_columns(pop, D))
    cases = synth.seed_case_counts(pop, spec, seed=POP_SEED)
    world = rh.RefWorld(spec, pop, seed_cases=cases)
    log_file = os.path.join(world.tmp, 'logs', 'model.log')
    try:
        np.random.seed(1000 + m)
        t0 = time.time()
        model = world.make_model(scenario, seed_infected=infect_num, log_file=log_file)
        t_load = time.time() - t0
        end = start + timedelta(days=days, hours=8)
        t0 = time.time()
        n_steps = 0
        while model.scheduler.real_time <= end:
            model.step()
            n_steps += 1
        t_run = time.time() - t0
        rows = []
        with open(log_file) as fl:
            for line in fl:
                rows.append(json.loads(line.split('$$', 1)[1]))
        return dict(member=m, rows=rows, t_load=t_load, t_run=t_run, n_steps=n_steps)
    finally:
        world.close()


def main(names):
    for case in names:
        scenario, n_agents, D, days, infect_num, members, R0 = CASES[case]
        with mp.Pool(min(int(os.environ.get('REF_PROCS', '8')), os.cpu_count())) as pool:
            res = pool.map(member, [(case, m) for m in range(members)], chunksize=1)
        keys = [k for k in res[0]['rows'][0] if k != 'date']
        data = {k: [[r[k] for r in x['rows']] for x in res] for k in keys}
        run_s = [x['t_run'] for x in res]
        out = dict(case=case, scenario=scenario, n_agents=n_agents, n_districts=D, days=days,
                   infect_num=infect_num, members=members, R0=R0, pop_seed=POP_SEED,
                   dates=[r['date'] for r in res[0]['rows']], counters=data,
                   reference_seconds_per_member=float(np.mean(run_s)),
                   reference_agent_days_per_s_1core=float(n_agents * days / np.mean(run_s)),
                   note=f"timings taken with {os.environ.get('REF_PROCS', '8')} members running concurrently on 8 vCPUs")
        with open(os.path.join(HERE, f'ref_ensemble_{case}.json'), 'w') as fl:
            json.dump(out, fl)
        act = np.array(data['current_contagious_cases'])
        print(case, 'members', members, 'mean active by day', act.mean(axis=0)[::5].round(1),
              's/member', np.mean(run_s))


if __name__ == '__main__':
    main(sys.argv[1:] or list(CASES))
