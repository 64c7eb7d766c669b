"""Generate tests/golden/ref_functions_<case>.npz -- outputs of the REAL reference functions.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_function_fixtures.py

For each case a real reference model (scenario class, real ParamsConfig, real Country vectors) is
advanced sub-step by sub-step by calling the reference's OWN methods in the order of
EpidemicScheduler.step (base_model.py:33-58, 177-179):
    model.log_model_output(); scheduler.update_agent_states(); scheduler.go_home() / go_out();
    the contagious promotion of base_model.py:54-58; model.set_epidemic_status(ids)
while np.random.random / gamma / exponential are replaced by a provider that serves, per agent, the
Philox-derived draw the CUDA kernels use (oracle/philox.py).  The provider identifies each call site by
the caller's line number inside base_model.py and reads the id vector the draw belongs to from the
caller's locals, so no reference logic is duplicated here.
The only step NOT executed by reference code is the explicit contact sampling (base_model.py:60-175),
which has no per-agent draw form: the ids to infect each sub-step are chosen by the
oracle's pressure-table restatement, evaluated in lockstep on the same state, and handed to the
reference's set_epidemic_status.
The reference's state after every sub-step is compared with the oracle's here (must be identical) and
written to the fixture at selected steps; tests/test_oracle_golden.py replays the oracle alone against
the stored reference states on any machine.
"""
import json
import os
import sys
from datetime import datetime, timedelta

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import ref_harness as rh                      # noqa: E402
from abm_b200 import synth, tables            # noqa: E402
from oracle import philox                     # noqa: E402
from oracle.abm_oracle import OracleModel, u31, sample_quantile   # noqa: E402

# call-site line ranges in /root/reference/src/covid19_abm/base_model.py
SITES = [
    ((309, 319), 'mover', 'active_ids'),
    ((372, 378), 'od', 'district_out_mover_ids'),
    ((426, 433), 'stay', 'actual_district_out_mover_ids'),
    ((853, 858), 'sym', 'neighbors_to_infect'),
    ((868, 875), 'incubation', 'symptomatic_neighbor_ids'),
    ((898, 902), 'hosp', 'symptomatic_neighbor_ids'),
    ((909, 913), 'crit', 'hospitalized_ids'),
    ((919, 922), 'crit_fatal', 'critical_ids'),
    ((930, 934), 'hosp_fatal', 'hospitalized_ids'),
    ((1011, 1018), 'sym_recovery', 'symptomatic_not_hospitalized_not_dead'),
    ((1036, 1042), 'asym_exposed', 'asymptomatic_neighbor_ids'),
    ((1044, 1051), 'asym_recovery', 'asymptomatic_neighbor_ids'),
]


class DrawProvider:
    """Serves np.random.* calls made by reference code with per-agent Philox draws."""

    def __init__(self, oracle: OracleModel, model):
        self.o, self.model = oracle, model
        self.calls = []

    def _site(self):
        f = sys._getframe(2)
        while f is not None and not f.f_code.co_filename.endswith('covid19_abm/base_model.py'):
            f = f.f_back
        assert f is not None, 'np.random called from outside base_model.py'
        for (lo, hi), name, var in SITES:
            if lo <= f.f_lineno <= hi:
                ids = np.asarray(f.f_locals[var], dtype=np.int64)
                self.calls.append(name)
                return name, ids, f
        raise AssertionError(f'unexpected np.random call at base_model.py:{f.f_lineno}')

    def _lanes(self, ids, site):
        return philox.agent_draws(self.o.seed, self.o.k, self.o.agent_ids[ids], site)

    def _group(self, ids, site):
        return philox.group_draws(self.o.seed, self.o.k, self.o.agent_ids[ids], site)

    def random(self, size=None):
        name, ids, _ = self._site()
        n = int(np.prod(size))
        assert n == ids.shape[0], (name, n, ids.shape)
        if name in ('mover', 'od'):
            u = u31(self._group(ids, philox.SITE_MOVER if name == 'mover' else philox.SITE_OD))
            return u.reshape(size) if size is not None else u
        lane = {'sym': (philox.SITE_COURSE_A, 0),
                'hosp': (philox.SITE_COURSE_A, 2), 'crit': (philox.SITE_COURSE_A, 3),
                'crit_fatal': (philox.SITE_COURSE_B, 0), 'hosp_fatal': (philox.SITE_COURSE_B, 1)}[name]
        u = u31(self._lanes(ids, lane[0])[lane[1]])
        return u.reshape(size) if size is not None else u

    def gamma(self, shape, scale, size=None):
        name, ids, f = self._site()
        tab = self.o.tab
        if name == 'stay':
            # length of stay in hours: Wilson-Hilferty transform of the table normal, floored to minutes
            dow = self.model.scheduler.real_time.weekday()
            src = self.model.district_ids[ids]
            dst = np.asarray(f.f_locals['dst_district_ids'], dtype=np.int64)
            par = tab['stay_par'][dow, src, dst]
            x3 = self._group(ids, philox.SITE_STAY)
            mag = sample_quantile(tab['znorm'], (x3.astype(np.uint64) << np.uint64(1)) & np.uint64(0xFFFFFFFF))
            z = np.where(x3 >> np.uint32(31) == 1, -mag, mag).astype(np.float64) / 65536.0
            t = par[:, 1] + par[:, 2] * z
            hours = par[:, 0] * ((t * t) * t)
            minutes = np.floor(np.clip(hours * 60.0, 0.0, 1.0e9))
            return minutes / 60.0
        which, site, lane = {'incubation': (0, philox.SITE_COURSE_A, 1),
                             'sym_recovery': (2, philox.SITE_COURSE_B, 2),
                             'asym_recovery': (3, philox.SITE_COURSE_B, 2)}[name]
        assert size == ids.shape[0]
        minutes = sample_quantile(tab['dwell'][which], self._lanes(ids, site)[lane])
        return minutes / 1440.0

    def exponential(self, scale, size=None):
        name, ids, _ = self._site()
        assert name == 'asym_exposed' and size == ids.shape[0]
        minutes = sample_quantile(self.o.tab['dwell'][1], self._lanes(ids, philox.SITE_COURSE_A)[1])
        return minutes / 1440.0

    def __enter__(self):
        self._saved = (np.random.random, np.random.gamma, np.random.exponential)
        np.random.random, np.random.gamma, np.random.exponential = self.random, self.gamma, self.exponential
        return self

    def __exit__(self, *exc):
        np.random.random, np.random.gamma, np.random.exponential = self._saved


def compare(ref_state, oracle, where):
    ost = oracle.state_dict()
    for k, v in ref_state.items():
        if not np.array_equal(v, ost[k]):
            bad = np.nonzero(v != ost[k])[0]
            raise AssertionError(f'{where}: {k} differs at {bad[:10]} ref={v[bad[:10]]} oracle={ost[k][bad[:10]]}')


def run_case(name, scenario, n_agents, n_districts, n_steps, seed, philox_seed, snapshot_every, infect_num,
             start=datetime(2020, 9, 1, 0), school_columns=False, open_phases=None):
    spec0 = synth.make_spec(n_districts=n_districts, R0=2.6, seed=seed, start_datetime=start)
    pop = synth.make_population(n_agents, spec0, seed=seed, scenario_flags=True)
    if school_columns:                 # census columns of the school-phase scenarios (school_id, phase, ...)
        pop.extras.update(synth.school_census_columns(pop, n_districts))
    seeds = synth.pick_seeds(pop, spec0, infect_num=infect_num, seed=seed)
    world = rh.RefWorld(spec0, pop, seed_cases={0: 1})
    log_file = os.path.join(world.tmp, 'logs', 'model.log')
    try:
        model = world.make_model(scenario, seed_infected=1, log_file=log_file, infect=False)
        params = model.params
        if open_phases:                # what Country.step's clock hook does on the 1st of a month (base_model.py:812-819)
            model.active_school_phases = list(open_phases)
            model.open_schools(np.unique(model.district_ids), model.active_school_phases)
        n_schools = len(params.SCHOOL_ID_TO_NAME) if getattr(model, 'is_school_scenario', False) else 0
        spec = rh.spec_from_reference_params(params, spec0.step_minutes, start, n_extra_pools=n_schools)
        # population as the reference holds it after its own (scenario) initialisation
        pop.district_mover = np.asarray(model.district_mover).astype(np.int8)
        pop.move_prob_weekday = np.asarray(model.economic_status_weekday_movement_probability, dtype=np.float64).copy()
        pop.move_prob_other = np.asarray(model.economic_status_other_day_movement_probability, dtype=np.float64).copy()
        assert np.array_equal(np.asarray(model.district_ids), pop.district)
        assert np.array_equal(np.asarray(model.household_ids), pop.household + n_districts)
        # where the scenario hook sends people at 08:00 (IsolateVulnerable keeps the elderly in their household; the
        # school-phase scenarios send the pupils of the open phases to their school, an outside location numbered
        # after the households, base_model.py:583-591, 801-805)
        econ = np.asarray(model.economic_activity_location_ids).astype(np.int64)
        at_household = econ == np.asarray(model.household_ids)
        first_school = n_districts + pop.n_households
        at_school = econ >= first_school if n_schools else np.zeros(pop.n, dtype=bool)
        assert (at_household | at_school | (econ == np.asarray(model.district_ids))).all()
        pop.econ_kind = np.where(at_household, rh.C.EKIND_HOUSEHOLD,
                                 np.where(at_school, rh.C.EKIND_POOL, rh.C.EKIND_HOME_DISTRICT)).astype(pop.econ_kind.dtype)
        if n_schools:
            assert min(params.SCHOOL_ID_TO_NAME) == first_school and at_school.sum() > 50
            pop.econ_pool = np.where(at_school, n_districts + (econ - first_school), 0).astype(np.int32)
        tab = tables.build_tables(spec)
        par = dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob)
        oracle = OracleModel(pop, par, tab, philox_seed)
        prov = DrawProvider(oracle, model)

        with prov:                                   # seeding = set_epidemic_status at t0 (base_model.py:796)
            model.set_epidemic_status(seeds)
        oracle.set_epidemic_status(seeds)
        compare(rh.reference_state(model, start), oracle, 'after seeding')

        snaps, snap_steps, site_hits = [], [], {}
        sched = model.scheduler
        for k in range(n_steps):
            now_dt = sched.real_time
            assert (now_dt - start) // timedelta(minutes=1) == oracle.now()
            with prov:
                model.log_model_output()                                         # :36
                sched.update_agent_states()                                      # :38
                if now_dt.hour in [params.WEEKDAY_END_DAY_HOUR, params.OTHER_DAY_END_DAY_HOUR]:
                    sched.go_home()                                              # :42-43
                if now_dt.hour in [params.WEEKDAY_START_DAY_HOUR, params.OTHER_DAY_START_DAY_HOUR]:
                    sched.go_out()                                               # :46-47
                newc = np.where((model.date_start_contagious < now_dt) &
                                (model.epidemic_state < model.STATE_CONTAGIOUS))[0]   # :54-58
                model.epidemic_state[newc] = model.STATE_CONTAGIOUS
            # oracle: same phases, then its mean-field choice of who is infected this sub-step
            now, hour = oracle.now(), oracle.hour()
            if hour == 8:
                row = oracle.counters(now); row['tick'] = now; oracle.log_rows.append(row)
            oracle.update_agent_states(now)
            if hour == 16:
                oracle.go_home(now)
            if hour == 8:
                oracle.go_out(now)
            onew = (oracle.date_start_contagious < now) & (oracle.epidemic_state < 2)
            oracle.epidemic_state[onew] = 2
            compare(rh.reference_state(model, start), oracle, f'step {k} before infection')
            newly = oracle.infect(now)
            with prov:
                model.set_epidemic_status(newly)                                 # :155 / :175
            sched.steps += 1; sched.time += 1; sched.real_time += sched.step_timedelta   # :177-179
            oracle.k += 1
            st = rh.reference_state(model, start)
            compare(st, oracle, f'step {k}')
            if (k + 1) % snapshot_every == 0 or k == n_steps - 1:
                snaps.append(st); snap_steps.append(k + 1)
        for c in prov.calls:
            site_hits[c] = site_hits.get(c, 0) + 1
        # the reference's own log lines
        ref_log = []
        with open(log_file) as fl:
            for line in fl:
                ref_log.append(json.loads(line.split('$$', 1)[1]))
        assert len(ref_log) == len(oracle.log_rows)
        for r, o in zip(ref_log, oracle.log_rows):
            for key, val in r.items():
                if key != 'date':
                    assert val == o[key], (key, r, o)
        print(name, 'ok;', 'final counters', ref_log[-1] if ref_log else None)
        print('  call sites exercised:', site_hits)
        out = dict(
            n_districts=n_districts, n_steps=n_steps, philox_seed=philox_seed, seeds=seeds,
            snap_steps=np.array(snap_steps), start=np.array([start.year, start.month, start.day, start.hour]),
            pop_district=pop.district, pop_household=pop.household, pop_age=pop.age, pop_status=pop.status,
            pop_econ_kind=pop.econ_kind, pop_district_mover=pop.district_mover,
            pop_move_prob_weekday=pop.move_prob_weekday, pop_move_prob_other=pop.move_prob_other,
            spec_R0=spec.R0, spec_step_minutes=spec.step_minutes, spec_od_cum=spec.od_cum,
            spec_stay_mean=spec.stay_mean, spec_stay_std=spec.stay_std,
            spec_interaction_matrix=spec.interaction_matrix, spec_interaction_sizes=spec.interaction_sizes,
            spec_mild=spec.mild_symptom_move_prob,
            spec_hw_risk=np.zeros(0) if spec.hw_risk is None else spec.hw_risk,
            ref_log=json.dumps(ref_log),
        )
        if n_schools:
            out['spec_n_extra_pools'] = n_schools
            out['pop_econ_pool'] = pop.econ_pool
        for i, st in enumerate(snaps):
            for key, val in st.items():
                small = val.astype(np.int32) if val.dtype == np.int64 else val
                out[f'snap{i}_{key}'] = small
        np.savez_compressed(os.path.join(HERE, f'ref_functions_{name}.npz'), **out)
    finally:
        world.close()


if __name__ == '__main__':
    run_case('unmitigated', 'UnmitigatedScenario', n_agents=4000, n_districts=6, n_steps=6 * 45, seed=11,
             philox_seed=0x5EED0001, snapshot_every=54, infect_num=60)
    run_case('isolate_symptomatic', 'IsolateSymptomaticScenario', n_agents=2500, n_districts=5, n_steps=6 * 30,
             seed=12, philox_seed=0x5EED0002, snapshot_every=60, infect_num=50)
    run_case('lockdown', 'ContinuedLockdownScenario', n_agents=2500, n_districts=5, n_steps=6 * 30,
             seed=13, philox_seed=0x5EED0003, snapshot_every=60, infect_num=50)
    run_case('handwashing_risk', 'HandWashingRiskScenario', n_agents=2500, n_districts=5, n_steps=6 * 30,
             seed=14, philox_seed=0x5EED0004, snapshot_every=60, infect_num=50)
    run_case('isolate_vulnerable', 'IsolateVulnerableScenario', n_agents=2500, n_districts=5, n_steps=6 * 30,
             seed=15, philox_seed=0x5EED0005, snapshot_every=60, infect_num=50)
    run_case('block_mobility', 'BlockGreatestMobilityScenario', n_agents=2500, n_districts=12, n_steps=6 * 30,
             seed=16, philox_seed=0x5EED0006, snapshot_every=60, infect_num=50)
    run_case('open_school_phases', 'AcceleratedGovernmentOpenSchoolsScenario', n_agents=4000, n_districts=5, n_steps=6 * 30,
             seed=17, philox_seed=0x5EED0007, snapshot_every=60, infect_num=60, school_columns=True, open_phases=[1, 2, 3])
