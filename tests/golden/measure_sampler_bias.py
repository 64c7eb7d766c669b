"""Per-sub-step comparison of the REAL reference's explicit contact sampler (base_model.py:60-175) with the
pressure-table restatement the oracle and the CUDA kernels use, on one and the same state.

Run in the build container only (needs /root/reference):
    python tests/golden/measure_sampler_bias.py [scenario] [n_agents] [repeats]

The reference model and the oracle are advanced in lock-step exactly as tests/golden/make_ref_function_fixtures.py does
(reference functions with the kernel's draws injected; infections chosen by the oracle), so both hold the same state.
At a few sub-steps without movement (12:00 and 20:00) the state just before the infection phase is frozen and
  * the reference's own `EpidemicScheduler.step()` is run `repeats` times from it with numpy's native RNG (state
    restored in between; the phases before the infection loop are idempotent at a fixed clock), counting who it infects;
  * the restatement's expectation is the sum of the per-susceptible infection probabilities (thresholds / 2^31).
Printed: mean new infections per sub-step of the reference (with its standard error) against that expectation, split
into household and outside locations and by economic status.  Result file: profiles/r01_sampler_bias_<scenario>.json.
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

import make_ref_ensemble as ens               # noqa: E402
import make_ref_function_fixtures as g        # noqa: E402
import ref_harness as rh                      # noqa: E402
from abm_b200 import synth, tables            # noqa: E402
from oracle.abm_oracle import OracleModel, hazard_threshold_fx   # noqa: E402

MEASURE_AT = [(10, 12), (20, 12), (20, 20), (30, 12), (30, 20), (38, 12)]       # (day, hour)


def snapshot(model):
    keep = {k: v.copy() for k, v in vars(model).items() if isinstance(v, np.ndarray)}
    s = model.scheduler
    return keep, (s.steps, s.time, s.real_time)


def restore(model, snap):
    keep, (steps, time_, real_time) = snap
    for k, v in keep.items():
        setattr(model, k, v.copy())
    s = model.scheduler
    s.steps, s.time, s.real_time = steps, time_, real_time


def expected_by_cell(o):
    """Restatement: expected new infections this sub-step, [2 (household, outside)][A]."""
    rsum, ncnt = o.pressure_tables()
    thr_out = o.outside_thresholds(rsum, ncnt)
    at_home, hsum = o.household_hazard()
    sus = o.epidemic_state < 1
    loc = o.current_location_ids
    pool = o.pool_of(loc)
    out = np.zeros((2, o.A))
    m = sus & (pool >= 0)
    np.add.at(out[1], o.economic_status_ids[m], thr_out[pool[m], o.economic_status_ids[m]] / 2147483648.0)
    h = np.nonzero(sus & at_home)[0]
    hs = hsum[loc[h] - o.D]
    ex = hs > 0
    p = hazard_threshold_fx(o.tab['omexp'], hs[ex] << np.uint64(8)) / 2147483648.0
    np.add.at(out[0], o.economic_status_ids[h[ex]], p)
    return out


def main():
    scenario = sys.argv[1] if len(sys.argv) > 1 else 'OpenManufacturingAndSchoolsScenario'
    n_agents = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 120
    D, start = 60, datetime(2020, 9, 1, 0)
    spec0 = synth.make_spec(n_districts=D, R0=1.9, seed=ens.POP_SEED, start_datetime=start)
    pop = synth.make_population(n_agents, spec0, seed=ens.POP_SEED, scenario_flags=True)
    pop = ens.scenario_columns(pop, D)
    cases = synth.seed_case_counts(pop, spec0, seed=ens.POP_SEED)
    world = rh.RefWorld(spec0, pop, seed_cases=cases)
    try:
        np.random.seed(1000)
        model = world.make_model(scenario, seed_infected=200, log_file=None, infect=False)
        params = model.params
        spec = rh.spec_from_reference_params(params, spec0.step_minutes, start)
        pop.district_mover = np.asarray(model.district_mover).astype(np.int8)
        pop.move_prob_weekday = np.asarray(model.economic_status_weekday_movement_probability, dtype=np.float64).copy()
        pop.move_prob_other = np.asarray(model.economic_status_other_day_movement_probability, dtype=np.float64).copy()
        econ = np.asarray(model.economic_activity_location_ids).astype(np.int64)
        pop.econ_kind = np.where(econ == np.asarray(model.household_ids), rh.C.EKIND_HOUSEHOLD,
                                 rh.C.EKIND_HOME_DISTRICT).astype(pop.econ_kind.dtype)
        tab = tables.build_tables(spec)
        oracle = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob), tab, 0x5A3B1E)
        prov = g.DrawProvider(oracle, model)
        seeds = synth.pick_seeds(pop, spec0, infect_num=200, seed=ens.POP_SEED)
        with prov:
            model.set_epidemic_status(seeds)
        oracle.set_epidemic_status(seeds)
        sched = model.scheduler
        want = {(d * 6 + h // 4) for d, h in MEASURE_AT}
        status_names = [params.ECON_STAT_ID_TO_NAME[i] for i in range(len(params.ECON_STAT_ID_TO_NAME))]
        results = []
        for k in range(max(want) + 1):
            now_dt = sched.real_time
            with prov:
                sched.update_agent_states()
                if now_dt.hour in [params.WEEKDAY_END_DAY_HOUR, params.OTHER_DAY_END_DAY_HOUR]:
                    sched.go_home()
                if now_dt.hour in [params.WEEKDAY_START_DAY_HOUR, params.OTHER_DAY_START_DAY_HOUR]:
                    sched.go_out()
                newc = np.where((model.date_start_contagious < now_dt) & (model.epidemic_state < model.STATE_CONTAGIOUS))[0]
                model.epidemic_state[newc] = model.STATE_CONTAGIOUS
            now, hour = oracle.now(), oracle.hour()
            oracle.update_agent_states(now)
            if hour == 16:
                oracle.go_home(now)
            if hour == 8:
                oracle.go_out(now)
            onew = (oracle.date_start_contagious < now) & (oracle.epidemic_state < 2)
            oracle.epidemic_state[onew] = 2
            if k in want:
                g.compare(rh.reference_state(model, start), oracle, f'step {k} before infection')
                exp = expected_by_cell(oracle)
                snap = snapshot(model)
                at_home = np.asarray(model.current_location_ids) >= D
                status = np.asarray(model.economic_status_ids)
                got = np.zeros((repeats, 2, oracle.A))
                for r in range(repeats):
                    np.random.seed(77_000 + 1000 * k + r)
                    sched.step()                                   # the reference's own step, native RNG
                    new = np.nonzero(np.asarray(model.date_infected) == now_dt)[0]
                    np.add.at(got[r, 0], status[new[at_home[new]]], 1)
                    np.add.at(got[r, 1], status[new[~at_home[new]]], 1)
                    restore(model, snap)
                mean, se = got.mean(axis=0), got.std(axis=0, ddof=1) / np.sqrt(repeats)
                tot, tot_se = got.sum(axis=(1, 2)).mean(), got.sum(axis=(1, 2)).std(ddof=1) / np.sqrt(repeats)
                row = dict(step=k, date=now_dt.isoformat(), contagious=int((oracle.epidemic_state == 2).sum()),
                           susceptible=int((oracle.epidemic_state == 0).sum()),
                           reference_total=float(tot), reference_total_se=float(tot_se), restated_total=float(exp.sum()),
                           reference_household=float(mean[0].sum()), restated_household=float(exp[0].sum()),
                           reference_outside=float(mean[1].sum()), restated_outside=float(exp[1].sum()),
                           outside_by_status={status_names[a]: [float(mean[1, a]), float(se[1, a]), float(exp[1, a])]
                                              for a in range(oracle.A)})
                results.append(row)
                print(f"{row['date']} contagious {row['contagious']}: reference {tot:.1f} +- {tot_se:.1f}  restated {exp.sum():.1f} "
                      f"({100 * (exp.sum() / tot - 1):+.1f} %)  household {mean[0].sum():.1f} / {exp[0].sum():.1f}  "
                      f"outside {mean[1].sum():.1f} / {exp[1].sum():.1f}", flush=True)
            newly = oracle.infect(now)
            with prov:
                model.set_epidemic_status(newly)
            sched.steps += 1; sched.time += 1; sched.real_time += sched.step_timedelta
            oracle.k += 1
        out = dict(scenario=scenario, n_agents=n_agents, repeats=repeats, rows=results,
                   note='reference = mean over `repeats` runs of the unmodified EpidemicScheduler.step() from the same state; '
                        'restated = sum of the pressure-table infection probabilities on that state')
        with open(os.path.join(REPO, 'profiles', f'r01_sampler_bias_{scenario}.json'), 'w') as fl:
            json.dump(out, fl, indent=1)
    finally:
        world.close()


if __name__ == '__main__':
    main()
