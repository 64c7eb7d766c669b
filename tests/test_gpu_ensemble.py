"""Tier-2 parity: ensemble epidemic curves of the CUDA path against the REAL reference run with its own RNG and its
explicit contact sampling (tests/golden/ref_ensemble_*.json, generator tests/golden/make_ref_ensemble.py).

Every member goes through the drop-in package the way the reference members were driven (same synthetic world, same
numpy seed per member -> same seed agents and same lockdown thinning; the disease course and infection draws come
from the device's Philox stream).  Stated tolerance: for each of daily active (current_contagious_cases), critical,
dead and cumulative infected counts, the ensemble MEAN of the CUDA members lies inside the reference members' central
95 % band (2.5th..97.5th percentile across members) on at least 95 % of the logged days, the two ensemble means never
differ by more than 3.5 combined standard errors, and the final cumulative infected mean is within 5 % of the
reference's.  (Measured with the bit-identical CPU oracle, tests/ensemble_compare_cpu.py: inside on 100 % of days,
largest difference 2.2 standard errors, final infected +0.9 % unmitigated / -0.6 % lockdown / +0.3 % selective lockdown;
with schools and manufacturing re-opened under lockdown the pressure-table form runs 4.6 % low after 45 days, 3.2
standard errors at the worst day, still inside the band on every day for active and infected counts.)  This bounds the one restated piece of the path -- the
pressure-table form of the contact sampler -- and the table-based dwell-time sampling."""
import os

import pytest

import ensemble_flow as ef

pytestmark = pytest.mark.gpu

# unmitigated / continued lockdown, and BASELINE.json configs[2] (selective lockdown of the high-mobility districts) and
# configs[3] (continued lockdown with schools and manufacturing re-opened)
CASES = [('unmitigated_100k', 64), ('lockdown_100k', 32), ('selective_lockdown_100k', 32), ('open_manufacturing_schools_100k', 32)]


@pytest.mark.parametrize('case,n_members', CASES)
def test_ensemble_mean_inside_reference_band(built_library, tmp_path, case, n_members):
    ref = ef.load_reference(case)
    assert ref['members'] >= n_members
    root = str(tmp_path / 'data')
    os.makedirs(root)
    old = os.environ.get('COVID19_ABM_DATA_DIR')
    try:
        ef.build_world(root, ref)
        members = []
        for m in range(n_members):
            curves, dates = ef.run_member(ref, m, os.path.join(root, 'logs', 'member.log'))
            assert dates == ref['dates'][:len(dates)] and len(dates) == len(ref['dates'])
            members.append(curves)
    finally:
        if old is None:
            os.environ.pop('COVID19_ABM_DATA_DIR', None)
        else:
            os.environ['COVID19_ABM_DATA_DIR'] = old
    res = ef.compare(ref, members)
    print(case, res)
    for key, r in res.items():
        assert r['inside'] >= 0.95, (case, key, r)
        assert r['max_z'] <= 3.5, (case, key, r)
    fin = res['infected_count']
    assert abs(fin['final_got'] - fin['final_ref']) <= 0.05 * fin['final_ref'], (case, fin)


# ---- round 2: the same comparison THROUGH THE PEAK and the decline (120 days; the 45-day windows above stop before the
# peak).  Reference: 32 members of the unmodified reference per scenario (make_ref_ensemble.py, ~2.5 min per member);
# CUDA: 64 members.  Tolerance as above, plus: the day of the peak of the mean active curve within 3 days of the
# reference's, its height and the final death toll within 5 %.  Measured (profiles/r02_ensemble_120d.txt): unmitigated
# inside the band on 98-100 % of the days, largest difference 2.8 standard errors, peak day 53 vs 54, peak height
# -0.2 %, deaths 204 vs 209, final infected +0.01 %; schools + manufacturing re-opened under lockdown inside on 99-100 %,
# 3.2 standard errors, peak day 64 vs 63, height +0.2 %, final infected -0.03 % -- the 4.6 % lag seen at day 45 is the
# growth phase running about a day late, it closes by the peak.
LONG_CASES = [('unmitigated_100k_120d', 64), ('open_manufacturing_schools_100k_120d', 64)]


@pytest.mark.parametrize('case,n_members', LONG_CASES)
def test_ensemble_through_the_peak(built_library, tmp_path, case, n_members):
    import numpy as np
    ref = ef.load_reference(case)
    root = str(tmp_path / 'data')
    os.makedirs(root)
    old = os.environ.get('COVID19_ABM_DATA_DIR')
    try:
        ef.build_world(root, ref)
        members = []
        for m in range(n_members):
            curves, dates = ef.run_member(ref, m, os.path.join(root, 'logs', 'member.log'))
            assert dates == ref['dates'][:len(dates)] and len(dates) == len(ref['dates'])
            members.append(curves)
    finally:
        if old is None:
            os.environ.pop('COVID19_ABM_DATA_DIR', None)
        else:
            os.environ['COVID19_ABM_DATA_DIR'] = old
    res = ef.compare(ref, members)
    print(case, res)
    for key, r in res.items():
        assert r['inside'] >= 0.95, (case, key, r)
        assert r['max_z'] <= 3.5, (case, key, r)
    fin = res['infected_count']
    assert abs(fin['final_got'] - fin['final_ref']) <= 0.05 * fin['final_ref'], (case, fin)
    ra = np.array(ref['counters']['current_contagious_cases'], dtype=float).mean(axis=0)
    ga = np.array([m['current_contagious_cases'] for m in members], dtype=float).mean(axis=0)
    assert ra.argmax() > 45 and abs(int(ra.argmax()) - int(ga.argmax())) <= 3, (int(ra.argmax()), int(ga.argmax()))
    assert abs(ga.max() - ra.max()) <= 0.05 * ra.max(), (ra.max(), ga.max())
    dead = res['died_count']
    assert dead['final_ref'] > 100 and abs(dead['final_got'] - dead['final_ref']) <= 0.05 * dead['final_ref'], dead


def test_concurrent_replicas_through_the_peak(built_library):
    """The ensemble runner (abm_b200.ensemble.run_ensemble: members as concurrent handles / streams on one GPU)
    on the unmitigated 120-day case: 64 members inside the reference band, and a members-per-second figure."""
    import time
    from datetime import datetime
    import numpy as np
    from abm_b200 import ensemble, synth
    ref = ef.load_reference('unmitigated_100k_120d')
    spec = synth.make_spec(n_districts=ref['n_districts'], R0=ref['R0'], seed=ef.POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(ref['n_agents'], spec, seed=ef.POP_SEED, scenario_flags=True)
    n = 64
    seeds = [synth.pick_seeds(pop, spec, infect_num=ref['infect_num'], seed=1000 + m) for m in range(n)]
    t0 = time.time()
    curves, dates = ensemble.run_ensemble(spec, pop, seeds, [5000 + m for m in range(n)], days=ref['days'], max_concurrent=32)
    secs = time.time() - t0
    from abm_b200.engine import COUNTER_NAMES
    members = [{k: curves[m, :, COUNTER_NAMES.index(k)].tolist() for k in ef.CURVES} for m in range(n)]
    res = ef.compare(ref, members)
    print(f'run_ensemble: {n} members x {ref["days"]} days x {ref["n_agents"]} agents in {secs:.1f} s = {n / secs:.1f} members/s, '
          f'{n * ref["days"] * ref["n_agents"] / secs:.3g} agent-days/s', res)
    # (seed agents come from synth.pick_seeds here, not from load_agents' own draw, so the first days differ by a few
    # cases where the member-to-member spread is tiny: the band and the end points are the test, not the z score)
    for key in ('current_contagious_cases', 'infected_count', 'died_count'):
        assert res[key]['inside'] >= 0.95, (key, res[key])
        assert abs(res[key]['final_got'] - res[key]['final_ref']) <= 0.05 * max(res[key]['final_ref'], 20.0), (key, res[key])
