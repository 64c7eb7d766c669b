"""Second batch of reference pins on the GPU (this file runs after test_gpu_parity.py).

1. CUDA path vs the real reference's functions on four more scenarios (fixtures added after
test_gpu_parity.py::test_cuda_matches_reference_fixture's first three): the hand-washing risk multiplier per district
(`HandWashingRiskScenario`), the elderly kept in their households (`IsolateVulnerableScenario`), travel blocked out
of the most mobile districts (`BlockGreatestMobilityScenario`) and pupils of three open phases attending their schools
-- outside locations numbered after the households (`AcceleratedGovernmentOpenSchoolsScenario` after `open_schools`).  Same bar: every state, location and tick vector and
every log counter bit-exact at every stored snapshot.
2. Tier-2 ensembles (tolerance and method: test_gpu_ensemble.py) for four more scenarios of the real reference: the mines
open under lockdown (camp households of tens to hundreds), a school-phase scenario (schools as extra infection pools, one
more phase opening on 1 October), symptomatic isolation, district hand-washing risk.  Measured with the bit-identical
CPU oracle (tests/ensemble_compare_cpu.py; profiles/r01_ensemble_compare_cpu.txt): final cumulative infections +1.8 % /
-0.3 % / +3.4 % / +1.7 % of the reference mean, means inside the band on >= 95.6 % of the days, <= 2.5 standard errors.
3. The OpenMiningScenario drop-in flow, person by person against the oracle."""
import os

import pytest

from abm_b200 import tables
from util import assert_state_equal, load_fixture

pytestmark = pytest.mark.gpu

CASES = ['handwashing_risk', 'isolate_vulnerable', 'block_mobility', 'open_school_phases']


def replay(name, simulation_class):
    fx = load_fixture(name)
    sim = simulation_class(fx['spec'], fx['pop'], seed=fx['philox_seed'], tables=tables.build_tables(fx['spec']))
    sim.seed_infections(fx['seeds'])
    done = 0
    for k, snap in fx['snaps']:
        sim.step(k - done)
        done = k
        assert_state_equal(sim.state(), snap, f'{name} after sub-step {k}')
    log = sim.log_dicts()
    assert len(log) == len(fx['ref_log'])
    for got, want in zip(log, fx['ref_log']):
        for key, val in want.items():
            if key != 'date':
                assert got[key] == val, (name, key, want['date'], got[key], val)
        assert got['date'] == want['date']
    sim.close()
    return fx


@pytest.mark.parametrize('name', CASES)
def test_cuda_matches_reference_functions(built_library, name):
    from abm_b200.engine import Simulation
    fx = replay(name, Simulation)
    assert fx['ref_log'][-1]['infected_count'] > 500


MORE_ENSEMBLES = [('open_mining_100k', 32), ('accelerated_schools_100k', 32), ('isolate_symptomatic_100k', 32),
                  ('handwashing_risk_100k', 32)]


@pytest.mark.parametrize('case,n_members', MORE_ENSEMBLES)
def test_more_ensembles_inside_reference_band(built_library, tmp_path, case, n_members):
    from test_gpu_ensemble import test_ensemble_mean_inside_reference_band as run_case
    run_case(built_library, tmp_path, case, n_members)


def test_open_mining_dropin_flow(built_library, tmp_path):
    import numpy as np
    import dropin_flow
    old = os.environ.get('COVID19_ABM_DATA_DIR')
    try:
        m, ora, rows = dropin_flow.run_flow(tmp_path, 'OpenMiningScenario', days=10, n_agents=20_000, mining_columns=True,
                                            mining_share=0.6)
        names = m.params.HOUSEHOLD_ID_TO_NAME
        camp_ids = [i for i, n in names.items() if n.startswith('mining_')]
        in_camp = np.isin(m.household_ids, camp_ids)
        assert len(camp_ids) == 12 and np.bincount(m.household_ids[in_camp] - min(camp_ids)).max() > 32
        assert rows[-1]['infected_count'] > 200 and not m.district_mover[in_camp].any()
        m.close()
    finally:
        if old is None:
            os.environ.pop('COVID19_ABM_DATA_DIR', None)
        else:
            os.environ['COVID19_ABM_DATA_DIR'] = old
