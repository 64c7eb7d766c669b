"""CUDA path vs the real reference's functions on three more scenarios (fixtures added after
test_gpu_parity.py::test_cuda_matches_reference_fixture's first three): the hand-washing risk multiplier per district
(`HandWashingRiskScenario`), the elderly kept in their households (`IsolateVulnerableScenario`) and travel blocked out
of the most mobile districts (`BlockGreatestMobilityScenario`).  Same bar: every state, location and tick vector and
every log counter bit-exact at every stored snapshot."""
import pytest

from abm_b200 import tables
from util import assert_state_equal, load_fixture

pytestmark = pytest.mark.gpu

CASES = ['handwashing_risk', 'isolate_vulnerable', 'block_mobility']


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
