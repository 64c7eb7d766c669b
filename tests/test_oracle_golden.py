"""Oracle vs the REAL reference: replay the oracle alone against states the reference's own functions
produced (tests/golden/make_ref_function_fixtures.py).  Integer/tick vectors must match exactly."""
import pytest

from oracle.abm_oracle import OracleModel
from abm_b200 import tables
from util import assert_state_equal, load_fixture, oracle_params

CASES = ['unmitigated', 'isolate_symptomatic', 'lockdown', 'handwashing_risk', 'isolate_vulnerable', 'block_mobility',
         'open_school_phases']


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_functions(name):
    fx = load_fixture(name)
    spec, pop = fx['spec'], fx['pop']
    ora = OracleModel(pop, oracle_params(spec), tables.build_tables(spec), fx['philox_seed'])
    ora.set_epidemic_status(fx['seeds'])
    snaps = dict(fx['snaps'])
    for k in range(fx['n_steps']):
        ora.step()
        if k + 1 in snaps:
            assert_state_equal(ora.state_dict(), snaps[k + 1], f'{name} after sub-step {k + 1}')
    assert len(ora.log_rows) == len(fx['ref_log'])
    for got, want in zip(ora.log_rows, fx['ref_log']):
        for key, val in want.items():
            if key != 'date':
                assert got[key] == val, (name, key, want['date'])
    assert fx['ref_log'][-1]['infected_count'] > 500     # the fixture exercises a real epidemic
