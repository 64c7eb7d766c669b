"""The scenario classes test_mirror_host.py does not cover -- improved hand-washing 1 / 2, interaction-matrix
sensitivity, open mining (camps that are household and work place at once), eased lockdown with schools, schools open
with the epidemic seeded among children, manufacturing + schools -- against the REAL reference
(tests/golden/ref_mirror_host_more.npz, generator tests/golden/make_ref_mirror_fixtures_more.py).  With
test_mirror_host.py and test_mirror_school_host.py every scenario class of scenario_models.py is pinned."""
import os
import sys
from datetime import datetime

import numpy as np
import pytest

from abm_b200 import synth
from util import GOLDEN

sys.path.insert(0, GOLDEN)
import make_ref_mirror_fixtures_more as gen          # noqa: E402  (constants only; no reference import at load)
from test_mirror_host import check_host_vectors, load_scenario   # noqa: E402

REFERENCE_CLASSES = [
    'UnmitigatedScenario', 'HandWashingRiskScenario', 'HandWashingRiskImproved1Scenario', 'HandWashingRiskImproved2Scenario',
    'InteractionSensitivityScenario', 'IsolateSymptomaticScenario', 'IsolateVulnerableScenario', 'BlockGreatestMobilityScenario',
    'LockdownGreatestMobilityScenario', 'ContinuedLockdownScenario', 'EasedLockdownScenario', 'OpenMiningScenario',
    'OpenSchoolsScenario', 'EasedOpenSchoolsScenario', 'OpenSchoolsSeedKidsScenario', 'OpenManufacturingScenario',
    'OpenManufacturingAndSchoolsScenario', 'Phase1GovernmentOpenSchoolsScenario',
    'DynamicPhase1GovernmentOpenSchoolsScenario', 'AcceleratedGovernmentOpenSchoolsScenario']      # scenario_models.py:9-440


@pytest.fixture(scope='module')
def ref(tmp_path_factory):
    root = str(tmp_path_factory.mktemp('abm_more_data'))
    spec = synth.make_spec(n_districts=gen.N_DISTRICTS, R0=1.9, seed=gen.POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(gen.N_AGENTS, spec, seed=gen.POP_SEED, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    cases = synth.seed_case_counts(pop, spec, seed=gen.POP_SEED)
    synth.write_data_dir(root, spec, pop, seed_cases=cases, seed=gen.POP_SEED, school_columns=True, mining_columns=True)
    old = os.environ.get('COVID19_ABM_DATA_DIR')
    os.environ['COVID19_ABM_DATA_DIR'] = root
    yield np.load(os.path.join(GOLDEN, 'ref_mirror_host_more.npz'))
    if old is None:
        os.environ.pop('COVID19_ABM_DATA_DIR', None)
    else:
        os.environ['COVID19_ABM_DATA_DIR'] = old


def test_every_reference_scenario_class_exists():
    from covid19_abm import scenario_models
    from covid19_abm.base_model import Country
    for name in REFERENCE_CLASSES:
        cls = getattr(scenario_models, name)
        assert issubclass(cls, Country) and cls.__name__ == name
    assert callable(scenario_models.run_scenario)


@pytest.mark.parametrize('name', gen.SCENARIOS)
def test_remaining_scenarios_host_vectors_match_reference(ref, name):
    m, p = load_scenario(name, gen.NP_SEED)
    check_host_vectors(ref, name, m, p)
    got = {}
    gen.snapshot(m, got, name + '/')
    for key in ('interaction_sizes', 'seed_age_window', 'n_households', 'first_household_names'):
        assert np.atleast_1d(got[name + '/' + key]).tolist() == np.atleast_1d(ref[name + '/' + key]).tolist(), key
    if name == 'OpenMiningScenario':
        # the camps: households named 'mining_<district>', numbered after every 'h_...' household; nobody travels from them
        camp_ids = [i for i, n in p.HOUSEHOLD_ID_TO_NAME.items() if n.startswith('mining_')]
        camp = np.isin(m.household_ids, camp_ids)
        assert len(camp_ids) >= 5 and camp.sum() > 8 and not m.district_mover[camp].any()
        assert np.array_equal(m.economic_activity_location_ids[camp], m.household_ids[camp])
    if name == 'OpenSchoolsSeedKidsScenario':
        assert (p.SEED_INFECT_AGE_MIN, p.SEED_INFECT_AGE_MAX) == (8, 18)
