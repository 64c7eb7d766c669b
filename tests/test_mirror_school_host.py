"""Host side of the three school-phase scenarios of the drop-in against the REAL reference
(tests/golden/ref_mirror_school.npz, generator tests/golden/make_ref_school_fixtures.py): what `load_agents` +
`set_school_params` build, `get_safe_and_unsafe_districts` on a prescribed epidemic state, `open_schools` and
`lockdown_schools` under the same numpy seed.  No GPU: the device store does not exist before seeding / stepping."""
import os
import sys
from datetime import datetime

import numpy as np
import pytest

from abm_b200 import synth
from util import GOLDEN

sys.path.insert(0, GOLDEN)
import make_ref_school_fixtures as gen          # noqa: E402  (constants + the shared call sequence; no reference import)


@pytest.fixture(scope='module')
def world(tmp_path_factory):
    root = str(tmp_path_factory.mktemp('abm_school_data'))
    spec = synth.make_spec(n_districts=gen.N_DISTRICTS, R0=1.9, seed=gen.POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(gen.N_AGENTS, spec, seed=gen.POP_SEED, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    cases = synth.seed_case_counts(pop, spec, seed=gen.POP_SEED)
    synth.write_data_dir(root, spec, pop, seed_cases=cases, seed=gen.POP_SEED, school_columns=True)
    old = os.environ.get('COVID19_ABM_DATA_DIR')
    os.environ['COVID19_ABM_DATA_DIR'] = root
    yield np.load(os.path.join(GOLDEN, 'ref_mirror_school.npz'))
    if old is None:
        os.environ.pop('COVID19_ABM_DATA_DIR', None)
    else:
        os.environ['COVID19_ABM_DATA_DIR'] = old


@pytest.mark.parametrize('name', gen.SCENARIOS)
def test_school_scenario_host_logic_matches_reference(world, name):
    from covid19_abm import scenario_models
    from test_mirror_host import make_params
    ref = world
    np.random.seed(gen.NP_SEED)
    p = make_params()
    m = getattr(scenario_models, name)(p, model_log_file=None, individual_log_file=None)
    p.SIMULATION_START_DATE = datetime(2020, 9, 1, 0)
    m.scheduler.real_time = p.SIMULATION_START_DATE
    m.load_agents(p.data_file_name, size=None, infect_num=None)
    assert m.is_school_scenario and m._sim is None
    got = {}
    gen.exercise(m, got, name + '/')
    assert m._sim is None and m._dirty                 # the hooks marked the (future) device store stale
    keys = [k for k in ref.files if k.startswith(name + '/')]
    assert len(keys) == len(got) == 18
    for k in keys:
        want, have = ref[k], got[k]
        if want.dtype.kind == 'f':
            assert np.array_equal(want, np.asarray(have, dtype=np.float64)), k
        elif want.dtype.kind in 'US':
            assert list(want) == list(have), k
        else:
            assert np.array_equal(want, np.asarray(have)), k
    first = int(ref[name + '/first_school_id'])
    assert (got[name + '/open/econ_loc'] >= first).sum() > (got[name + '/locked/econ_loc'] >= first).sum() > 500
