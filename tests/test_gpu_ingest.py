"""Device-side ingestion on the GPU: same arrays as the host build, a working simulation on top of them, and the drop-in
`Country` using it (the trajectory does not depend on where the store was built)."""
import os
import time
from datetime import datetime

import numpy as np
import pytest

from abm_b200 import synth, tables
from test_ingest import COLUMNS, host_reference, person_order_columns

pytestmark = pytest.mark.gpu


def test_device_store_on_gpu_equals_host_store_and_steps(built_library):
    from abm_b200.engine import Simulation
    from abm_b200.ingest import DeviceStore
    spec = synth.make_spec(n_districts=12, R0=2.4, seed=5)
    pop = synth.make_population(300_000, spec, seed=5)
    cols = person_order_columns(pop, seed=9)
    order, ref_pop, want = host_reference(spec, cols)
    got = DeviceStore(spec, device=0, **cols)
    assert np.array_equal(got.order, order)
    for k, v in want.host.items():
        assert np.array_equal(got.host[k].cpu().numpy(), v.numpy()), k
    assert np.array_equal(got.sm.slot_of, want.sm.slot_of) and np.array_equal(got.sm.sub_info, want.sm.sub_info)
    tab = tables.build_tables(spec)
    seeds = synth.pick_seeds(ref_pop, spec, infect_num=600, seed=5)
    a = Simulation(spec, got.pop, seed=7, tables=tab, host_store=got)
    b = Simulation(spec, ref_pop, seed=7, tables=tab, host_store=want)
    for sim in (a, b):
        sim.seed_infections(seeds)
        sim.step(6 * 8)
    assert np.array_equal(a.read_log(), b.read_log()) and a.read_log()[-1, 1] > 1500
    sa, sb = a.state(), b.state()
    for k in sb:
        assert np.array_equal(sa[k], sb[k]), k
    a.close(); b.close()


def test_dropin_trajectory_does_not_depend_on_where_the_store_is_built(built_library, tmp_path):
    import dropin_flow
    from covid19_abm import base_model
    root = str(tmp_path / 'data')
    os.makedirs(root)
    dropin_flow.build_world(root, n_agents=30_000)
    logs = []
    try:
        for on_device in (True, False):
            base_model.Country.ingest_on_device = on_device
            m = dropin_flow.make_model('ContinuedLockdownScenario', datetime(2020, 9, 1, 0), log_file=None, np_seed=5, philox_seed=99)
            m.run_days(10)
            logs.append((m._sim.read_log().copy(), np.asarray(m.epidemic_state).copy(), np.asarray(m.current_location_ids).copy()))
            m.close()
    finally:
        base_model.Country.ingest_on_device = True
    for x, y in zip(logs[0], logs[1]):
        assert np.array_equal(x, y)
    assert logs[0][0][-1, 1] > 150


def test_ingest_time_at_scale(built_library):
    """4 M agents from person-order integer columns to the first step: a number for profiles/, and a sanity bound."""
    from abm_b200.engine import Simulation
    from abm_b200.ingest import DeviceStore
    spec = synth.make_spec(n_districts=60, R0=1.9, seed=2020)
    pop = synth.make_population(4_000_000, spec, seed=2020)
    cols = person_order_columns(pop, seed=1)
    import torch
    torch.cuda.synchronize()
    t0 = time.time()
    store = DeviceStore(spec, device=0, **cols)
    sim = Simulation(spec, store.pop, seed=1, tables=tables.build_tables(spec), host_store=store)
    sim.step(1)
    sim.sync()
    secs = time.time() - t0
    print(f'ingest: {pop.n} agents, person-order integer columns -> first sub-step in {secs:.2f} s')
    assert secs < 10
    sim.close()
