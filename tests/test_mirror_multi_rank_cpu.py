"""The drop-in `Country` on several ranks (abm_b200/multi.py), host logic on CPU: two gloo processes drive the
reference's own sequence (ParamsConfig, scenario class, load_agents + seeding, scenario hooks, step loop, log file) with
the oracle-backed engine double on their shards; the log file rank 0 writes and the gathered per-agent vectors must be
those of the single-process run, bit for bit.  The same flow runs on two B200s in tests/test_multi_gpu.py."""
import json
import os
import socket
from datetime import datetime

import numpy as np
import pytest
import torch.multiprocessing as mp

import dropin_flow

SCENARIOS = ['UnmitigatedScenario', 'ContinuedLockdownScenario']
DAYS = 6


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(root, scenario, log_file, out_file):
    from covid19_abm import base_model
    from fake_sim import FakeSimulation
    base_model.Country._sim_class = FakeSimulation
    os.environ['COVID19_ABM_DATA_DIR'] = root
    m = dropin_flow.make_model(scenario, datetime(2020, 9, 1, 0), log_file=log_file, np_seed=5, philox_seed=None)
    for _ in range(6 * DAYS):
        m.step()
        m.active_infections()          # the stop test of run_scenario: a collective on several ranks
    m.log_model_output()
    vec = {k: np.asarray(getattr(m, k)).astype(np.int64) for k in
           ('epidemic_state', 'clinical_state', 'current_location_ids', 'current_district_ids', 'infected_at_location_ids')}
    vec['date_recovered'] = m.event_ticks('date_recovered')
    vec['philox_seed'] = np.array([m._philox_seed])
    if out_file:
        np.savez(out_file, **vec)
    m.close()
    return vec


def _rank_main(rank, world, port, root, scenario, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    _run(root, scenario, os.path.join(out_dir, f'log_rank{rank}.log'), os.path.join(out_dir, f'rank{rank}.npz'))
    dist.destroy_process_group()


@pytest.mark.parametrize('scenario', SCENARIOS)
def test_two_rank_dropin_equals_single_process(tmp_path, scenario):
    root = str(tmp_path / 'data')
    os.makedirs(root)
    dropin_flow.build_world(root, n_agents=12_000, n_districts=60)
    single_log = str(tmp_path / 'single.log')
    want = _run(root, scenario, single_log, None)
    out_dir = str(tmp_path)
    mp.spawn(_rank_main, args=(2, _free_port(), root, scenario, out_dir), nprocs=2, join=True)
    got_log = dropin_flow.read_log(os.path.join(out_dir, 'log_rank0.log'))
    want_log = dropin_flow.read_log(single_log)
    assert len(want_log) == DAYS and got_log == want_log                 # rank 0 wrote the reference's log, summed over ranks
    assert not os.path.exists(os.path.join(out_dir, 'log_rank1.log'))     # ... and only rank 0
    for rank in range(2):
        z = np.load(os.path.join(out_dir, f'rank{rank}.npz'))
        for k, w in want.items():
            assert np.array_equal(z[k], w), (rank, k)                     # every rank reads the whole population
    assert want_log[-1]['infected_count'] > 150
