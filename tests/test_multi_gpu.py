"""Two ranks on two GPUs reproduce the single-rank oracle trajectory (tests/multi_gpu_check.py).
Skipped on boxes with fewer than two GPUs; the host-side sharding logic is covered on CPU by
tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('exchange,shards', [('p2p', 'contiguous'), ('nccl', 'contiguous'), ('p2p', 'dealt')])
def test_two_ranks_match_single_rank_oracle(built_library, exchange, shards):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    port = {('p2p', 'contiguous'): '29533', ('nccl', 'contiguous'): '29534', ('p2p', 'dealt'): '29535'}[(exchange, shards)]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', port, os.path.join(ROOT, 'tests', 'multi_gpu_check.py'),
           '--agents', '80000', '--days', '12', '--exchange', exchange, '--shards', shards]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'multi_gpu_check ok' in res.stdout


def test_lost_peer_is_reported_not_summed(built_library):
    """a rank that never publishes its table rows must surface as ABM_ERR_STATE at the next synchronising
    call of its peers (and stay sticky), not as silently partial district x status sums."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29536', os.path.join(ROOT, 'tests', 'multi_gpu_check.py'),
           '--agents', '40000', '--days', '1', '--exchange', 'p2p', '--lost-peer']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'multi_gpu_check ok (lost peer reported)' in res.stdout


def test_run_scenario_on_two_gpus_writes_the_single_gpu_log(built_library, tmp_path):
    """`torchrun ... scripts/run_scenario.py UnmitigatedScenario ...` (the reference's driver function,
    scenario_models.py:443-500, through the drop-in package on 2 ranks) gives the single-GPU log bit for bit."""
    import glob
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import dropin_flow
    logs = {}
    for n in (1, 2):
        root = str(tmp_path / f'data{n}')
        os.makedirs(root)
        dropin_flow.build_world(root, n_agents=40_000)
        env = dict(os.environ, COVID19_ABM_DATA_DIR=root)
        args = ['UnmitigatedScenario', 'Unmitigated', '3', '2.6', '100', '300', '--start-date', '2021-05-20T00:00:00',
                '--np-seed', '5', '--philox-seed', '4711']
        script = os.path.join(ROOT, 'scripts', 'run_scenario.py')
        cmd = [sys.executable, script] + args if n == 1 else \
            [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
             '--master-port', '29537', script] + args
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        found = glob.glob(os.path.join(root, 'logs', 'model_log_file_Unmitigated_03_R2.6_samp100_seed300.*.log'))
        assert len(found) == 1, found
        logs[n] = dropin_flow.read_log(found[0])
        assert len(glob.glob(os.path.join(root, 'logs', 'model_dump_file_*.pickle'))) == 1
    assert len(logs[1]) == 12 and logs[1][-1]['infected_count'] > 600
    assert logs[2] == logs[1]
