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
