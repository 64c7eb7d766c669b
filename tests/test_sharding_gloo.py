"""Multi-rank host logic on CPU (gloo, world_size 2): household-aligned shards + one all-reduce of the
district x status tables per sub-step reproduce the single-rank trajectory exactly.  The data path of the ranks is
the oracle here (no GPU in this suite); the same check runs on two B200s in tests/test_multi_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from abm_b200 import sharding, synth, tables
from oracle.abm_oracle import OracleModel
from util import STATE_KEYS, oracle_params

N_AGENTS, N_DISTRICTS, SEED, DAYS = 24_000, 6, 23, 10


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _world():
    spec = synth.make_spec(n_districts=N_DISTRICTS, R0=2.8, seed=SEED)
    pop = synth.make_population(N_AGENTS, spec, seed=SEED)
    seeds = synth.pick_seeds(pop, spec, infect_num=200, seed=SEED)
    return spec, pop, seeds, tables.build_tables(spec)


def _rank_main(rank, world_size, port, out_dir, dealt=False):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    from abm_b200 import distributed
    spec, pop, seeds, tab = _world()
    if dealt:
        idx = sharding.deal_population(pop, world_size)[rank]
        shard = sharding.take_population(pop, idx)
    else:
        lo, hi = sharding.split_population(pop, world_size)[rank]
        shard = sharding.slice_population(pop, lo, hi)
        idx = np.arange(lo, hi, dtype=np.int64)

    def reduce_fn(buf):
        t = torch.from_numpy(buf.view(np.int64).copy())
        dist.all_reduce(t)
        return t.numpy().view(np.uint64)

    if dealt:
        ora = OracleModel(shard, oracle_params(spec), tab, 4242, agent_ids=idx, reduce_fn=reduce_fn)
    else:
        ora = OracleModel(shard, oracle_params(spec), tab, 4242, agent_id_base=lo, reduce_fn=reduce_fn)
    mine = seeds[np.isin(seeds, idx)]
    ora.set_epidemic_status(np.searchsorted(idx, mine))
    for _ in range(6 * DAYS):
        ora.step()
    rows = np.array([[r[k] for k in sorted(r) if k != 'tick'] for r in ora.log_rows], dtype=np.int64)
    total = distributed.sum_over_ranks(rows, 'cpu')
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), idx=idx, log=total, **ora.state_dict())
    dist.destroy_process_group()


@pytest.mark.parametrize('dealt', [False, True], ids=['contiguous', 'dealt'])
def test_two_gloo_ranks_equal_single_rank(tmp_path, dealt):
    spec, pop, seeds, tab = _world()
    single = OracleModel(pop, oracle_params(spec), tab, 4242)
    single.set_epidemic_status(seeds)
    for _ in range(6 * DAYS):
        single.step()
    want = single.state_dict()
    want_log = np.array([[r[k] for k in sorted(r) if k != 'tick'] for r in single.log_rows], dtype=np.int64)

    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path), dealt), nprocs=2, join=True)
    D = spec.n_districts
    covered = np.zeros(pop.n, dtype=int)
    for rank in range(2):
        z = np.load(tmp_path / f'rank{rank}.npz')
        idx = z['idx']
        starts = idx[np.r_[True, np.diff(idx) != 1]]                                # first agent of every id run
        assert all(pop.household[s] != pop.household[s - 1] for s in starts if s)  # cut on household boundaries
        assert (np.unique(pop.district[idx]).shape[0] == D) if dealt else True     # dealt: a piece of every district
        for key in STATE_KEYS:
            g = z[key].astype(np.int64)
            if key in ('current_location_ids', 'infected_at_location_ids'):
                g = np.where(g >= D, pop.household[idx] + D, g)                     # the rank numbers its households locally
            assert np.array_equal(g, want[key][idx].astype(np.int64)), (rank, key)
        assert np.array_equal(z['log'], want_log), rank
        covered[idx] += 1
    assert np.all(covered == 1)
    assert want_log[-1].max() > 500          # a real epidemic was simulated


def test_synthetic_shards_tile_the_population():
    """make_synthetic_shard builds each rank's agents without materialising the others; together they are exactly
    the population make_population builds in one piece."""
    spec = synth.make_spec(n_districts=12, seed=3)
    full = synth.make_population(50_000, spec, seed=3)
    for world in (1, 2, 3, 4, 8):
        parts = [sharding.make_synthetic_shard(50_000, spec, 3, world, r) for r in range(world)]
        pos = 0
        for shard, base, dstart in parts:
            assert base == pos
            n = shard.n
            assert np.array_equal(shard.district, full.district[pos:pos + n])
            assert np.array_equal(shard.age, full.age[pos:pos + n])
            assert np.array_equal(shard.status, full.status[pos:pos + n])
            assert np.array_equal(np.diff(shard.household) != 0, np.diff(full.household[pos:pos + n]) != 0)
            if pos:
                assert full.household[pos] != full.household[pos - 1]
            pos += n
        assert pos == full.n
        assert np.array_equal(parts[0][2], np.searchsorted(full.district, np.arange(13)))
        # dealt shards: piece r of every district, built one district at a time
        seen = np.zeros(full.n, dtype=int)
        want_idx = sharding.deal_population(full, world)
        for r in range(world):
            shard, ids, dstart = sharding.make_synthetic_shard_dealt(50_000, spec, 3, world, r)
            assert np.array_equal(ids, want_idx[r])
            for name in ('district', 'age', 'status', 'econ_kind', 'district_mover', 'move_prob_weekday'):
                assert np.array_equal(getattr(shard, name), getattr(full, name)[ids]), name
            assert np.array_equal(np.diff(shard.household) != 0, np.diff(full.household[ids]) != 0)
            assert shard.household[0] == 0 and np.all(np.diff(shard.household) >= 0)
            seen[ids] += 1
        assert np.all(seen == 1)
