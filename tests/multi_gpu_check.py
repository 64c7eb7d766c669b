#!/usr/bin/env python
"""Sharding invariance on real GPUs: W ranks, each holding a household-aligned shard, must reproduce the
single-rank oracle trajectory bit for bit (draws are keyed by the global agent id; the only exchange is
the all-reduce of the district x status tables).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [--agents 120000] [--days 20]

Every rank runs the full-population oracle itself (test infrastructure, CPU) and compares its own shard.
Exit code 0 and a line `multi_gpu_check ok` on rank 0 when all ranks agree.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'covid19-agent-based-model_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=120_000)
    ap.add_argument('--districts', type=int, default=8)
    ap.add_argument('--days', type=int, default=20)
    ap.add_argument('--exchange', default='p2p', choices=['p2p', 'nccl'])
    ap.add_argument('--shards', default='contiguous', choices=['contiguous', 'dealt'],
                    help='contiguous ranges of the canonical order, or one piece of every district per rank')
    ap.add_argument('--lost-peer', action='store_true',
                    help='rank 1 never steps: rank 0 must get an error from its next synchronising call, not partial sums')
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    from abm_b200 import distributed, sharding, synth, tables
    from abm_b200.engine import Simulation, load_library, COUNTER_NAMES
    from oracle.abm_oracle import OracleModel
    from util import STATE_KEYS

    rank, world, local = distributed.init_from_env()
    lib = load_library()
    device = torch.device('cuda', local)
    spec = synth.make_spec(n_districts=a.districts, R0=2.6, seed=41)
    pop = synth.make_population(a.agents, spec, seed=41)
    seeds = synth.pick_seeds(pop, spec, infect_num=max(50, a.agents // 400), seed=41)
    tab = tables.build_tables(spec)
    dstart = np.searchsorted(pop.district, np.arange(spec.n_districts + 1)).astype(np.int64)
    uid = distributed.exchange_unique_id(lib, rank, world, device) if (world > 1 and a.exchange == 'nccl') else None
    if a.shards == 'dealt':
        idx = sharding.deal_population(pop, world)[rank]               # global ids of this rank's agents
        shard = sharding.take_population(pop, idx)
        lo, hi = int(idx[0]), int(idx[-1]) + 1
        sim = Simulation(spec, shard, seed=4242, device=local, tables=tab, agent_ids=idx, district_start=dstart,
                         world_size=world, rank=rank, nccl_unique_id=uid, max_days=a.days + 4,
                         exchange=a.exchange if world > 1 else None)
    else:
        lo, hi = sharding.split_population(pop, world)[rank]
        shard = sharding.slice_population(pop, lo, hi)
        idx = np.arange(lo, hi, dtype=np.int64)
        sim = Simulation(spec, shard, seed=4242, device=local, tables=tab, agent_id_base=lo, district_start=dstart,
                         world_size=world, rank=rank, nccl_unique_id=uid, max_days=a.days + 4,
                         exchange=a.exchange if world > 1 else None)
    sim.seed_infections(seeds[np.isin(seeds, idx)])
    if a.lost_peer:
        # a peer that does not publish its table rows: the bounded wait in the table kernel ends, the thresholds of the
        # sub-step are zero, and every synchronising call reports it (sticky)
        ok = True
        if rank == 0:
            sim.step(1)
            errors = []
            for call in (sim.sync, sim.read_log, sim.counters, lambda: sim.step(1)):
                try:
                    call()
                    errors.append(None)
                except RuntimeError as e:
                    errors.append(str(e))
            ok = all(e is not None and 'table exchange failed' in e for e in errors)
            print('lost peer:', errors, flush=True)
        flag = torch.tensor([0 if ok else 1], device=device)
        dist.all_reduce(flag)
        if rank == 0 and int(flag.item()) == 0:
            print('multi_gpu_check ok (lost peer reported)', flush=True)
        sys.stdout.flush()
        os._exit(int(flag.item()))
    sim.step(6 * a.days)

    ora = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob), tab, 4242)
    ora.set_epidemic_status(seeds)
    for _ in range(6 * a.days):
        ora.step()

    got, want = sim.state(), ora.state_dict()
    D = spec.n_districts
    bad = []
    for key in STATE_KEYS:
        g = np.asarray(got[key]).astype(np.int64)
        w = want[key][idx].astype(np.int64)
        if key in ('current_location_ids', 'infected_at_location_ids'):
            g = np.where(g >= D, pop.household[idx] + D, g)          # shard-local household numbering -> global
        if not np.array_equal(g, w):
            bad.append((key, int((g != w).sum())))
    rows = distributed.sum_over_ranks(sim.read_log()[:, :12].copy(), device)
    for r, o in zip(rows, ora.log_rows):
        for i, name in enumerate(COUNTER_NAMES):
            if int(r[i]) != o[name]:
                bad.append((name, o['tick']))
    if len(rows) != len(ora.log_rows):
        bad.append(('log rows', len(rows)))
    # the summed tables every rank holds after the all-reduce equal the oracle's global tables
    rs, nc, th = sim.debug_tables()
    if not (np.array_equal(rs, ora.last_tables['rsum']) and np.array_equal(nc, ora.last_tables['ncnt'])
            and np.array_equal(th, ora.last_tables['thr_out'])):
        bad.append(('tables', 0))
    flag = torch.tensor([len(bad)], device=device)
    if world > 1:
        dist.all_reduce(flag)
    print(f'rank {rank}/{world}: {idx.shape[0]} agents in [{lo}, {hi}) ({a.shards}) infected {int((want["epidemic_state"][idx] > 0).sum())} '
          f'mismatches {bad}', flush=True)
    failed = int(flag.item()) != 0
    if rank == 0 and not failed:
        print(f'multi_gpu_check ok ({a.exchange if world > 1 else "single rank"}): world {world}, {a.agents} agents, {a.days} days, '
              f'infected_count {ora.log_rows[-1]["infected_count"]}', flush=True)

    import threading

    def teardown():
        sim.close()
        if world > 1:
            dist.destroy_process_group()
    t = threading.Thread(target=teardown, daemon=True)
    t.start()
    t.join(timeout=30)
    if t.is_alive():
        print(f'rank {rank}: tear-down stuck, exiting anyway', flush=True)
    sys.stdout.flush()
    os._exit(1 if failed else 0)


if __name__ == '__main__':
    main()
