"""Device-side timeline of a few simulated days (graph replay, PDL): where the time of a sub-step goes.

    python tools/timeline.py [--agents N] [--spinup D] [--days D]

Every kernel stamps [first CTA start, last CTA end] with %globaltimer (abm_set_trace); printed per sub-step: duration of
the sweep, gap to the event kernel, its duration, gap, table kernel, gap to the next sweep (negative = overlap)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'covid19-agent-based-model_b200')):
    sys.path.insert(0, p)

import numpy as np
import torch

from abm_b200 import synth, tables
from abm_b200.engine import Simulation

ap = argparse.ArgumentParser()
ap.add_argument('--agents', type=int, default=15_000_000)
ap.add_argument('--spinup', type=int, default=40)
ap.add_argument('--days', type=int, default=3)
ap.add_argument('--seed-fraction', type=float, default=0.002)
ap.add_argument('--districts', type=int, default=60)
ap.add_argument('--shards', default='contiguous', choices=['contiguous', 'dealt'])
a = ap.parse_args()
spec = synth.make_spec(n_districts=a.districts, R0=1.9, seed=2020)
world = int(os.environ.get('WORLD_SIZE', '1'))
if world > 1:                      # under torchrun: --agents per GPU, weak scaling like bench.py
    from abm_b200 import distributed, sharding
    rank, world, local = distributed.init_from_env()
    torch.cuda.set_device(local)
    if a.shards == 'dealt':
        pop, ids, dstart = sharding.make_synthetic_shard_dealt(a.agents * world, spec, 2020, world, rank)
        seeds = ids[synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=2020 + rank)]
        sim = Simulation(spec, pop, seed=2020, device=local, tables=tables.build_tables(spec), max_days=a.spinup + a.days + 4,
                         agent_ids=ids, district_start=dstart, world_size=world, rank=rank, exchange='p2p')
    else:
        pop, base, dstart = sharding.make_synthetic_shard(a.agents * world, spec, 2020, world, rank)
        seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=2020 + rank) + base
        sim = Simulation(spec, pop, seed=2020, device=local, tables=tables.build_tables(spec), max_days=a.spinup + a.days + 4,
                         agent_id_base=base, district_start=dstart, world_size=world, rank=rank, exchange='p2p')
else:
    rank = 0
    pop = synth.make_population(a.agents, spec, seed=2020)
    seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=2020)
    sim = Simulation(spec, pop, seed=2020, tables=tables.build_tables(spec), max_days=a.spinup + a.days + 4)
n_sub = (a.spinup + a.days) * 6
sim.set_trace(n_sub)
sim.seed_infections(seeds)
sim.step(a.spinup * 6)
torch.cuda.synchronize()
sim.step(a.days * 6)
tr = sim.read_trace(n_sub).astype(np.int64)
dbg = sim.trace_debug.astype(np.int64)
k0 = (a.spinup + 1) * 6          # skip the first traced day after the spin-up boundary
import time
time.sleep(0.5 * rank)           # ranks print one after the other
print(f'rank {rank}/{world} ({a.shards})' if world > 1 else 'single GPU')
print('sub-step  hour | sweep  gap  events  gap  table  gap->next | total   (us)')
tot = []
for k in range(k0, n_sub - 1):
    sw, ev, tb, nxt = tr[k, 0], tr[k, 1], tr[k, 2], tr[k + 1, 0]
    row = [(sw[1] - sw[0]), (ev[0] - sw[1]), (ev[1] - ev[0]), (tb[0] - ev[1]), (tb[1] - tb[0]), (nxt[0] - tb[1]), (nxt[0] - sw[0])]
    tot.append(row)
    print(f'{k:7d}  {(k * 4) % 24:4d} | ' + '  '.join(f'{v / 1e3:6.1f}' for v in row) + '   table-kernel stamps (rel. its start): ' + ' '.join(f'{(d - tb[0]) / 1e3:6.1f}' for d in dbg[k, :4]))
m = np.array(tot, dtype=float).mean(0) / 1e3
print('mean          | ' + '  '.join(f'{v:6.1f}' for v in m), flush=True)
if world > 1:
    torch.distributed.barrier()
    sim.close()
    os._exit(0)
