"""Per-day device time of the agent step over a whole epidemic (CUDA events per simulated day)."""
import argparse
import json
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
ap.add_argument('--days', type=int, default=365)
ap.add_argument('--seed-fraction', type=float, default=0.00015)
ap.add_argument('--out', default=None)
a = ap.parse_args()

spec = synth.make_spec(n_districts=60, R0=1.9, seed=2020)
pop = synth.make_population(a.agents, spec, seed=2020)
seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=2020)
sim = Simulation(spec, pop, seed=2020, tables=tables.build_tables(spec), max_days=a.days + 4)
sim.seed_infections(seeds)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.days + 1)]
ev[0].record(sim.stream)
for d in range(a.days):
    sim.step(6)
    ev[d + 1].record(sim.stream)
torch.cuda.synchronize()
ms = np.array([ev[d].elapsed_time(ev[d + 1]) for d in range(a.days)])
log = sim.log_dicts()
active = np.array([r['current_contagious_cases'] for r in log])
print('total ms', ms.sum(), 'agent-days/s whole run', a.agents * a.days / ms.sum() * 1e3)
print('peak active', active.max(), 'on day', int(active.argmax()), 'final', {k: log[-1][k] for k in ('infected_count', 'died_count', 'recovered_count')})
for d in range(0, a.days, max(1, a.days // 24)):
    print(f'day {d:3d}  {ms[d]:.3f} ms  active {active[min(d, len(active) - 1)]}')
if a.out:
    json.dump(dict(ms=ms.tolist(), active=active.tolist(), agents=a.agents), open(a.out, 'w'))
