"""Small driver for ncu: spin a 15 M-agent epidemic up, then run a few simulated days."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'covid19-agent-based-model_b200')):
    sys.path.insert(0, p)

import torch

from abm_b200 import synth, tables
from abm_b200.engine import Simulation

ap = argparse.ArgumentParser()
ap.add_argument('--agents', type=int, default=15_000_000)
ap.add_argument('--spinup', type=int, default=40)
ap.add_argument('--days', type=int, default=2)
ap.add_argument('--seed-fraction', type=float, default=0.002)
ap.add_argument('--districts', type=int, default=60)
a = ap.parse_args()
spec = synth.make_spec(n_districts=a.districts, R0=1.9, seed=2020)
pop = synth.make_population(a.agents, spec, seed=2020)
seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=2020)
sim = Simulation(spec, pop, seed=2020, tables=tables.build_tables(spec), max_days=a.spinup + a.days + 4)
sim.seed_infections(seeds)
sim.step(a.spinup * 6)
torch.cuda.synchronize()
sim.set_profiling(True)
sim.step(a.days * 6)
ms, n = sim.kernel_time()
print('main kernel avg ms', ms / n, 'launches', n, 'step', sim.step_index, sim.counters())
