"""Where does a simulated day go?  Graph replay vs plain launches vs the sum of the kernels (CUDA events)."""
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
ap.add_argument('--days', type=int, default=20)
ap.add_argument('--seed-fraction', type=float, default=0.002)
a = ap.parse_args()
spec = synth.make_spec(n_districts=60, R0=1.9, seed=2020)
pop = synth.make_population(a.agents, spec, seed=2020)
seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=2020)
tab = tables.build_tables(spec)


def timed(sim, days):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(sim.stream)
    sim.step(days * 6)
    e1.record(sim.stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / days


for use_graph in (True, False):
    sim = Simulation(spec, pop, seed=2020, tables=tab, max_days=a.spinup + 3 * a.days + 8, use_graph=use_graph)
    sim.seed_infections(seeds)
    sim.step(a.spinup * 6)
    day_ms = timed(sim, a.days)
    ck = sim.checkpoint()
    sim.restore(ck, arrays_on_device=True)      # same state again (pending draws were flushed by the checkpoint)
    sim.set_profiling(True)
    sim.step(a.days * 6)
    kms, n = sim.kernel_time()
    tms = sim.table_time()
    sim.set_profiling(False)
    print(f'use_graph={use_graph}: {day_ms:.4f} ms/day; events: agent kernels {kms / a.days:.4f} ms/day ({kms / n * 1e3:.1f} us each), '
          f'table kernels {tms / a.days:.4f} ms/day ({tms / n * 1e3:.1f} us each); unaccounted {day_ms - (kms + tms) / a.days:.4f} ms/day')
    sim.close()
