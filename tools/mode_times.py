"""Per-launch CUDA-event times of the three kernels of a sub-step, by hour of day -- no trace buffer attached, so the sweep
instantiation that runs is the production one (tools/timeline.py attaches one and measures the stamped instantiation).

    python tools/mode_times.py [--agents N] [--spinup D] [--days D] [--seed-fraction F] [--districts D]
                               [--libs a.so,b.so,...]        (build-time variants, tools/build_variants.sh; default: the tree's)

One sub-step per abm_step call in profiling mode (events between the launches: no programmatic overlap), the same timing
`bench.py` uses for `roofline.avg_launch_ms`; then whole days in graph-replay mode (the production path) between one CUDA
event pair.  Several libraries are measured one after the other in this process on the same packed population, and the
final state digest of each is printed: variants of one source must agree bit for bit."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'covid19-agent-based-model_b200')):
    sys.path.insert(0, p)

import numpy as np
import torch

from abm_b200 import engine, synth, tables
from abm_b200.engine import HostStore, Simulation

ap = argparse.ArgumentParser()
ap.add_argument('--agents', type=int, default=15_000_000)
ap.add_argument('--spinup', type=int, default=40)
ap.add_argument('--days', type=int, default=3)
ap.add_argument('--seed-fraction', type=float, default=0.002)
ap.add_argument('--districts', type=int, default=60)
ap.add_argument('--libs', default='')
a = ap.parse_args()
spec = synth.make_spec(n_districts=a.districts, R0=1.9, seed=2020)
pop = synth.make_population(a.agents, spec, seed=2020)
seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=2020)
tab = tables.build_tables(spec)
store = HostStore(spec, pop)
print(f'{a.agents} agents, {a.districts} districts, days {a.spinup}..{a.spinup + 2 * a.days}, seed fraction {a.seed_fraction}; '
      f'us per launch (CUDA events), us per simulated day (graph replay)')
print(f'{"library":14s} | ' + '  '.join(f'{h:02d}h sweep/events' for h in range(0, 24, 4)) + ' | mean sweep  events  table | frac   | day (graph) | state digest')
for path in (a.libs.split(',') if a.libs else [engine.LIB_PATH]):
    engine._lib = None                       # one dlopen per variant (RTLD_LOCAL: the libraries do not see each other)
    engine.load_library(os.path.abspath(path))
    sim = Simulation(spec, pop, seed=2020, tables=tab, max_days=a.spinup + 2 * a.days + 6, host_store=store)
    sim.seed_infections(seeds)
    sim.step(a.spinup * 6)
    sim.sync()
    rows = {}
    for k in range(a.days * 6):
        hour = (sim.step_index * 4) % 24
        sim.set_profiling(True)
        sim.step(1)
        sweep, n = sim.kernel_time()
        rows.setdefault(hour, []).append((sweep / n * 1e3, sim.event_time() * 1e3, sim.table_time() * 1e3))
        sim.set_profiling(False)
    sim.step(6)                              # builds the day graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.sync()
    e0.record(sim.stream)
    sim.step(a.days * 6)
    e1.record(sim.stream)
    sim.sync()
    day_us = e0.elapsed_time(e1) / a.days * 1e3
    sim.flush()
    sim.sync()
    digest = '%x' % (sum(int(sim.buf[k].to(torch.int64).sum()) * (i + 1) for i, k in enumerate(('flags', 'aux', 'cold'))) & (2 ** 64 - 1))
    per_hour = [np.array(rows[h]).mean(0) for h in sorted(rows)]
    m = np.array(per_hour).mean(0)
    frac = 104 / 6 * a.agents / (m[0] * 1e-6) / 6545.9e9
    print(f'{os.path.basename(path):14s} | ' + '  '.join(f'{p[0]:8.1f}/{p[1]:6.1f}' for p in per_hour) +
          f' | {m[0]:10.1f}  {m[1]:6.1f}  {m[2]:5.1f} | {frac:.4f} | {day_us:11.1f} | {digest}', flush=True)
    sim.close()
