"""Ensembles as concurrent replicas on one GPU.

The reference runs an ensemble as a pool of OS processes, one single-threaded Python process per
(scenario, R0, simulation) (scripts/run_simulation_scenarios_full_data.py:35-82).  Here the members of an ensemble
are `Simulation` handles that share one packed host store and differ only in the Philox key and the seed agents;
each has its own CUDA stream, so their (small) kernels overlap on the device.  A 100 k-agent population fills about
a twelfth of a B200, which is why replicas -- not a bigger grid -- are the way to use the rest of it.
"""
import numpy as np

from . import tables as tables_mod
from .engine import COUNTER_NAMES, HostStore, Simulation


def run_ensemble(spec, pop, seed_agents, philox_seeds, days, device=0, max_concurrent=16, tables=None):
    """Run len(philox_seeds) members for `days` simulated days.

    seed_agents   list (one per member) of canonical agent ids to infect at the start, or one array used by all
    philox_seeds  list of 64-bit draw keys, one per member
    Returns int64 [members, log rows, 12] (columns in engine.COUNTER_NAMES order) and the list of ISO dates."""
    import torch
    tab = tables if tables is not None else tables_mod.build_tables(spec)
    store = HostStore(spec, pop)
    steps_per_day = (24 * 60) // spec.step_minutes
    n = len(philox_seeds)
    if not isinstance(seed_agents, (list, tuple)):
        seed_agents = [seed_agents] * n
    out, dates = [None] * n, None
    for lo in range(0, n, max_concurrent):
        group = []
        for m in range(lo, min(n, lo + max_concurrent)):
            sim = Simulation(spec, pop, seed=int(philox_seeds[m]), device=device, tables=tab, max_days=days + 2,
                             host_store=store)
            sim.seed_infections(np.asarray(seed_agents[m], dtype=np.int64))
            group.append(sim)
        for _ in range(days):                       # round robin: every member's day is enqueued on its own stream
            for sim in group:
                sim.step(steps_per_day)
        for i, sim in enumerate(group):
            rows = sim.read_log()
            out[lo + i] = rows[:, :len(COUNTER_NAMES)].copy()
            if dates is None:
                dates = [d['date'] for d in sim.log_dicts(rows)]
            sim.close()
        torch.cuda.synchronize(device)
    return np.stack(out), dates


def summarize(curves, q=(2.5, 97.5)):
    """Mean and central band over the members axis of `run_ensemble` output."""
    c = np.asarray(curves, dtype=np.float64)
    return dict(mean=c.mean(axis=0), lo=np.percentile(c, q[0], axis=0), hi=np.percentile(c, q[1], axis=0))
