#!/usr/bin/env python
"""Run one scenario through the drop-in package, on one GPU or -- under torchrun -- on several.

    python scripts/run_scenario.py UnmitigatedScenario Unmitigated 1 1.9 100 2200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        scripts/run_scenario.py UnmitigatedScenario Unmitigated 1 1.9 100 2200

Same arguments as one job of the reference's launcher (scripts/run_simulation_scenarios_full_data.py:41-46 ->
scenario_models.run_scenario, scenario_models.py:443-500): scenario class, simulation name, simulation number, R0,
sample size, number of seed infections.  With several ranks every process executes the host sequence on the full census
(same numpy seed), keeps its household-aligned shard on its GPU (abm_b200/multi.py) and rank 0 writes the log -- the
same file, bit for bit, as the single-GPU run.  Data directory: COVID19_ABM_DATA_DIR (covid19_abm.dir_manager).
"""
import argparse
import os
import sys
from datetime import datetime, timedelta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'covid19-agent-based-model_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('scenario')
    ap.add_argument('sim_name')
    ap.add_argument('sim_num')
    ap.add_argument('R0', type=float)
    ap.add_argument('sample_size', type=int)
    ap.add_argument('seed_num', type=int)
    ap.add_argument('--start-date', default=None, help='ISO date, default params.SIMULATION_START_DATE')
    ap.add_argument('--timestep-hours', type=int, default=4)
    ap.add_argument('--np-seed', type=int, default=None, help='numpy seed (seed agents, lockdown thinning, Philox key)')
    ap.add_argument('--philox-seed', type=int, default=None)
    a = ap.parse_args()

    import numpy as np
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        from abm_b200 import distributed
        distributed.init_from_env()
        if a.np_seed is None:
            raise SystemExit('several ranks need --np-seed: every rank must draw the same seed agents and thinning')
    if a.np_seed is not None:
        np.random.seed(a.np_seed)
    from covid19_abm import scenario_models
    cls = getattr(scenario_models, a.scenario)
    if a.philox_seed is not None:
        cls.philox_seed = a.philox_seed
    model = scenario_models.run_scenario(cls, a.sim_name, a.sim_num, a.R0, a.sample_size, a.seed_num,
                                         start_date=None if a.start_date is None else datetime.fromisoformat(a.start_date),
                                         timestep=timedelta(hours=a.timestep_hours))
    if int(os.environ.get('RANK', '0')) == 0:
        print(f'run_scenario done: {model.scheduler.steps} sub-steps, clock {model.scheduler.real_time.isoformat()}', flush=True)
    model.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
