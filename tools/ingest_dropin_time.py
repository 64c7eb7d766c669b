"""Ingestion through the reference's own API at scale: Country.load_agents on a census pickle in the REFERENCE format
(string household / district / status / location ids, base_model.py:760-799) -> first step().

    python tools/ingest_dropin_time.py --agents 15000000

Prints the stages: read_pickle, name resolution + vectors (host, pandas), seeding = device store build (abm_b200/ingest.py)
+ first kernel, first step."""
import argparse
import json
import os
import sys
import tempfile
import time
from datetime import datetime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'covid19-agent-based-model_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)

import numpy as np
import pandas as pd

ap = argparse.ArgumentParser()
ap.add_argument('--agents', type=int, default=15_000_000)
ap.add_argument('--out', default=None)
a = ap.parse_args()
import dropin_flow
root = tempfile.mkdtemp(prefix='abm_ingest_')
t0 = time.time()
dropin_flow.build_world(root, n_agents=a.agents, n_districts=60, shuffle=True)
print(f'synthetic census written in {time.time() - t0:.1f} s (untimed set-up)', flush=True)
import torch
torch.zeros(1, device='cuda')                      # CUDA context: not part of ingestion
from covid19_abm import scenario_models
np.random.seed(5)
p = dropin_flow.make_params(R0=1.9, seed_infected=int(0.00015 * a.agents))
m = scenario_models.UnmitigatedScenario(p, model_log_file=None)
m.philox_seed = 1
p.SIMULATION_START_DATE = datetime(2020, 9, 1, 0)
m.scheduler.real_time = p.SIMULATION_START_DATE
stages = {}
t = time.time()
df = pd.read_pickle(p.data_file_name)
stages['read_pickle_s'] = time.time() - t
del df
t = time.time()
m.load_agents(p.data_file_name, size=None, infect_num=p.SEED_INFECT_NUM)        # includes its own read_pickle
torch.cuda.synchronize()
stages['load_agents_s'] = time.time() - t
t = time.time()
m.step()
m._sim.sync()
stages['first_step_s'] = time.time() - t
stages['dataframe_to_first_step_s'] = stages['load_agents_s'] - stages['read_pickle_s'] + stages['first_step_s']
stages['agents'] = a.agents
stages['agents_per_s_from_dataframe'] = a.agents / stages['dataframe_to_first_step_s']
print(json.dumps(stages))
if a.out:
    json.dump(stages, open(a.out, 'w'))
m.close()
