"""Time `load_agents` ingestion of the REAL reference and of the drop-in on the same synthetic
census, on this container's CPU, and write profiles/r01_ingestion_cpu.json.

Run in the build container only (needs /root/reference):
    python tests/golden/measure_ref_ingestion.py [n_agents ...]
Both packages are called `covid19_abm`, so each side runs in its own interpreter (`--side ref|ours`).
  ref   ParamsConfig + UnmitigatedScenario + load_agents(census, infect_num=None) of the unmodified reference
        (base_model.py:760-799: pandas -> string-keyed id maps -> per-agent numpy vectors)
  ours  the same calls through the drop-in, plus everything the reference does not have to do before the first step:
        canonical (district, household) ordering, bit packing into the device layout (abm_b200.engine.HostStore).
        The host -> device copy itself is in bench.py's e2e leg (it needs the GPU).
"""
import json
import os
import subprocess
import sys
import tempfile
import time
from datetime import datetime

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [HERE, os.path.join(REPO, 'tests'), os.path.join(REPO, 'covid19-agent-based-model_b200'), REPO]


def world(n_agents):
    from abm_b200 import synth
    spec = synth.make_spec(n_districts=60, R0=1.9, seed=7, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(n_agents, spec, seed=7, scenario_flags=True)
    pop.extras['manufacturing_workers'] = np.where(pop.extras['manufacturing_workers'] == 1, 1.0, np.nan)
    return synth, spec, pop


def side_ref(n_agents):
    import ref_harness as rh
    synth, spec, pop = world(n_agents)
    w = rh.RefWorld(spec, pop, seed_cases=synth.seed_case_counts(pop, spec, seed=7))
    np.random.seed(1)
    t0 = time.perf_counter()
    m = w.make_model('UnmitigatedScenario', seed_infected=25, load=True, infect=False)
    dt = time.perf_counter() - t0
    assert np.asarray(m.household_ids).shape[0] == n_agents
    return dict(seconds=dt)


def side_ours(n_agents):
    import dropin_flow
    from abm_b200 import engine
    synth, spec, pop = world(n_agents)
    root = tempfile.mkdtemp(prefix='abmours_')
    synth.write_data_dir(root, spec, pop, seed=7)
    os.environ['COVID19_ABM_DATA_DIR'] = root
    import torch  # noqa: F401  (import cost is not ingestion)
    from covid19_abm import scenario_models
    packed = {}

    class PackOnly:                                     # stands in for the device: pack the store, upload nothing
        def __init__(self, spec, pop, **kw):
            packed['store'] = engine.HostStore(spec, pop)
    engine.Simulation = PackOnly
    np.random.seed(1)
    t0 = time.perf_counter()
    p = dropin_flow.make_params(R0=1.9, seed_infected=25)
    m = scenario_models.UnmitigatedScenario(p, model_log_file=None, individual_log_file=None)
    m.philox_seed = 1
    m.load_agents(p.data_file_name, size=None, infect_num=None)
    t1 = time.perf_counter()
    m._ensure_device()
    t2 = time.perf_counter()
    assert m.household_ids.shape[0] == n_agents and packed['store'].nbytes > 30 * n_agents
    return dict(seconds=t2 - t0, load_agents_seconds=t1 - t0, canonical_order_and_packing_seconds=t2 - t1)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == '--side':
        fn = side_ref if sys.argv[2] == 'ref' else side_ours
        print('RESULT ' + json.dumps(fn(int(sys.argv[3]))))
        return
    sizes = [int(a) for a in sys.argv[1:]] or [300_000, 1_000_000]
    out = dict(what='load_agents ingestion on the build container CPU (one core), synthetic Zimbabwe-shaped census, '
                    '60 districts; ref = unmodified reference, ours = drop-in incl. canonical ordering and device-layout packing',
               cases=[])
    for n in sizes:
        row = dict(n_agents=n)
        for side in ('ref', 'ours'):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--side', side, str(n)],
                               capture_output=True, text=True, check=True)
            res = json.loads([l for l in r.stdout.splitlines() if l.startswith('RESULT ')][-1][7:])
            row[side] = res
            row[side]['agents_per_second'] = n / res['seconds']
        row['speedup'] = row['ref']['seconds'] / row['ours']['seconds']
        out['cases'].append(row)
        print(json.dumps(row))
    with open(os.path.join(REPO, 'profiles', 'r01_ingestion_cpu.json'), 'w') as fl:
        json.dump(out, fl, indent=1)


if __name__ == '__main__':
    main()
