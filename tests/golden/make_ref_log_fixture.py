"""Generate tests/golden/ref_log_consumer.json -- what the REAL reference's `load_model_logs_df`
(/root/reference/src/covid19_abm/data_plotting.py:30-64) returns for a daily log written with the reference's own
`log_to_file` (params.py:610-621), for the drop-in's reader to be compared against.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_log_fixture.py
matplotlib / seaborn / geopandas are absent here; the reference module imports them at the top for its plotting
helpers, so empty stand-in modules are registered before the import (the reader itself uses json + pandas only).
"""
import json
import os
import sys
import tempfile
import types
from datetime import datetime, timedelta

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

NAMES = ['current_contagious_count', 'infected_count', 'current_exposed_cases', 'current_contagious_cases',
         'current_hospitalized_cases', 'current_critical_cases', 'asymptomatic_count', 'symptomatic_count',
         'hospitalized_count', 'critical_count', 'died_count', 'recovered_count']


def synthetic_rows(days=40, seed=3):
    """Monotone cumulative counters and bounded current ones, shaped like a small epidemic."""
    rng = np.random.default_rng(seed)
    cum = {n: 0 for n in NAMES}
    rows = []
    t = datetime(2020, 9, 1, 8)
    for d in range(days):
        new = int(rng.poisson(5 + 3 * d))
        cum['infected_count'] += new
        cum['current_contagious_count'] = int(0.8 * cum['infected_count'])
        cum['hospitalized_count'] += int(rng.binomial(new, 0.08))
        cum['critical_count'] += int(rng.binomial(new, 0.02))
        cum['died_count'] += int(rng.binomial(new, 0.01))
        cum['recovered_count'] = int(0.5 * cum['infected_count'])
        cum['current_exposed_cases'] = int(rng.integers(0, new + 1))
        cum['current_contagious_cases'] = cum['current_contagious_count'] - cum['recovered_count']
        cum['current_hospitalized_cases'] = int(rng.integers(0, cum['hospitalized_count'] + 1))
        cum['current_critical_cases'] = int(rng.integers(0, cum['critical_count'] + 1))
        cum['symptomatic_count'] = int(0.6 * cum['infected_count'])
        cum['asymptomatic_count'] = cum['infected_count'] - cum['symptomatic_count']
        rows.append(dict(date=(t + timedelta(days=d)).isoformat(), **{n: cum[n] for n in NAMES}))
    return rows


def main():
    import ref_harness as rh
    rh._install_import_shims()
    for name in ('matplotlib', 'seaborn', 'geopandas'):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.set = lambda *a, **k: None          # data_plotting.py:11 calls sns.set(...)
            sys.modules[name] = mod
    from covid19_abm import data_plotting, params
    assert data_plotting.__file__.startswith('/root/reference/'), data_plotting.__file__
    tmp = tempfile.mkdtemp()
    log = os.path.join(tmp, 'model_log_file_fixture.log')
    for row in synthetic_rows():
        params.log_to_file(log, json.dumps(row), as_log=True, verbose=False)
    df = data_plotting.load_model_logs_df(log)
    lines = [l.split(' $$ ', 1)[1] for l in open(log).read().strip().split('\n')]
    out = dict(log_messages=lines, columns=list(df.columns),
               records=json.loads(df.drop(columns=['version']).to_json(orient='records')))
    with open(os.path.join(HERE, 'ref_log_consumer.json'), 'w') as fl:
        json.dump(out, fl)
    print('wrote', len(out['records']), 'rows; columns', out['columns'])


if __name__ == '__main__':
    main()
