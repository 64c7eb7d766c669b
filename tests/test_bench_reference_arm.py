"""bench.py's CPU arm machinery, without a GPU: the unmodified reference (baseline/_ref, installed by
`__graft_entry__.build()` wherever the reference checkout exists) steps through `baseline/ref_harness.py`, and the numpy
port leg runs.  Tiny populations: this checks plumbing, not speed."""
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)


def test_numpy_port_replica_runs():
    import bench
    secs = bench._cpu_replica((4_000, 12, 0.01, 0, 1, 0))
    assert secs > 0


def test_unmodified_reference_replica_runs():
    import bench
    if not bench._reference_available():
        pytest.skip('baseline/_ref not installed here (needs the reference checkout: python baseline/make_ref.py)')
    secs, t_load, prev = bench._ref_replica((4_000, 12, 0.02, 2, 0, 1, 0))
    assert secs > 0 and t_load > 0
    assert prev['agents'] == 4_000 and prev['ever_infected'] >= 80          # at least the seeds (int(p * n) + 1 per district)
    # the sources it ran are the unmodified reference's: same bytes as the manifest says
    import hashlib
    import json
    man = json.load(open(os.path.join(ROOT, 'baseline', '_ref', 'MANIFEST.json')))
    for rel, digest in man['files'].items():
        if rel == 'MANIFEST.json':
            continue
        assert hashlib.sha256(open(os.path.join(ROOT, 'baseline', '_ref', rel), 'rb').read()).hexdigest() == digest, rel
    ref_file = '/root/reference/src/covid19_abm/base_model.py'
    if os.path.exists(ref_file):
        assert open(ref_file, 'rb').read() == open(os.path.join(ROOT, 'baseline', '_ref', 'src', 'covid19_abm', 'base_model.py'), 'rb').read()
