"""bench.py's CPU arm machinery, without a GPU: the unmodified reference (baseline/_ref, installed by
`__graft_entry__.build()` wherever the reference checkout exists) steps through `baseline/ref_harness.py`, and the numpy
port leg runs.  Tiny populations: this checks plumbing, not speed."""
import hashlib
import json
import os
import subprocess
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
    # in a child process, as bench.py runs it: the harness puts the reference's own `covid19_abm` package on sys.path,
    # which must not shadow the drop-in package of the same name for the tests that follow
    code = ('import json, sys; sys.path.insert(0, %r); import bench; '
            'print(json.dumps(bench._ref_replica((4_000, 12, 0.02, 2, 0, 1, 0))))' % ROOT)
    out = subprocess.run([sys.executable, '-c', code], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    secs, t_load, prev = json.loads(out.strip().splitlines()[-1])
    assert secs > 0 and t_load > 0
    assert prev['agents'] == 4_000 and prev['ever_infected'] >= 80          # at least the seeds (int(p * n) + 1 per district)
    # the sources it ran are the unmodified reference's: same bytes as the manifest says
    man = json.load(open(os.path.join(ROOT, 'baseline', '_ref', 'MANIFEST.json')))
    for rel, digest in man['files'].items():
        if rel == 'MANIFEST.json':
            continue
        assert hashlib.sha256(open(os.path.join(ROOT, 'baseline', '_ref', rel), 'rb').read()).hexdigest() == digest, rel
    ref_file = '/root/reference/src/covid19_abm/base_model.py'
    if os.path.exists(ref_file):
        assert open(ref_file, 'rb').read() == open(os.path.join(ROOT, 'baseline', '_ref', 'src', 'covid19_abm', 'base_model.py'), 'rb').read()


def test_clock_sampler_windows():
    """The SM-clock summary of the bench line: median over the timed region plus the profiled days, first range counted."""
    import bench
    c = bench.ClockSampler(0)
    c.samples, c.max_mhz = [1200, 1965, 1965, 300, 1965, 1950], 1965
    s = c.summary([(1, 3), (4, 6)])
    assert s['sm_mhz'] == 1965.0 and s['samples'] == 4 and s['samples_in_timed_region'] == 2 and s['sm_max_mhz'] == 1965.0
    assert c.summary()['samples'] == 6 and 'window' not in c.summary()
    assert bench.ClockSampler(0).summary([(0, 0), (0, 0)])['sm_mhz'] is None
