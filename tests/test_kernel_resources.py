"""Register / stack budget of the built kernels, read from the library without a GPU (`cuobjdump -res-usage`).

Round 2 measured what a spill costs here (profiles/r02b_variants_*.txt, r02c_*): two spilled registers in the night sweep
were 3.7 % of the mean sweep launch, 236 spilled bytes per thread in the event kernel 14 % of that kernel at 08:00.  The
shapes below are the ones the screens selected; a change that pushes a production kernel back over its budget shows up
here before it shows up on a B200."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import PKG

CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'


def _resources(built_library):
    out = subprocess.run([CUOBJDUMP, '-res-usage', os.path.join(PKG, 'libabm_b200.so')], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for line in out.splitlines():
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r'\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', line)
        if m and name:
            res[name] = dict(zip(('reg', 'stack', 'shared', 'local'), map(int, m.groups())))
            name = None
    return res


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason='cuobjdump not available')
def test_production_kernels_stay_inside_their_register_budget_without_spills(built_library):
    res = _resources(built_library)
    sweeps = {k: v for k, v in res.items() if 'abm_sweep_kernelILj' in k}
    assert len(sweeps) == 10                                  # 5 modes x (production, TRACE)
    for name, r in sweeps.items():
        mode = int(re.search(r'ILj(\d+)ELb([01])E', name).group(1))
        trace = re.search(r'ILj(\d+)ELb([01])E', name).group(2) == '1'
        if trace:
            continue                                          # the stamped instantiation only runs under abm_set_trace
        budget = 32 if mode in (259, 451) else 48             # night / noon: 8 CTAs per SM; movement and generic: 5
        assert r['reg'] <= budget and r['stack'] == 0 and r['local'] == 0, (name, r)
    event = [v for k, v in res.items() if 'abm_event_kernel' in k]
    assert len(event) == 1 and event[0]['reg'] <= 64 and event[0]['stack'] == 0 and event[0]['local'] == 0, event
