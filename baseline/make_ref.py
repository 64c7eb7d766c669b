#!/usr/bin/env python
"""Install the UNMODIFIED reference under baseline/_ref/ (git-ignored, travels to the GPU box with the working tree).

    python baseline/make_ref.py            # needs /root/reference (the build container)

The reference cannot be pip-installed offline (it imports mesa, pylab and openpyxl, none of which exist here, and reads
private data files), so the install is a plain copy of its package sources plus the one shipped data file the step
needs; `baseline/ref_harness.py` supplies the three import shims (stub mesa / pylab, patched pandas.read_excel) and
synthesises the private inputs.  A manifest with the sha256 of every copied file is written next to them.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/src/covid19_abm'
RISK = '/root/reference/data/preprocessed/risk/hw_and_severe_disease_risk.csv'
DST = os.path.join(HERE, '_ref')


def main():
    if not os.path.isdir(SRC):
        sys.exit(f'{SRC} not found: baseline/_ref can only be built where the reference checkout is')
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, 'src'))
    os.makedirs(os.path.join(DST, 'data'))
    shutil.copytree(SRC, os.path.join(DST, 'src', 'covid19_abm'), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    shutil.copy(RISK, os.path.join(DST, 'data', 'hw_and_severe_disease_risk.csv'))
    manifest = {}
    for root, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, 'rb').read()).hexdigest()
    json.dump(dict(source='/root/reference (worldbank/covid19-agent-based-model), unmodified', files=manifest),
              open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=1)
    print(f'baseline/_ref: {len(manifest)} files')


if __name__ == '__main__':
    main()
