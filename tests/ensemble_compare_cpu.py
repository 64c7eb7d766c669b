#!/usr/bin/env python
"""Build-container check of the tier-2 tolerance: run the ensemble through the drop-in package with the engine
replaced by the oracle-backed test double (bit-identical to the CUDA path) and print the comparison with the
stored reference ensembles.  Not part of any product path.
    python tests/ensemble_compare_cpu.py unmitigated_100k 64"""
import multiprocessing as mp
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'covid19-agent-based-model_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)


def work(args):
    case, member, root = args
    os.environ['COVID19_ABM_DATA_DIR'] = root
    from abm_b200 import engine
    from fake_sim import FakeSimulation
    engine.Simulation = FakeSimulation
    import ensemble_flow as ef
    ref = ef.load_reference(case)
    curves, dates = ef.run_member(ref, member, os.path.join(root, 'logs', f'm{member}.log'))
    return curves


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else 'unmitigated_100k'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    import ensemble_flow as ef
    ref = ef.load_reference(case)
    root = tempfile.mkdtemp(prefix='abm_ens_')
    ef.build_world(root, ref)
    with mp.get_context('spawn').Pool(8) as pool:
        members = pool.map(work, [(case, m, root) for m in range(n)], chunksize=1)
    res = ef.compare(ref, members)
    for k, v in res.items():
        print(case, k, v)


if __name__ == '__main__':
    main()
