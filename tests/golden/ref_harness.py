"""The reference harness lives in baseline/ref_harness.py (bench.py's reference arm uses it too); the fixture generators
of this directory keep importing it under this name."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'baseline', 'ref_harness.py')
_spec = importlib.util.spec_from_file_location('abm_baseline_ref_harness', _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith('__')})
