"""The C-ABI library builds, loads and exports every symbol include/abm_b200.h declares (no compute)."""
import ctypes as ct
import os
import re

from conftest import ROOT


def test_library_exports_all_declared_symbols(built_library):
    header = open(os.path.join(ROOT, 'include', 'abm_b200.h')).read()
    declared = set(re.findall(r'\b(abm_[a-z_]+)\s*\(', header))
    assert len(declared) >= 19
    for name in sorted(declared):
        assert hasattr(built_library, name), name
    assert built_library.abm_version() == 100


def test_python_structs_match_header_layout(built_library):
    from abm_b200 import engine
    assert ct.sizeof(engine.AbmTables) == 16 * 8
    assert ct.sizeof(engine.AbmAgentPtrs) == 11 * 8
    assert engine.AbmConfig.seed.offset % 8 == 0 and engine.AbmConfig.stream.offset % 8 == 0


def test_create_fails_loudly_without_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        return
    from abm_b200 import engine
    cfg = engine.AbmConfig()
    cfg.n_slots, cfg.n_tiles, cfg.n_districts, cfg.n_status, cfg.n_pools, cfg.hw_rows = 1024, 1, 2, 9, 2, 1
    cfg.step_ticks, cfg.max_log_rows, cfg.world_size = 240, 4, 1
    h = ct.c_void_p()
    rc = built_library.abm_create(ct.byref(cfg), ct.byref(h))
    assert rc == -2 and b'no CPU fallback' in built_library.abm_last_error(h)
    built_library.abm_destroy(h)
