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
    assert built_library.abm_version() == 200


def test_python_structs_match_header_layout(built_library):
    from abm_b200 import engine
    assert ct.sizeof(engine.AbmTables) == 16 * 8
    assert ct.sizeof(engine.AbmAgentPtrs) == 6 * 8
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


def _config(**kw):
    from abm_b200 import engine
    cfg = engine.AbmConfig()
    cfg.n_slots, cfg.n_tiles, cfg.n_districts, cfg.n_status, cfg.n_pools, cfg.hw_rows = 1024, 1, 2, 9, 2, 1
    cfg.step_ticks, cfg.max_log_rows, cfg.world_size = 240, 4, 1
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def test_create_rejects_bad_configurations(built_library):
    """Argument errors are reported before any device is touched: code ABM_ERR_ARG (-1) + a message on the handle,
    which is still returned so the caller can read it and must destroy it (include/abm_b200.h)."""
    lib = built_library
    cases = [(dict(n_slots=1000), b'n_slots must be n_tiles * 1024'), (dict(n_slots=0, n_tiles=0), b'n_slots'),
             (dict(n_status=0), b'n_status'), (dict(n_status=17), b'n_status'), (dict(n_districts=0), b'district'),
             (dict(n_districts=70000, n_pools=70000), b'district'), (dict(n_pools=1), b'pool'),
             (dict(hw_rows=3), b'hw_rows'), (dict(step_ticks=1), b'step_ticks'), (dict(max_log_rows=0), b'max_log_rows'),
             (dict(n_slots=1 << 31, n_tiles=1 << 21), b'shard the population')]
    for kw, fragment in cases:
        h = ct.c_void_p()
        rc = lib.abm_create(ct.byref(_config(**kw)), ct.byref(h))
        assert rc == -1, kw
        assert h.value and fragment in lib.abm_last_error(h), (kw, lib.abm_last_error(h))
        lib.abm_destroy(h)
    h = ct.c_void_p()
    assert lib.abm_create(None, ct.byref(h)) == -1 and not h.value
    assert lib.abm_create(ct.byref(_config()), None) == -1


def test_null_handles_are_argument_errors(built_library):
    lib = built_library
    assert lib.abm_last_error(None) == b'null handle'
    assert lib.abm_step(None, 1) == -1 and lib.abm_flush(None) == -1 and lib.abm_sync(None) == -1
    assert lib.abm_set_tables(None, None) == -1 and lib.abm_bind_agents(None, None) == -1
    assert lib.abm_seed(None, None, None, 0) == -1 and lib.abm_counters(None, None) == -1
    assert lib.abm_read_log(None, None, 0, None) == -1 and lib.abm_district_symptomatic(None, None, None) == -1
    assert lib.abm_peer_export(None, None) == -1 and lib.abm_peer_attach(None, None, 0) == -1
    assert lib.abm_nccl_unique_id(None) == -1
    lib.abm_destroy(None)          # a no-op, like free(NULL)


def test_integration_guide_names_every_entry_point():
    """INTEGRATION.md maps each C entry point to the reference interface it replaces: none may be missing."""
    header = open(os.path.join(ROOT, 'include', 'abm_b200.h')).read()
    declared = set(re.findall(r'\b(abm_[a-z_]+)\s*\(', header))
    guide = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    mentioned = set(re.findall(r'\babm_[a-z_]+\b', guide))
    assert declared - mentioned == set(), sorted(declared - mentioned)
