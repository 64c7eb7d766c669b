"""Error behaviour of the C ABI on a live device (include/abm_b200.h: every call returns 0 or a negative abm_status
and leaves a message on the handle; a refused call leaves the simulation in a consistent state)."""
import ctypes as ct

import numpy as np
import pytest

from util import assert_state_equal, small_case

pytestmark = pytest.mark.gpu

ERR_ARG, ERR_STATE = -1, -3


def _bare_handle(lib, **kw):
    from abm_b200 import engine
    cfg = engine.AbmConfig()
    cfg.n_slots, cfg.n_tiles, cfg.n_districts, cfg.n_status, cfg.n_pools, cfg.hw_rows = 1024, 1, 2, 9, 2, 1
    cfg.step_ticks, cfg.max_log_rows, cfg.world_size, cfg.n_move_classes = 240, 4, 1, 1
    for k, v in kw.items():
        setattr(cfg, k, v)
    h = ct.c_void_p()
    rc = lib.abm_create(ct.byref(cfg), ct.byref(h))
    return rc, h


def test_calls_out_of_order_are_state_errors(built_library):
    import torch
    from abm_b200 import engine
    lib = built_library
    rc, h = _bare_handle(lib)
    assert rc == 0, lib.abm_last_error(h)
    assert lib.abm_step(h, 1) == ERR_STATE and b'tables / agents not set' in lib.abm_last_error(h)
    one = np.zeros(1, dtype=np.int64)
    assert lib.abm_seed(h, one.ctypes.data_as(ct.c_void_p), one.ctypes.data_as(ct.c_void_p), 1) == ERR_STATE
    assert lib.abm_step(h, -1) == ERR_ARG
    assert lib.abm_step(h, 0) == ERR_STATE            # still nothing bound
    assert lib.abm_peer_export(h, ct.cast((ct.c_char * 64)(), ct.c_void_p)) == ERR_STATE      # single-rank handle
    assert b'exchange' in lib.abm_last_error(h)
    assert lib.abm_peer_detach(h) == 0                # nothing attached: a no-op
    # agent buffers: null and misaligned pointers are refused
    ap = engine.AbmAgentPtrs()
    assert lib.abm_bind_agents(h, ct.byref(ap)) == ERR_ARG and b'null agent buffer' in lib.abm_last_error(h)
    buf = torch.zeros(1024 * 16, dtype=torch.int32, device='cuda')
    for name, _ in engine.AbmAgentPtrs._fields_:
        setattr(ap, name, buf.data_ptr())
    ap.flags = buf.data_ptr() + 4
    assert lib.abm_bind_agents(h, ct.byref(ap)) == ERR_ARG and b'16-byte aligned' in lib.abm_last_error(h)
    ap.flags = buf.data_ptr()
    ap.move_class = buf.data_ptr() + 2
    assert lib.abm_bind_agents(h, ct.byref(ap)) == ERR_ARG and b'misaligned' in lib.abm_last_error(h)
    assert lib.abm_sync(h) == 0
    lib.abm_destroy(h)


def test_multi_rank_configuration_errors(built_library):
    lib = built_library
    rc, h = _bare_handle(lib, world_size=2, rank=0, exchange=0)          # NCCL exchange without a unique id
    assert rc == ERR_ARG and b'ncclUniqueId' in lib.abm_last_error(h)
    lib.abm_destroy(h)
    rc, h = _bare_handle(lib, world_size=64, rank=0, exchange=1)         # more ranks than peer slots
    assert rc == ERR_ARG and b'at most' in lib.abm_last_error(h)
    lib.abm_destroy(h)
    rc, h = _bare_handle(lib, world_size=2, rank=0, exchange=1)          # peer exchange: attach is mandatory
    assert rc == 0, lib.abm_last_error(h)
    blob = (ct.c_char * 64)()
    assert lib.abm_peer_export(h, ct.cast(blob, ct.c_void_p)) == 0 and any(bytes(blob))
    three = (ct.c_char * (64 * 3))()
    assert lib.abm_peer_attach(h, ct.cast(three, ct.c_void_p), 3) == ERR_ARG
    assert b'3 handles for world size 2' in lib.abm_last_error(h)
    lib.abm_destroy(h)


def test_full_day_log_refuses_further_steps_and_keeps_the_state(built_library):
    from abm_b200.engine import Simulation
    spec, pop, seeds, tab = small_case(n_agents=6000, n_districts=3, infect_num=80)
    sim = Simulation(spec, pop, seed=11, tables=tab, max_days=2)         # room for 4 log rows
    sim.seed_infections(seeds)
    with pytest.raises(RuntimeError, match='day log full'):
        sim.step(6 * 10)
    k = sim.step_index
    rows = sim.read_log()
    assert rows.shape[0] == 4 and 6 * 3 < k <= 6 * 5
    with pytest.raises(RuntimeError, match='day log full'):              # still refused, nothing moves
        sim.step(6)
    assert sim.step_index == k
    sim.step(0)
    got = sim.state()
    twin = Simulation(spec, pop, seed=11, tables=tab, max_days=20)
    twin.seed_infections(seeds)
    twin.step(k)
    assert_state_equal(got, twin.state(), where=f'after the refused call at sub-step {k}')
    assert np.array_equal(rows, twin.read_log()[:4])
    sim.close(); twin.close()


def test_debug_tables_need_a_step_and_unknown_exchange_is_rejected(built_library):
    from abm_b200.engine import Simulation
    spec, pop, seeds, tab = small_case(n_agents=3000, n_districts=2, infect_num=20)
    sim = Simulation(spec, pop, seed=1, tables=tab, max_days=3)
    with pytest.raises(RuntimeError, match='no sub-step executed yet'):
        sim.debug_tables()
    sim.step(1)
    sim.debug_tables()
    sim.close()
    with pytest.raises(ValueError, match='unknown exchange'):
        Simulation(spec, pop, seed=1, tables=tab, exchange='mpi')


@pytest.mark.parametrize('k_ckpt,k_used', [(9, 17), (10, 16), (12, 6)])
def test_restore_onto_a_handle_that_has_already_run(built_library, k_ckpt, k_used):
    """abm_set_state is the general resume call (include/abm_b200.h): restoring a checkpoint onto a handle that has
    executed other sub-steps -- same or different parity of the sub-step index, mid-day or not -- continues exactly like
    the uninterrupted run (stale accumulators / a reset exchange generation used to leak in)."""
    from abm_b200.engine import Simulation
    spec, pop, seeds, tab = small_case(n_agents=20_000, n_districts=5, infect_num=400)
    ref = Simulation(spec, pop, seed=21, tables=tab, max_days=12)
    ref.seed_infections(seeds)
    ref.step(k_ckpt)
    ck = ref.checkpoint()
    ref.step(30)
    want, want_tables = ref.state(), ref.debug_tables()

    used = Simulation(spec, pop, seed=21, tables=tab, max_days=12)
    used.seed_infections(seeds[: len(seeds) // 2])       # a different history on this handle
    used.step(k_used)
    used.restore(ck)
    assert used.step_index == k_ckpt
    used.step(30)
    assert_state_equal(used.state(), want, where=f'restored at {k_ckpt} onto a handle at {k_used}')
    for g, w in zip(used.debug_tables(), want_tables):
        assert np.array_equal(g, w)
    ref.close(); used.close()


def test_active_infections_is_the_stop_test_without_a_flush(built_library):
    """abm_active_infections (run_scenario's stop test, scenario_models.py:494-495): agrees with the flushed state on
    zero / non-zero after every sub-step, launches nothing and leaves the infection draw pending."""
    from abm_b200.engine import Simulation
    spec, pop, seeds, tab = small_case(n_agents=5_000, n_districts=3, infect_num=3, R0=0.3)
    sim = Simulation(spec, pop, seed=3, tables=tab, max_days=80)
    twin = Simulation(spec, pop, seed=3, tables=tab, max_days=80)
    assert sim.active_infections() == 0
    sim.seed_infections(seeds); twin.seed_infections(seeds)
    seen_zero = False
    for k in range(6 * 70):
        sim.step(1); twin.step(1)
        launches = sim.launch_count
        n = sim.active_infections()
        assert sim.launch_count == launches
        epi = np.asarray(twin.state()['epidemic_state'])          # flushes the twin: the reference's view
        assert (n == 0) == (not np.isin(epi, [1, 2]).any()), k
        if n == 0:
            seen_zero = True
            break
    assert seen_zero, 'the sub-critical epidemic should have died out'
    assert_state_equal(sim.state(), twin.state(), where='unflushed vs flushed every sub-step')
    sim.close(); twin.close()
