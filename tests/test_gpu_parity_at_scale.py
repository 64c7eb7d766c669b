"""CUDA path == numpy oracle, bit for bit, at the sizes BASELINE.json names (missing item 6):
  * configs[1]: 15 M agents, 60 districts (state vectors, tables and day log after 3 days at 1 % seeded),
  * configs[4]'s shape: 500 districts, millions of agents,
  * configs[4]'s per-rank id range: canonical ids beyond 2^32 as Philox keys (a shard of a 1 B+ population).
The oracle (oracle/abm_oracle.py, test infrastructure) does ~1.5e6 agent-days/s on one host core, so these take about a
minute each; everything else in the suite runs at sizes it finishes in seconds."""
import time

import numpy as np
import pytest

from abm_b200 import synth, tables
from oracle.abm_oracle import OracleModel
from util import assert_state_equal, oracle_params

pytestmark = pytest.mark.gpu


def _compare(spec, pop, seeds, days, seed, agent_id_base=0, check_days=None, label=''):
    from abm_b200.engine import Simulation
    tab = tables.build_tables(spec)
    t0 = time.time()
    sim = Simulation(spec, pop, seed=seed, tables=tab, max_days=days + 4, agent_id_base=agent_id_base,
                     district_start=np.searchsorted(pop.district, np.arange(spec.n_districts + 1)).astype(np.int64) + agent_id_base)
    ora = OracleModel(pop, oracle_params(spec), tab, seed, agent_id_base=agent_id_base)
    sim.seed_infections(seeds + agent_id_base)
    ora.set_epidemic_status(seeds)
    check_days = check_days or [days - 1]
    for day in range(days):
        sim.step(6)
        for _ in range(6):
            ora.step()
        if day in check_days:
            assert_state_equal(sim.state(), ora.state_dict(), f'{label} day {day}')
            r, n, t = sim.debug_tables()
            assert np.array_equal(r, ora.last_tables['rsum']) and np.array_equal(n, ora.last_tables['ncnt'])
            assert np.array_equal(t, ora.last_tables['thr_out'])
    got = sim.log_dicts()
    assert len(got) == len(ora.log_rows)
    for g, w in zip(got, ora.log_rows):
        for key, val in w.items():
            assert g[key] == val, (label, key, g['date'])
    infected = got[-1]['infected_count']
    sim.close()
    print(f'{label}: {pop.n} agents, {spec.n_districts} districts, {days} days, infected {infected}, {time.time() - t0:.0f} s')
    return infected


def test_config1_15m_agents_60_districts(built_library):
    """BASELINE configs[1] size: the population bench.py times, 1 % seeded, 3 days (two movement sub-steps a day, tens of
    thousands of infection events, transitions, trips)."""
    spec = synth.make_spec(n_districts=60, R0=1.9, seed=2020)
    pop = synth.make_population(15_000_000, spec, seed=2020)
    seeds = synth.pick_seeds(pop, spec, infect_num=150_000, seed=2020)
    infected = _compare(spec, pop, seeds, days=3, seed=2020, label='15M/60')
    assert infected > 150_000


def test_500_districts_4m_agents(built_library):
    """BASELINE configs[4] shape: 500 districts (OD rows of 500 columns, 5 000 pool x status cells), 4 M agents."""
    spec = synth.make_spec(n_districts=500, R0=1.9, seed=11)
    pop = synth.make_population(4_000_000, spec, seed=11)
    seeds = synth.pick_seeds(pop, spec, infect_num=40_000, seed=11)
    infected = _compare(spec, pop, seeds, days=3, seed=5, label='4M/500')
    assert infected > 40_000


def test_shard_with_ids_beyond_32_bits(built_library):
    """A rank of the 1 B-agent run holds canonical ids far beyond 2^30; here a 2 M-agent shard whose ids start at
    5 * 2^32 + 12 (both Philox counter words in use, base not a multiple of 4: leading filler slots)."""
    spec = synth.make_spec(n_districts=40, R0=2.2, seed=13)
    pop = synth.make_population(2_000_000, spec, seed=13)
    seeds = synth.pick_seeds(pop, spec, infect_num=20_000, seed=13)
    base = 5 * 2**32 + 12 + 1          # odd on purpose
    _compare(spec, pop, seeds, days=4, seed=9, agent_id_base=base, check_days=[1, 3], label='2M ids>2^32')
