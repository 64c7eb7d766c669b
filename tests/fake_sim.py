"""Test double for abm_b200.engine.Simulation backed by the CPU oracle.

TEST INFRASTRUCTURE ONLY: it lets the `-m "not gpu"` suite exercise the host logic of the drop-in `Country`
(person <-> canonical permutation, log writing, scenario clock hooks, refresh, pickling) in a container without
a GPU.  Nothing in the product imports it; tests install it with monkeypatch."""
import copy
from datetime import timedelta

import numpy as np

from abm_b200 import constants as C
from abm_b200.engine import COUNTER_NAMES
from oracle.abm_oracle import OracleModel


class FakeSimulation:
    instances = []

    def __init__(self, spec, pop, seed=0, device=0, tables=None, max_days=400, upload=True, agent_id_base=0, world_size=1,
                 rank=0, **kw):
        self.spec, self.pop, self.tab, self.seed = spec, pop, tables, seed
        self.agent_id_base = int(agent_id_base)
        reduce_fn = None
        if world_size > 1:            # a shard: the district x status tables are summed over the (gloo) ranks
            import torch
            import torch.distributed as dist

            def reduce_fn(buf):
                t = torch.from_numpy(buf.view(np.int64).copy())
                dist.all_reduce(t)
                return t.numpy().view(np.uint64)
        self.ora = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob), tables, seed,
                               agent_id_base=agent_id_base, reduce_fn=reduce_fn)
        self.n_refresh = 0
        self.log_base = 0
        FakeSimulation.instances.append(self)

    @property
    def step_index(self):
        return self.ora.k

    def seed_infections(self, ids):
        self.ora.set_epidemic_status(np.asarray(ids, dtype=np.int64) - self.agent_id_base)      # global -> local ids

    def step(self, n=1):
        for _ in range(int(n)):
            self.ora.step()

    def flush(self):
        pass

    def sync(self):
        pass

    def refresh(self):
        o, pop = self.ora, self.pop
        o.district_mover = pop.district_mover.astype(np.int64).copy()
        o.economic_status_weekday_movement_probability = pop.move_prob_weekday.astype(np.float64).copy()
        o.economic_status_other_day_movement_probability = pop.move_prob_other.astype(np.float64).copy()
        econ = np.where(pop.econ_kind == C.EKIND_HOUSEHOLD, o.household_ids, o.district_ids)
        if pop.econ_pool is not None:
            econ = np.where(pop.econ_kind == C.EKIND_POOL, o.D + o.H + (pop.econ_pool.astype(np.int64) - o.D), econ)
        o.economic_activity_location_ids = econ
        self.n_refresh += 1

    def state(self, household_ids=None, econ_loc=None):
        st = self.ora.state_dict()
        if household_ids is not None:          # oracle numbering (D + dense index) -> the caller's location ids
            o = self.ora
            for key in ('current_location_ids', 'infected_at_location_ids'):
                loc = st[key].copy()
                hh = (loc >= o.D) & (loc < o.D + o.H)
                school = loc >= o.D + o.H
                own = hh & (loc == o.household_ids)
                loc[own] = household_ids[own]
                if econ_loc is not None:
                    loc[school] = econ_loc[school]
                st[key] = loc
        return st

    def read_log(self):
        rows = np.zeros((len(self.ora.log_rows) - self.log_base, 16), dtype=np.int64)
        for i, r in enumerate(self.ora.log_rows[self.log_base:]):
            rows[i, :12] = [r[k] for k in COUNTER_NAMES]
            rows[i, 12] = r['tick']
        return rows

    def log_dicts(self, rows=None):
        rows = self.read_log() if rows is None else rows
        out = []
        for r in rows:
            d = {name: int(r[i]) for i, name in enumerate(COUNTER_NAMES)}
            d['tick'] = int(r[12])
            d['date'] = (self.spec.start_datetime + timedelta(minutes=int(r[12]))).isoformat()
            out.append(d)
        return out

    def counters(self):
        c = self.ora.counters(self.ora.now())
        return {k: c[k] for k in COUNTER_NAMES}

    def active_infections(self):
        return int(np.isin(self.ora.epidemic_state, [1, 2]).sum())

    def district_symptomatic(self):
        o = self.ora
        sick = np.isin(o.epidemic_state, [1, 2]) & (o.infected_symptomatic_status == 1)
        return (np.bincount(o.district_ids[sick], minlength=o.D).astype(np.int64),
                np.bincount(o.district_ids, minlength=o.D).astype(np.int64))

    def checkpoint(self):
        return dict(oracle=copy.deepcopy(self.ora))

    def restore(self, ck):
        self.ora = copy.deepcopy(ck['oracle'])
        self.ora.tab = self.tab
        self.log_base = len(self.ora.log_rows)

    def close(self):
        pass
