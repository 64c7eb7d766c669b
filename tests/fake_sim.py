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

    def __init__(self, spec, pop, seed=0, device=0, tables=None, max_days=400, upload=True, **kw):
        self.spec, self.pop, self.tab, self.seed = spec, pop, tables, seed
        self.ora = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob), tables, seed)
        self.n_refresh = 0
        self.log_base = 0
        FakeSimulation.instances.append(self)

    @property
    def step_index(self):
        return self.ora.k

    def seed_infections(self, ids):
        self.ora.set_epidemic_status(np.asarray(ids, dtype=np.int64))

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

    def log_dicts(self):
        out = []
        for r in self.ora.log_rows[self.log_base:]:
            d = dict(r)
            d['date'] = (self.spec.start_datetime + timedelta(minutes=int(r['tick']))).isoformat()
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
