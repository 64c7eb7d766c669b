"""`Country` and `EpidemicScheduler` -- drop-in for the reference's model facade
(/root/reference/src/covid19_abm/base_model.py:20-441 scheduler, 444-1106 model), with the daily agent
step executed on a B200 through libabm_b200.so.

What stays as in the reference: class / method / attribute names, constructor arguments, the
`load_agents` flow with its two scenario hooks, the helper methods scenario classes call
(`set_lockdown_movers_and_movement_probabilities`, `setup_selective_movement_restriction_scenarios`,
school helpers, `get_safe_and_unsafe_districts`), the public per-agent numpy vectors and the
`STATE_* / CLINICAL_* / SYMPTOM_* / DISTRICT_MOVER_*` constants, the daily `$$`-delimited JSON log line,
picklability.  Reference quirks that change results are kept: `model.lockdown` /
`model.blocked` are read before the scenario sets them (Q1), `severe_disease_risk` is all ones (Q2).

What is different: the agent vectors live on the GPU in the packed layout of abm_b200/layout.py.  The
device store is created at the first `step()` / `set_epidemic_status()` from the host vectors the load
built (so every load-time hook has run), the dynamic vectors (`epidemic_state`, `current_location_ids`,
`date_*` ...) are materialised from the device on attribute access, and the four vectors scenario code
mutates later (`district_mover`, the two movement probabilities, `economic_activity_location_ids`) are
re-encoded and pushed by `refresh_device()` -- the school helpers call it themselves.  There is no CPU
path: without the CUDA library or a GPU, `step()` raises.
"""
import gc
import json
from datetime import datetime, timedelta

import numpy as np
import pandas as pd

from abm_b200 import constants as C
from abm_b200.spec import ModelSpec, Population
from covid19_abm.params import log_to_file

T_NEVER = int(C.T_NEVER)

# vectors that the agent step changes: served from the device once it exists
_TICK_VECTORS = ('date_infected', 'date_start_contagious', 'date_start_symptomatic', 'date_recovered',
                 'date_start_hospitalized', 'date_end_hospitalized', 'date_start_critical', 'date_end_critical',
                 'date_died', 'return_district_at')
_STATE_VECTORS = ('epidemic_state', 'clinical_state', 'infected_symptomatic_status', 'current_location_ids',
                  'current_district_ids', 'infected_at_district_ids', 'infected_at_location_ids')
_DERIVED_VECTORS = ('infection_rate', 'mild_symptom_movement_probability')
_DYNAMIC = _TICK_VECTORS + _STATE_VECTORS + _DERIVED_VECTORS
_COUNTER_KEYS = ('current_contagious_count', 'infected_count', 'current_exposed_cases', 'current_contagious_cases',
                 'current_hospitalized_cases', 'current_critical_cases', 'asymptomatic_count', 'symptomatic_count',
                 'hospitalized_count', 'critical_count', 'died_count', 'recovered_count')     # order of base_model.py:1086-1104


def _dynamic_property(name):
    def get(self):
        return self._dynamic_vector(name)

    def put(self, value):
        self._host[name] = value
        self._fresh.pop(name, None)
    return property(get, put, doc=f'per-agent vector `{name}` (person order); read from the device after the first step')


class EpidemicScheduler:
    """Clock + `step()` of the reference's scheduler (base_model.py:20-179).  One call = one sub-step
    (`params.step_timedelta`); the work is `abm_step(handle, 1)`."""

    def __init__(self, model):
        self.model = model
        self.steps = 0
        self.time = 0
        self._real_time = model.start_datetime
        self.step_timedelta = model.step_timedelta
        p = model.params
        self.weekday_move_hours = {p.WEEKDAY_START_DAY_HOUR, p.WEEKDAY_END_DAY_HOUR}
        self.other_day_move_hours = {p.OTHER_DAY_START_DAY_HOUR, p.OTHER_DAY_END_DAY_HOUR}
        self.day_location_map = {}

    @property
    def real_time(self):
        return self._real_time

    @real_time.setter
    def real_time(self, value):
        if self.model._sim is not None:
            raise RuntimeError('scheduler.real_time can only be set before the first step (the device clock is running)')
        self._real_time = value

    def step(self):
        m = self.model
        sim = m._ensure_device()
        logs = self._real_time.hour == 8                     # base_model.py:1083
        sim.step(1)
        self.steps += 1
        self.time += 1
        self._real_time = self._real_time + self.step_timedelta          # :177-179
        m._fresh.clear()
        if logs:
            m._log_rows_due += 1
            if m.model_log_file is not None and m._log_rows_due - m._log_rows_written >= m.log_flush_days:
                m.log_model_output()

    # ---- the pieces of the reference's step are phases of one fused kernel here
    def _fused(self, name):
        raise NotImplementedError(f'EpidemicScheduler.{name} is a phase of abm_sweep_kernel (csrc/abm_kernels.cuh) and '
                                  'cannot be run on its own; call step()')

    def update_agent_states(self):
        self._fused('update_agent_states')

    def go_home(self):
        self._fused('go_home')

    def go_out(self):
        self._fused('go_out')

    def return_to_home_district(self, mover_ids):
        self._fused('return_to_home_district')

    def move_to_other_district(self, mover_ids):
        self._fused('move_to_other_district')

    def get_active_ids(self, return_type=None):
        """People that may move: not hospitalised (or recovered), at a non-negative location (base_model.py:195-209)."""
        m = self.model
        ok = ((m.clinical_state < m.CLINICAL_HOSPITALIZED) | (m.epidemic_state == m.STATE_RECOVERED)) & (m.current_location_ids >= 0)
        ids = np.nonzero(ok)[0]
        return ids if return_type is None else return_type(ids)

    def get_infection_rates(self, contagious_ids, location):
        """Per-contact infection rates of `contagious_ids` (base_model.py:181-193): x3 inside a household, x the
        hand-washing multiplier of the first agent's district in the HANDWASHING_RISK scenarios."""
        m = self.model
        contagious_ids = np.asarray(contagious_ids)
        rates = m.infection_rate[contagious_ids].reshape(contagious_ids.size, 1)
        if location == 'house':
            rates = rates * 3
        if m.params.SCENARIO.startswith('HANDWASHING_RISK'):
            rates = rates * m.hw_risk[m.current_district_ids[contagious_ids[0]]]
        return rates


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class _IdNameTable(dict):
    """`{id: name}` (or, inverted, `{name: id}`) of millions of consecutive ids, as the plain dict the reference keeps
    (base_model.py:577-611), but filled on first use: building 3.7 M-entry dictionaries costs seconds at load time and
    nothing in the step reads them.  A real dict subclass, so every dict operation works once it is filled."""

    _pending = None

    def __init__(self, names, first_id, inverted=False, extra=()):
        super().__init__()
        self._pending = (names, int(first_id), bool(inverted), tuple(extra))

    def __reduce__(self):                       # pickles (and deep-copies) as the plain dict it stands for
        return (dict, (dict(self._fill()),))

    def _fill(self):
        if self._pending is not None:
            (names, first, inverted, extra), self._pending = self._pending, None
            for e in extra:                      # earlier tables first: later entries win, like dict.update
                dict.update(self, e)
            seq = names.tolist() if hasattr(names, 'tolist') else list(names)
            dict.update(self, zip(seq, range(first, first + len(seq))) if inverted else enumerate(seq, first))
        return self

    def last_id(self):
        """max(keys) of an id -> name table without filling it."""
        if self._pending is not None and not self._pending[2] and not self._pending[3]:
            return self._pending[1] + len(self._pending[0]) - 1
        return max(self._fill().keys())


def _forward(name):
    def method(self, *a, **kw):
        return getattr(dict, name)(self._fill(), *a, **kw)
    method.__name__ = name
    return method


for _n in ('__getitem__', '__iter__', '__len__', '__contains__', '__eq__', '__repr__', 'get', 'items', 'keys', 'values',
           '__setitem__', '__delitem__', 'update', 'pop', 'copy', '__or__', '__ror__', 'setdefault'):
    setattr(_IdNameTable, _n, _forward(_n))


class _InvertedLocations(_IdNameTable):
    """`{location name: id}` = inversion of districts + households (+ schools) merged in that order."""

    def __init__(self, district_id_to_name, hh_names, first_household, schools):
        dict.__init__(self)
        self._pending = (district_id_to_name, hh_names, int(first_household), schools)

    def _fill(self):
        if self._pending is not None:
            (districts, hh_names, first, schools), self._pending = self._pending, None
            dict.update(self, {n: i for i, n in districts.items()})
            seq = hh_names.tolist() if hasattr(hh_names, 'tolist') else list(hh_names)
            dict.update(self, zip(seq, range(first, first + len(seq))))
            for sc in schools:
                dict.update(self, {n: i for i, n in sc.items()})
        return self

    def last_id(self):
        return max(self._fill().values())


def _last_id(table):
    return table.last_id() if isinstance(table, _IdNameTable) else max(table)


class _Never:
    """Stand-in for `np.full(size, datetime.max)` (object dtype: 8 bytes and a reference count per element) until
    somebody reads the vector before the device store exists."""

    def __init__(self, size, value):
        self.size, self.value = size, value

    def make(self):
        return np.full(shape=self.size, fill_value=self.value)


class Country:
    STATE_SUSCEPTIBLE, STATE_INFECTED, STATE_CONTAGIOUS, STATE_RECOVERED, STATE_DEAD = 0, 1, 2, 3, 4
    DEAD_LOCATION_ID = -1
    CLINICAL_NOT_HOSPITALIZED, CLINICAL_HOSPITALIZED, CLINICAL_CRITICAL, CLINICAL_RELEASED_OR_DEAD = 0, 1, 2, 3
    SYMPTOM_NONE, SYMPTOM_ASYMPTOMATIC, SYMPTOM_SYMPTOMATIC = -1, 0, 1
    DISTRICT_MOVER_FALSE, DISTRICT_MOVER_TRUE = 0, 1

    # device options (class-level so scenario subclasses and run_scenario need no extra arguments)
    device_index = 0
    philox_seed = None            # None: drawn from numpy's global RNG at device creation (so np.random.seed pins a run)
    log_flush_days = 1            # write log lines every N simulated days (each flush synchronises with the GPU)
    max_days = 800
    _sim_class = None             # tests substitute the engine (CPU double); None = abm_b200.engine.Simulation
    ingest_on_device = True       # build the packed agent store on the GPU (abm_b200/ingest.py) instead of in numpy

    def __init__(self, params, model_log_file=None, individual_log_file=None):
        self.individual_log_file = individual_log_file       # accepted and never written, as in the reference
        self.model_log_file = model_log_file
        self.params = params
        self.start_datetime = params.start_datetime
        self.step_timedelta = params.step_timedelta

        self._host = {name: None for name in _DYNAMIC}        # host copies of the dynamic vectors (valid until the device exists)
        self._fresh = {}                                       # dynamic vectors already read back since the last step
        self._sim = None
        self._dirty = False
        self._log_rows_due = 0
        self._log_rows_written = 0
        self._restore = None
        self._status_person_ids = None

        for name in ('person_ids', 'district_ids', 'household_ids', 'age', 'sex', 'economic_status_ids',
                     'economic_activity_location_ids', 'left_district_at', 'economic_status_weekday_movement_probability',
                     'economic_status_other_day_movement_probability', 'district_mover', 'severe_disease_risk'):
            setattr(self, name, None)
        # running totals the reference declares and never updates (base_model.py:510-516)
        self.infected_count = self.asymptomatic_count = self.symptomatic_count = 0
        self.hospitalized_count = self.critical_count = self.died_count = self.recovered_count = 0

        self.scheduler = EpidemicScheduler(self)
        # quirk Q1: read here, before the scenario subclass flips them on `params` -> always False
        self.lockdown = params.lockdown
        self.blocked = params.blocked
        self.school_phase = None
        self.is_school_scenario = False

    # ------------------------------------------------------------------ initialisation (base_model.py:527-634)
    def initialize_epidemic_vectors(self, size):
        h = self._host
        h['epidemic_state'] = np.zeros(shape=size)            # float64 like the reference (Q11)
        h['infection_rate'] = np.zeros(shape=size)
        h['infected_symptomatic_status'] = np.full(shape=size, fill_value=self.SYMPTOM_NONE)
        h['clinical_state'] = np.full(shape=size, fill_value=self.CLINICAL_NOT_HOSPITALIZED)
        for name in ('date_infected', 'date_start_contagious', 'date_start_symptomatic', 'date_recovered'):
            h[name] = _Never(size, datetime.max)
        h['infected_at_district_ids'] = np.full(shape=size, fill_value=-1)
        h['infected_at_location_ids'] = np.full(shape=size, fill_value=-1)
        self.severe_disease_risk = np.ones(shape=size)        # quirk Q2: whatever the scenario hook stored is overwritten

    def initialize_clinical_vectors(self, size):
        for name in ('date_start_hospitalized', 'date_end_hospitalized', 'date_start_critical', 'date_end_critical',
                     'date_died'):
            self._host[name] = _Never(size, datetime.max)

    def initialize_agent_vectors(self, df):
        p = self.params
        size = df.shape[0]
        first_household = max(p.DISTRICT_ID_TO_NAME) + 1       # households are numbered after the districts
        # (factorize hashes each string once and sorts only the distinct names: the dict maps of the reference,
        # base_model.py:577-611, are what dominates its load time at millions of agents)
        hh_codes, hh_names = pd.factorize(df['household_id'].astype(str), sort=True)
        hh_names = np.asarray(hh_names)
        p.HOUSEHOLD_ID_TO_NAME = _IdNameTable(hh_names, first_household)
        p.HOUSEHOLD_NAME_TO_ID = _IdNameTable(hh_names, first_household, inverted=True)
        if self.is_school_scenario:                             # schools after the households
            first_school = _last_id(p.HOUSEHOLD_ID_TO_NAME) + 1
            school_names = sorted(df.loc[df['school_id'] != '', 'school_id'].unique())
            p.SCHOOL_ID_TO_NAME = dict(enumerate(school_names, first_school))
            p.SCHOOL_NAME_TO_ID = {n: i for i, n in p.SCHOOL_ID_TO_NAME.items()}
        p.SEX_ID_TO_NAME = dict(enumerate(sorted(df['sex'].unique())))
        p.SEX_NAME_TO_ID = {n: i for i, n in p.SEX_ID_TO_NAME.items()}
        schools = (p.SCHOOL_ID_TO_NAME,) if self.is_school_scenario else ()
        p.LOCATION_ID_TO_NAME = _IdNameTable(hh_names, first_household, extra=(p.DISTRICT_ID_TO_NAME,) + schools)
        # (inverted: districts, then households, then schools -- a name that is both a district and a household resolves
        # to the household, a school name to the school, exactly like inverting the merged id -> name dict)
        p.LOCATION_NAME_TO_ID = _InvertedLocations(p.DISTRICT_ID_TO_NAME, hh_names, first_household, schools)

        def codes_of(column, names_by_id):
            """ids of a string column whose names are those of `names_by_id` (id -> name): the column is hashed once
            (factorize) and only its few distinct names go through the dictionary."""
            codes, names = pd.factorize(df[column])
            id_of = {n: i for i, n in names_by_id.items()}
            lut = np.array([id_of.get(n, -1) for n in names] + [-1], dtype=np.int64)     # last entry: missing value
            ids = lut[codes]
            if (ids < 0).any():                                  # unknown name: let the reference-style map decide (NaN)
                return np.array(df[column].map(id_of))
            return ids

        def by_status(mapping):
            """Per-agent value of a {status name: value} table (`df['economic_status'].map(mapping)` in the reference,
            base_model.py:617-620), looked up through the status ids."""
            ids = self.economic_status_ids
            if ids.dtype.kind != 'i' or any(name not in mapping for name in p.ECON_STAT_NAME_TO_ID):
                return np.array(df['economic_status'].map(mapping))
            table = np.array([mapping[p.ECON_STAT_ID_TO_NAME[i]] for i in range(len(p.ECON_STAT_ID_TO_NAME))])
            return table[ids]

        self.person_ids = np.arange(size, dtype=int)
        self.district_ids = codes_of('district_id', p.DISTRICT_ID_TO_NAME)
        self.household_ids = hh_codes.astype(np.int64) + first_household
        self.age = np.array(df['age'])
        self.sex = codes_of('sex', p.SEX_ID_TO_NAME)
        self.economic_status_ids = codes_of('economic_status', p.ECON_STAT_ID_TO_NAME)
        # the work / school place is almost always the person's own household or home district: compare names
        # column-wise and send only the rest through the name -> id dictionary
        place = df['economic_activity_location_id'].astype(str).values
        own_household = place == df['household_id'].astype(str).values
        own_district = ~own_household & (place == df['district_id'].astype(str).values)
        loc = np.where(own_household, self.household_ids, np.where(own_district, self.district_ids, -1)).astype(np.int64)
        # a name that is both a district and a household (mining camps) resolves to the household, as in the inverted
        # dictionary of the reference (base_model.py:594-599)
        hh_name_set = pd.Index(hh_names)
        both = {n for n in p.DISTRICT_NAME_TO_ID if n in hh_name_set}
        rest = loc < 0
        if both:
            rest |= own_district & pd.Series(place).isin(both).values
        if rest.any():
            mapped = pd.Series(place[rest]).map(p.LOCATION_NAME_TO_ID).values
            if np.isnan(mapped.astype(np.float64)).any():
                loc = loc.astype(np.float64)
            loc[rest] = mapped
        self.economic_activity_location_ids = loc
        self._status_person_ids = None                          # built on first use (economic_status_ids_person_ids)
        self.economic_status_weekday_movement_probability = by_status(p.ECONOMIC_STATUS_WEEKDAY_MOVEMENT_PROBABILITY)
        self.economic_status_other_day_movement_probability = by_status(p.ECONOMIC_STATUS_OTHER_DAY_MOVEMENT_PROBABILITY)
        self._host['mild_symptom_movement_probability'] = np.ones(shape=size)
        self._host['current_location_ids'] = np.array(self.household_ids)
        self._host['current_district_ids'] = np.array(self.district_ids)
        self.left_district_at = np.full(shape=size, fill_value=datetime.max)
        self._host['return_district_at'] = np.full(shape=size, fill_value=datetime.max)
        self.district_mover = self.DISTRICT_MOVER_TRUE * self._may_travel(self.age >= p.DISTRICT_MOVEMENT_ALLOWED_AGE)

    @property
    def economic_status_ids_person_ids(self):
        """{status id: set of person ids} (base_model.py:613).  Only the reference's explicit contact sampler reads it
        (:84); here it is built when somebody asks, not at load time (python sets of every person id)."""
        if self._status_person_ids is None and self.economic_status_ids is not None:
            self._status_person_ids = {i: set(np.nonzero(self.economic_status_ids == i)[0])
                                       for i in self.params.ECON_STAT_ID_TO_NAME}
        return self._status_person_ids

    def _may_travel(self, age_ok):
        moving = [self.params.ECON_STAT_NAME_TO_ID[s] for s in self.params.DISTRICT_MOVING_ECONOMIC_STATUS
                  if s in self.params.ECON_STAT_NAME_TO_ID]
        return np.isin(self.economic_status_ids, moving) & age_ok

    def scenario_data_preprocessing(self, df):
        pass

    def scenario_data_postprocessing(self, df):
        pass

    # ------------------------------------------------------------------ movement restrictions (base_model.py:642-679)
    def _thin_district_movers(self, ids, allowed_probability):
        """Movers among `ids` keep travelling with `allowed_probability` (scalar or per-agent array)."""
        movers = ids[self.district_mover[ids] == self.DISTRICT_MOVER_TRUE]
        keep = movers[np.random.random(len(movers)) < allowed_probability(movers)]
        self.district_mover[movers] = self.DISTRICT_MOVER_FALSE
        self.district_mover[keep] = self.DISTRICT_MOVER_TRUE

    def _district_table(self, mapping, ids):
        return np.array([mapping[d] for d in self.district_ids[ids]])

    def set_blocked_movers_and_movement_probabilities(self):
        blocked = self.person_ids[np.isin(self.district_ids, self.params.BLOCK_DISTRICTS_IDS)]
        self._thin_district_movers(blocked, lambda movers: self.params.BLOCK_ALLOWED_PROBABILITY)
        self._dirty = True

    def set_lockdown_movers_and_movement_probabilities(self, unrestricted_ids=None):
        locked = self.person_ids[np.isin(self.district_ids, self.params.LOCKDOWN_DISTRICTS_IDS)]
        if unrestricted_ids is not None and len(unrestricted_ids) > 0:
            locked = np.setdiff1d(locked, unrestricted_ids)
        self._thin_district_movers(locked, lambda movers: self._district_table(self.params.LOCKDOWN_ALLOWED_PROBABILITY, movers))
        slowdown = self._district_table(self.params.LOCKDOWN_DECREASED_MOBILITY_RATE, locked)
        self.economic_status_weekday_movement_probability[locked] *= slowdown
        self.economic_status_other_day_movement_probability[locked] *= slowdown
        self._dirty = True

    def setup_selective_movement_restriction_scenarios(self, unrestricted_ids, set_lockdown):
        if set_lockdown:
            self.set_lockdown_movers_and_movement_probabilities(unrestricted_ids=unrestricted_ids)
        # the re-opened group works / studies locally: no inter-district trips, weekend mobility still reduced
        self.district_mover[unrestricted_ids] = self.DISTRICT_MOVER_FALSE
        self.economic_status_other_day_movement_probability[unrestricted_ids] *= self._district_table(
            self.params.LOCKDOWN_DECREASED_MOBILITY_RATE, unrestricted_ids)
        self._dirty = True

    # ------------------------------------------------------------------ schools (base_model.py:681-731, 801-805)
    def set_school_params(self, df):
        if self.params.SCENARIO.endswith('GOVERNMENT_OPEN_SCHOOLS'):
            self.school_phase = np.array(df['phase'])
            self.max_school_phase = max(self.school_phase)
            self.active_school_phases = [1]
            self.school_ids = np.array(df['school_id'].map(self.params.SCHOOL_NAME_TO_ID))

    def _school_population(self, district_ids, active_school_phases):
        return self.person_ids[np.isin(self.district_ids, district_ids) & np.isin(self.school_phase, active_school_phases)]

    def lockdown_schools(self, district_ids, active_school_phases):
        ids = self._school_population(district_ids, active_school_phases)
        self._thin_district_movers(ids, lambda movers: self._district_table(self.params.LOCKDOWN_ALLOWED_PROBABILITY, movers))
        base = self.params.ECONOMIC_STATUS_WEEKDAY_MOVEMENT_PROBABILITY['In School']
        self.economic_status_weekday_movement_probability[ids] = base
        self.economic_status_weekday_movement_probability[ids] *= self._district_table(
            self.params.LOCKDOWN_DECREASED_MOBILITY_RATE, ids)
        self.update_school_economic_activity_location(ids, is_lockdown=True)
        self._dirty = True

    def open_schools(self, district_ids, active_school_phases):
        ids = self._school_population(district_ids, active_school_phases)
        self.economic_status_weekday_movement_probability[ids] = self.params.ECONOMIC_STATUS_WEEKDAY_MOVEMENT_PROBABILITY['In School']
        self.update_school_economic_activity_location(ids, is_lockdown=False)
        self._dirty = True

    def update_school_economic_activity_location(self, school_educ_ids, is_lockdown):
        src = self.district_ids if is_lockdown else self.school_ids
        self.economic_activity_location_ids[school_educ_ids] = src[school_educ_ids]
        self._dirty = True

    def get_safe_and_unsafe_districts(self, quantile_value=0.75):
        """Symptomatic infected-or-contagious share per home district, split at a quantile
        (base_model.py:733-758).  Like the reference it returns the RATES on both sides of the quantile, not district
        ids (quirk Q10).  The per-district counts come from a device reduction (abm_district_symptomatic)."""
        if self._sim is not None:
            self._sync_device_inputs()
            sym, pop = self._sim.district_symptomatic()
            present = pop > 0
            infectious = pd.Series(sym[present] / pop[present], index=np.nonzero(present)[0])
        else:
            sick = np.isin(self.epidemic_state, [self.STATE_INFECTED, self.STATE_CONTAGIOUS]) & (
                self.infected_symptomatic_status == self.SYMPTOM_SYMPTOMATIC)
            D = len(self.params.DISTRICT_IDS)
            pop = np.bincount(self.district_ids, minlength=D)
            present = pop > 0
            infectious = pd.Series(np.bincount(self.district_ids[sick], minlength=D)[present] / pop[present],
                                   index=np.nonzero(present)[0])
        q = infectious.quantile(quantile_value)
        return infectious[infectious <= q].values, infectious[infectious > q].values

    # ------------------------------------------------------------------ load (base_model.py:760-799)
    def load_agents(self, filename, size=None, infect_num=None):
        if filename.endswith('.csv'):
            df = pd.read_csv(filename)
        elif filename.endswith('.pickle'):
            df = pd.read_pickle(filename)
        else:
            raise ValueError('Invalid file type!')
        if size is not None:
            df = df.head(size)
        else:
            size = df.shape[0]

        self.scenario_data_preprocessing(df)
        self.initialize_agent_vectors(df)
        self.scenario_data_postprocessing(df)
        self.initialize_epidemic_vectors(size)
        self.initialize_clinical_vectors(size)

        if infect_num is not None:
            p = self.params
            eligible = (self.age > p.SEED_INFECT_AGE_MIN) & (self.age < p.SEED_INFECT_AGE_MAX)
            home = self._host['current_district_ids']
            eligible &= np.isin(home, p.SEED_INFECT_DISTRICT_IDS)
            chosen = []
            for did, share in p.DISTRICT_ID_INFECTED_PROB.items():     # int(share * n) + 1 per line-list district (Q14)
                cands = self.person_ids[eligible & (home == did)]
                chosen.append(np.random.choice(cands, size=int(share * infect_num) + 1, replace=False))
            self.set_epidemic_status(np.concatenate(chosen))
        del df
        gc.collect()

    # ------------------------------------------------------------------ device store
    def _spec(self):
        p = self.params
        names = p.DISTRICT_IDS
        D = len(names)
        od = np.ascontiguousarray(p.DAILY_DISTRICT_TRANSITION_PROBABILITY.values.reshape(7, D, D), dtype=np.float64)
        stay = p.DISTRICT_WEEKDAY_OD_STAY_COUNT_MATRIX.reindex(pd.MultiIndex.from_product([range(7), names, names]))
        mean = stay['avg_duration'].values.astype(np.float64)
        std = stay['stddev_duration'].values.astype(np.float64)
        missing = np.isnan(mean)
        if missing.any():                                     # pairs without CDR observations: 24 h, average spread
            mean[missing] = 24.0 + 0.001
            std[missing] = np.nanmean(std) if (~missing).any() else 24.0
        status_names = [p.ECON_STAT_ID_TO_NAME[i] for i in range(len(p.ECON_STAT_ID_TO_NAME))]
        sizes = np.array([p.ECONOMIC_STATUS_INTERACTION_SIZE_MAP[s] for s in status_names], dtype=np.float64)
        hw = getattr(self, 'hw_risk', None) if p.SCENARIO.startswith('HANDWASHING_RISK') else None
        n_schools = len(p.SCHOOL_ID_TO_NAME) if self.is_school_scenario else 0
        step = self.step_timedelta
        if step % timedelta(minutes=1) != timedelta(0):
            raise ValueError('step_timedelta must be a whole number of minutes')
        school_district = None
        if n_schools:
            # a school's hand-washing multiplier is that of the district most of its pupils live in
            first_school = _last_id(p.HOUSEHOLD_ID_TO_NAME) + 1
            sid = self.school_ids - first_school
            ok = ~np.isnan(sid.astype(np.float64)) & (sid >= 0)
            school_district = np.zeros(n_schools, dtype=np.int32)
            school_district[sid[ok].astype(np.int64)] = self.district_ids[ok]
        return ModelSpec(
            n_districts=D, R0=p.R0, step_minutes=int(step / timedelta(minutes=1)), start_datetime=self.scheduler.real_time,
            status_names=status_names, interaction_matrix=np.asarray(p.ECONOMIC_STATUS_INTERACTION_MATRIX_VALUES, dtype=np.float64),
            interaction_sizes=sizes, od_cum=od, stay_mean=mean.reshape(7, D, D), stay_std=std.reshape(7, D, D),
            hw_risk=None if hw is None else np.asarray(hw, dtype=np.float64),
            mild_symptom_move_prob=float(p.MILD_SYMPTOM_MOVEMENT_PROBABILITY), n_extra_pools=n_schools,
            extra_pool_district=school_district, symptomatic_rate=float(p.SYMPTOMATIC_RATE),
            critical_fatality_rate=float(p.CRITICAL_FATALITY_RATE))

    def _econ_encoding(self, order):
        """economic_activity_location_ids -> (EKIND_* code, outside-pool row) in canonical order."""
        D = len(self.params.DISTRICT_IDS)
        loc = self.economic_activity_location_ids[order].astype(np.int64)
        hh = self.household_ids[order]
        home = self.district_ids[order]
        kind = np.full(loc.shape[0], -1, dtype=np.int8)
        kind[loc == home] = C.EKIND_HOME_DISTRICT
        kind[loc == hh] = C.EKIND_HOUSEHOLD
        pool = np.zeros(loc.shape[0], dtype=np.int32)
        if self.is_school_scenario:
            first_school = _last_id(self.params.HOUSEHOLD_ID_TO_NAME) + 1
            at_school = loc >= first_school
            kind[at_school] = C.EKIND_POOL
            pool[at_school] = D + (loc[at_school] - first_school)
        if (kind < 0).any():
            bad = order[np.nonzero(kind < 0)[0][:5]]
            raise NotImplementedError(
                f'economic_activity_location_ids of persons {bad.tolist()} is neither their household, their home district '
                'nor a school: not representable in the device layout')
        return kind, (pool if self.is_school_scenario else None)

    def _ensure_device(self):
        if self._sim is not None:
            if self._dirty:
                self.refresh_device()
            return self._sim
        if self.person_ids is None:
            raise RuntimeError('load_agents() has not been called')
        from abm_b200 import multi, tables
        from abm_b200.engine import Simulation
        n = self.person_ids.shape[0]
        if self.age.min() < 0 or self.age.max() > 99:
            raise ValueError('ages must lie in 0..99 (params.ages)')
        n_ranks = multi.world_size()
        store = None
        if self._sim_class is None and n_ranks == 1 and self._restore is None and self.ingest_on_device and _cuda_available():
            # the store is built where it is going to live: the integer columns go up in person order; sort, gather,
            # sub-tile packing and bit packing run on the device (abm_b200/ingest.py); the canonical-order columns come back
            from abm_b200.ingest import DeviceStore
            kind_p, pool_p = self._econ_encoding(np.arange(n))
            store = DeviceStore(self._spec(), district=self.district_ids, household=self.household_ids, age=self.age,
                                status=self.economic_status_ids, econ_kind=kind_p, district_mover=self.district_mover,
                                move_prob_weekday=self.economic_status_weekday_movement_probability,
                                move_prob_other=self.economic_status_other_day_movement_probability, econ_pool=pool_p,
                                device=self.device_index)
            order, pop = store.order, store.pop
        else:
            # canonical order: (home district, household), ties in person order -- one stable sort of a combined key
            span = int(self.household_ids.max()) + 1
            if int(self.district_ids.max()) < (2 ** 62) // span:
                order = np.argsort(self.district_ids.astype(np.int64) * span + self.household_ids, kind='stable')
            else:
                order = np.lexsort((self.household_ids, self.district_ids))
            hh = self.household_ids[order]
            new_hh = np.r_[True, hh[1:] != hh[:-1]]
            dense = np.cumsum(new_hh) - 1
            lowest = int(self.household_ids.min())
            if int(dense[-1]) + 1 != int(np.count_nonzero(np.bincount(self.household_ids - lowest))):
                raise NotImplementedError('a household has members with different home districts: the household-aligned device '
                                          'layout cannot hold it')
            kind, pool = self._econ_encoding(order)
            pop = Population(
                district=self.district_ids[order].astype(np.int32), household=dense.astype(np.int64),
                age=self.age[order].astype(np.int32), status=self.economic_status_ids[order].astype(np.int32), econ_kind=kind,
                district_mover=self.district_mover[order].astype(np.int8),
                move_prob_weekday=self.economic_status_weekday_movement_probability[order].astype(np.float64),
                move_prob_other=self.economic_status_other_day_movement_probability[order].astype(np.float64), econ_pool=pool)
        self._order = order
        self._canon_of_person = np.empty(n, dtype=np.int64)
        self._canon_of_person[order] = np.arange(n)
        restore = self._restore
        if restore is not None:                               # unpickled model: same clock origin, same draw key
            spec, seed = self._spec_obj, self._philox_seed
        else:
            spec = store.spec if store is not None else self._spec()
            seed = self.philox_seed
            if seed is None:
                seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        pop.validate(spec)
        # Under torchrun (one process per GPU, torch.distributed initialised) every rank has built the same host vectors
        # from the same census and numpy seed; each keeps its household-aligned range of the canonical order on its GPU
        # and `ShardedSimulation` makes the N stores look like one (abm_b200/multi.py).
        if n_ranks > 1:
            import os
            device = int(os.environ.get('LOCAL_RANK', self.device_index))
            seed = multi.broadcast_int(seed, device if self._sim_class is None else 'cpu')
            if getattr(self, '_resumable', True) is False:
                raise RuntimeError('this model was pickled from a multi-rank run without its device state: it can be read, not stepped')
        self._philox_seed = int(seed)
        self._spec_obj, self._pop = spec, pop
        if n_ranks > 1:
            self._sim = multi.ShardedSimulation(spec, pop, seed=self._philox_seed, device=device, tables=tables.build_tables(spec),
                                                max_days=self.max_days, upload=restore is None, sim_class=self._sim_class)
        else:
            if getattr(self, '_resumable', True) is False:
                raise RuntimeError('this model was pickled from a multi-rank run without its device state: it can be read, not stepped')
            kw = dict(host_store=store) if store is not None else {}
            self._sim = (self._sim_class or Simulation)(spec, pop, seed=self._philox_seed, device=self.device_index,
                                                        tables=tables.build_tables(spec), max_days=self.max_days,
                                                        upload=restore is None, **kw)
        if restore is not None:
            self._sim.restore(restore)
            self._restore = None
        self._dirty = False
        self._fresh.clear()
        return self._sim

    def refresh_device(self):
        """Push host-side edits of `district_mover`, the movement probabilities and
        `economic_activity_location_ids` to the device (what the reference's scenario hooks mutate in place)."""
        if self._sim is None:
            return
        order, pop = self._order, self._pop
        pop.district_mover = self.district_mover[order].astype(np.int8)
        pop.move_prob_weekday = self.economic_status_weekday_movement_probability[order].astype(np.float64)
        pop.move_prob_other = self.economic_status_other_day_movement_probability[order].astype(np.float64)
        pop.econ_kind, pool = self._econ_encoding(order)
        if pool is not None:
            pop.econ_pool = pool
        self._sim.refresh()
        self._dirty = False
        self._fresh.clear()

    def _sync_device_inputs(self):
        if self._dirty:
            self.refresh_device()

    # ------------------------------------------------------------------ read-back of the dynamic vectors
    def _to_datetimes(self, ticks):
        """int32 minute ticks -> object array of `datetime` (`datetime.max` = never), as the reference stores dates."""
        out = np.full(ticks.shape[0], datetime.max, dtype=object)
        hit = np.nonzero(ticks != T_NEVER)[0]
        if hit.size:
            t0 = np.datetime64(self._spec_obj.start_datetime, 'us')
            out[hit] = (t0 + ticks[hit].astype('timedelta64[m]')).astype(object)      # converted in C, not in a Python loop
        return out

    def _read_back(self):
        """All dynamic vectors, person order, reference dtypes; cached until the next step."""
        sim = self._sim
        canon = sim.state(household_ids=self.household_ids[self._order],
                          econ_loc=self.economic_activity_location_ids[self._order].astype(np.int64))
        cp = self._canon_of_person
        f = self._fresh
        for name in _STATE_VECTORS:
            v = canon[name][cp]
            f[name] = v.astype(np.float64) if name == 'epidemic_state' else v.astype(np.int64)
        for name in _TICK_VECTORS:
            f['_ticks_' + name] = canon[name][cp]
        at = self._spec_obj.age_tables()
        sym = f['infected_symptomatic_status']
        age = self.age.astype(np.int64)
        f['infection_rate'] = np.where(sym == 1, at['r_sym'][age], np.where(sym == 0, at['r_asym'][age], 0.0))
        mild = np.ones(age.shape[0])
        if self.params.MILD_SYMPTOM_MOVEMENT_PROBABILITY < 1:
            last = (sim.step_index - 1) * self._spec_obj.step_minutes          # clock of the last executed sub-step
            onset_seen = canon['date_start_symptomatic'][cp].astype(np.int64) < last
            mild[onset_seen & (f['epidemic_state'] < self.STATE_RECOVERED)] = self.params.MILD_SYMPTOM_MOVEMENT_PROBABILITY
        f['mild_symptom_movement_probability'] = mild

    def _dynamic_vector(self, name):
        if self._sim is None:
            v = self._host[name]
            if isinstance(v, _Never):
                v = self._host[name] = v.make()
            return v
        if name not in self._fresh:
            if 'epidemic_state' not in self._fresh:
                self._read_back()
            if name in _TICK_VECTORS:
                self._fresh[name] = self._to_datetimes(self._fresh['_ticks_' + name])
        return self._fresh[name]

    def event_ticks(self, name):
        """Integer minute ticks since the simulation start of a `date_*` vector (2**31 - 1 = never): the cheap form of
        the object-dtype datetime vectors."""
        if name not in _TICK_VECTORS:
            raise KeyError(name)
        if self._sim is None:
            raise RuntimeError('the device store does not exist yet')
        self._dynamic_vector('epidemic_state')
        return self._fresh['_ticks_' + name]

    # ------------------------------------------------------------------ the step (base_model.py:807-836)
    def step(self):
        scenario = self.params.SCENARIO
        now = self.scheduler.real_time
        if scenario == 'DYNAMIC_PHASE1_GOVERNMENT_OPEN_SCHOOLS':
            if now.day == 1:                                   # every sub-step of the 1st of a month (Q10)
                safe, unsafe = self.get_safe_and_unsafe_districts(quantile_value=0.75)
                self.open_schools(safe, self.active_school_phases)
                self.lockdown_schools(unsafe, self.active_school_phases)
        elif scenario == 'ACCELERATED_GOVERNMENT_OPEN_SCHOOLS':
            if now.day == 1:
                nxt = self.active_school_phases[-1] + 1
                if nxt <= self.max_school_phase:
                    self.active_school_phases.append(nxt)
                    self.open_schools(np.unique(self.district_ids), self.active_school_phases)
        elif scenario == 'PHASE1_GOVERNMENT_OPEN_SCHOOLS':
            if now.year == 2021 and now.day == 1:
                if now.month == 1:
                    self.active_school_phases += [2, 4]
                elif now.month == 5:
                    self.active_school_phases += [3, 5]
                self.open_schools(np.unique(self.district_ids), self.active_school_phases)
        self.scheduler.step()

    def run_days(self, n_days):
        """Extension: enqueue `n_days` whole simulated days in one call (replayed as CUDA graphs), then write the log
        lines.  Equivalent to 6 * n_days calls of step() for scenarios without clock hooks."""
        if self.params.SCENARIO.endswith('GOVERNMENT_OPEN_SCHOOLS'):
            for _ in range(int(n_days) * int(timedelta(days=1) / self.step_timedelta)):
                self.step()
            return
        sim = self._ensure_device()
        per_day = int(timedelta(days=1) / self.step_timedelta)
        hour0 = self.scheduler.real_time
        sim.step(per_day * int(n_days))
        for i in range(per_day * int(n_days)):
            if (hour0 + i * self.step_timedelta).hour == 8:
                self._log_rows_due += 1
        self.scheduler.steps += per_day * int(n_days)
        self.scheduler.time += per_day * int(n_days)
        self.scheduler._real_time = hour0 + per_day * int(n_days) * self.step_timedelta
        self._fresh.clear()
        if self.model_log_file is not None:
            self.log_model_output()

    def active_infections(self):
        """Number of exposed + contagious agents -- run_scenario's stop test (scenario_models.py:494) without
        reading the state vector back."""
        if self._sim is None:
            s = self._host['epidemic_state']
            return int(((s >= self.STATE_INFECTED) & (s < self.STATE_RECOVERED)).sum())
        return self._sim.active_infections()      # running device counters: no flush, the fused launch sequence stays

    # ------------------------------------------------------------------ infection (base_model.py:838-1055)
    def set_epidemic_status(self, neighbors_to_infect):
        """Infect the given persons now: the whole disease course (symptomatic?, incubation, hospital / critical /
        fatal outcome, every dwell time) is sampled on the device by `infect_agent`."""
        ids = np.asarray(neighbors_to_infect, dtype=np.int64)
        if ids.size == 0:
            return
        sim = self._ensure_device()
        sim.seed_infections(np.sort(self._canon_of_person[ids]))
        self._fresh.clear()

    def get_location_person_ids(self, person_ids=None):
        """{location id: person ids there} (base_model.py:1057-1080).  The step itself never groups agents (it
        accumulates district x status tables instead); this is for analysis code."""
        loc = self.current_location_ids
        ids = np.arange(loc.shape[0]) if person_ids is None else np.asarray(person_ids)
        order = np.argsort(loc[ids], kind='stable')
        sorted_loc = loc[ids][order]
        cuts = np.nonzero(np.r_[True, sorted_loc[1:] != sorted_loc[:-1], True])[0]
        return {int(sorted_loc[cuts[i]]): ids[order[cuts[i]:cuts[i + 1]]] for i in range(cuts.shape[0] - 1)}

    # ------------------------------------------------------------------ daily log (base_model.py:1082-1106)
    def log_model_output(self):
        """Write the day-log rows the device has produced since the last call, one JSON line per simulated day at
        08:00, in the reference's format and key order."""
        if self._sim is None or self.model_log_file is None:
            self._log_rows_written = self._log_rows_due
            return
        rows = self._sim.log_dicts()             # multi-rank: summed over ranks (a collective: every rank calls it)
        if getattr(self._sim, 'rank', 0) == 0:   # ... and rank 0 writes the file
            for r in rows[self._log_rows_written:self._log_rows_due]:
                data = dict(date=r['date'])
                data.update({k: r[k] for k in _COUNTER_KEYS})
                log_to_file(self.model_log_file, json.dumps(data), as_log=True, verbose=False)
        self._log_rows_written = max(self._log_rows_written, min(self._log_rows_due, len(rows)))

    # ------------------------------------------------------------------ pickling (scenario_models.py:497-500)
    def __getstate__(self):
        d = dict(self.__dict__)
        if self._sim is not None:
            if self.model_log_file is not None:
                self.log_model_output()
            try:
                d['_restore'] = self._sim.checkpoint()
            except NotImplementedError:                          # multi-rank: the gathered host vectors only
                d['_restore'] = None
                d['_resumable'] = False
            d['_log_rows_written'] = d['_log_rows_due'] = 0      # the day log restarts with the restored handle
            d['_host'] = {name: self._dynamic_vector(name) for name in _DYNAMIC}
        d['_sim'] = None
        d['_fresh'] = {}
        d['_status_person_ids'] = None                           # derived; rebuilt on demand
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.__dict__.setdefault('_status_person_ids', None)

    def close(self):
        if self._sim is not None:
            if self.model_log_file is not None:
                self.log_model_output()
            self._sim.close()
            self._sim = None


for _name in _DYNAMIC:
    setattr(Country, _name, _dynamic_property(_name))
