"""Plain-numpy description of one model instance: the inputs the agent step needs.

`ModelSpec` is what `ParamsConfig` (reference: src/covid19_abm/params.py:199-276) boils down to for
the hot path; `Population` is what `Country.initialize_agent_vectors` (base_model.py:572-634) builds
from the census DataFrame, in canonical order (sorted by home district, then household).
"""
from dataclasses import dataclass, field
from datetime import datetime
from typing import Optional

import numpy as np

from . import constants as C


def band_lookup(bands, ages):
    """Value of the first band whose inclusive upper bound is >= age (params.py:223-227 idiom)."""
    uppers = np.array([b[0] for b in bands], dtype=np.float64)
    values = np.array([b[1] for b in bands], dtype=np.float64)
    idx = np.searchsorted(uppers, np.asarray(ages, dtype=np.float64), side='left')
    return values[idx]


@dataclass
class ModelSpec:
    n_districts: int
    R0: float = 1.9
    step_minutes: int = 240
    start_datetime: datetime = datetime(2020, 9, 1, 0)
    status_names: list = field(default_factory=lambda: list(C.ECONOMIC_STATUS_NAMES))
    interaction_matrix: np.ndarray = None      # [A, A] row-stochastic, infector row -> contact column
    interaction_sizes: np.ndarray = None       # [A] contacts per sub-step
    od_cum: np.ndarray = None                  # [7, D, D] cumulative destination probabilities
    stay_mean: np.ndarray = None               # [7, D, D] hours (already + 0.001, params.py:250)
    stay_std: np.ndarray = None                # [7, D, D]
    hw_risk: Optional[np.ndarray] = None       # [D] or None (HANDWASHING_RISK* scenarios only)
    severe_risk: Optional[np.ndarray] = None   # [D]; ones under reference parity (quirk Q2)
    mild_symptom_move_prob: float = 1.0        # params.py:192
    n_extra_pools: int = 0                     # school / other outside locations beyond the districts
    extra_pool_district: Optional[np.ndarray] = None   # [n_extra_pools] district owning each extra pool
    symptomatic_rate: float = C.SYMPTOMATIC_RATE
    critical_fatality_rate: float = C.CRITICAL_FATALITY_RATE

    def __post_init__(self):
        if self.interaction_matrix is None:
            self.interaction_matrix = np.array(C.INTERACTION_MATRIX, dtype=np.float64)
        if self.interaction_sizes is None:
            self.interaction_sizes = np.array(C.INTERACTION_SIZES, dtype=np.float64)
        self.interaction_matrix = np.asarray(self.interaction_matrix, dtype=np.float64)
        self.interaction_sizes = np.asarray(self.interaction_sizes, dtype=np.float64)
        if self.severe_risk is None:
            self.severe_risk = np.ones(self.n_districts)
        if self.extra_pool_district is None:
            self.extra_pool_district = np.zeros(self.n_extra_pools, dtype=np.int32)

    # ---- derived, following params.py:220-242
    @property
    def n_status(self):
        return int(self.interaction_matrix.shape[0])

    @property
    def n_pools(self):
        return self.n_districts + self.n_extra_pools

    def age_tables(self):
        ages = np.arange(100)
        p_hosp = band_lookup(C.HOSPITALIZATION_PCT_BY_AGE, ages) / 100 / self.symptomatic_rate
        p_crit = band_lookup(C.CRITICAL_PCT_BY_AGE, ages) / 100
        p_hfat = band_lookup(C.HOSPITAL_DEATH_PCT_BY_AGE, ages) / 100
        contact = band_lookup(C.PER_CAPITA_CONTACT_RATE_BY_AGE, ages)
        steps_per_day = (24 * 60) / self.step_minutes
        r_asym = self.R0 / ((C.ASYM_CONTAGIOUS_MEAN_DAYS * steps_per_day) * contact)
        r_sym = self.R0 / ((C.SYM_CONTAGIOUS_MEAN_DAYS * steps_per_day) * contact)
        return dict(p_hosp=p_hosp, p_crit=p_crit, p_hfat=p_hfat, r_asym=r_asym, r_sym=r_sym)


@dataclass
class Population:
    """Agents in canonical order.  `household` is a dense index 0..H-1, non-decreasing."""
    district: np.ndarray           # int32 [N] home district id
    household: np.ndarray          # int64 [N]
    age: np.ndarray                # int32 [N]
    status: np.ndarray             # int32 [N] economic status id
    econ_kind: np.ndarray          # int8  [N] EKIND_*
    district_mover: np.ndarray     # int8  [N]
    move_prob_weekday: np.ndarray  # float64 [N]
    move_prob_other: np.ndarray    # float64 [N]
    econ_pool: Optional[np.ndarray] = None   # int32 [N] pool row for EKIND_POOL agents
    extras: dict = field(default_factory=dict)   # scenario flags (school_goers, manufacturing_workers, ...)

    @property
    def n(self):
        return int(self.district.shape[0])

    @property
    def n_households(self):
        return int(self.household[-1]) + 1 if self.n else 0

    def validate(self, spec: ModelSpec):
        n = self.n
        assert np.all(np.diff(self.household) >= 0), 'agents must be grouped by household'
        assert np.all(np.diff(self.district) >= 0), 'agents must be sorted by home district'
        assert self.age.min() >= 0 and self.age.max() <= 99, 'age tables cover 0..99 (params.py:124)'
        assert self.status.min() >= 0 and self.status.max() < spec.n_status
        assert self.district.min() >= 0 and self.district.max() < spec.n_districts
        for a in (self.household, self.age, self.status, self.econ_kind, self.district_mover,
                  self.move_prob_weekday, self.move_prob_other):
            assert a.shape[0] == n
