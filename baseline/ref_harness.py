"""Run the REAL, unmodified reference (worldbank/covid19-agent-based-model, src/covid19_abm).

Where the reference comes from: `baseline/_ref/src` when it exists (the copy `baseline/make_ref.py` makes from
/root/reference; git-ignored, but it travels to the GPU box with the working tree), else /root/reference/src (the build
container).  Users: the fixture generators under tests/golden/ (build container only) and `bench.py`'s reference arm /
`cpu_baseline` leg (GPU box).  Nothing in the product path imports it.  It
  * puts stub `mesa` / `pylab` packages (ref_shims/) and the unmodified reference sources on sys.path,
  * patches pandas.read_excel (no openpyxl here) to return the shipped interaction matrix values,
  * synthesises the private data files the reference loads (census pickle, OD csv, stay pickle,
    line list, intra-district rates) from our seedable synthetic generator, into a temp data dir that
    the reference's `get_data_dir` is redirected to,
  * builds a model exactly like scenario_models.run_scenario (scenario_models.py:471-487).
"""
import os
import pickle
import shutil
import sys
import tempfile
import warnings
from datetime import datetime, timedelta

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
_REF_LOCAL = os.path.join(HERE, '_ref')
if os.path.isdir(os.path.join(_REF_LOCAL, 'src', 'covid19_abm')):
    REF_SRC = os.path.join(_REF_LOCAL, 'src')
    RISK_CSV = os.path.join(_REF_LOCAL, 'data', 'hw_and_severe_disease_risk.csv')
else:
    REF_SRC = '/root/reference/src'
    RISK_CSV = '/root/reference/data/preprocessed/risk/hw_and_severe_disease_risk.csv'


def available():
    """True when the reference sources can be imported here."""
    return os.path.isdir(os.path.join(REF_SRC, 'covid19_abm')) and os.path.exists(RISK_CSV)


if os.path.join(REPO, 'covid19-agent-based-model_b200') not in sys.path:
    sys.path.insert(0, os.path.join(REPO, 'covid19-agent-based-model_b200'))

from abm_b200 import constants as C          # noqa: E402
from abm_b200 import synth                   # noqa: E402
from abm_b200.spec import ModelSpec, Population   # noqa: E402

T_NEVER = int(C.T_NEVER)


def _install_import_shims():
    for p in (os.path.join(HERE, 'ref_shims'), REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings('ignore', category=DeprecationWarning)
    warnings.filterwarnings('ignore', category=FutureWarning)


def _fake_read_excel(status_names, matrix, sizes):
    def read_excel(fname, sheet_name=None, index_col=None):
        if sheet_name == 'interaction_matrix':
            return pd.DataFrame(np.asarray(matrix, dtype=float), index=status_names, columns=status_names)
        if sheet_name == 'interactions':
            return pd.DataFrame({'interactions': [int(v) for v in sizes]}, index=status_names)
        raise ValueError(sheet_name)
    return read_excel


def population_to_dataframe(pop: Population, n_districts, seed=0):
    """Census-style DataFrame.  Zero-padded household names keep the
    reference's string sort (base_model.py:577-581) equal to our numeric household order."""
    names = np.array(synth.district_names(n_districts))
    rng = np.random.default_rng(seed + 5)
    hh_names = np.char.add('h_', np.char.zfill(pop.household.astype(str), 9))
    d_names = names[pop.district]
    econ_loc = np.where(pop.econ_kind == C.EKIND_HOUSEHOLD, hh_names, d_names)
    status_names = np.array(C.ECONOMIC_STATUS_NAMES)
    df = pd.DataFrame({
        'person_id': np.arange(pop.n),
        'age': pop.age.astype(np.int64),
        'sex': rng.choice(['f', 'm'], pop.n),
        'household_id': hh_names,
        'district_id': d_names,
        'economic_status': status_names[pop.status],
        'economic_activity_location_id': econ_loc,
    })
    for k, v in pop.extras.items():
        df[k] = v
    return df


class RefWorld:
    """A temp data directory + the imported reference modules."""

    def __init__(self, spec: ModelSpec, pop: Population, seed_cases: dict, lockdown_rates=None,
                 sample_tag=100):
        _install_import_shims()
        self.spec, self.pop = spec, pop
        self.tmp = tempfile.mkdtemp(prefix='abmref_')
        D = spec.n_districts
        names = synth.district_names(D)
        for sub in ('raw', 'preprocessed/census', 'preprocessed/mobility', 'preprocessed/line_list',
                    'preprocessed/risk', 'logs'):
            os.makedirs(os.path.join(self.tmp, sub), exist_ok=True)
        # OD cumulative csv: index (weekday, home_region), columns d_* (appendix B.2)
        idx = pd.MultiIndex.from_product([range(7), names], names=['weekday', 'home_region'])
        od = pd.DataFrame(spec.od_cum.reshape(7 * D, D), index=idx, columns=names)
        self.od_file = os.path.join(self.tmp, 'preprocessed/mobility/od.csv')
        od.to_csv(self.od_file, float_format='%.17g')
        # stay pickle: MultiIndex (weekday, home, dest); the reference adds 0.001 itself (params.py:250)
        sidx = pd.MultiIndex.from_product([range(7), names, names], names=['weekday', 'home_region', 'region'])
        stay = pd.DataFrame({'avg_duration': self.raw_stay_mean.reshape(-1),
                             'stddev_duration': self.raw_stay_std.reshape(-1)}, index=sidx)
        self.stay_file = os.path.join(self.tmp, 'preprocessed/mobility/stay.pickle')
        stay.to_pickle(self.stay_file)
        with open(os.path.join(self.tmp, 'preprocessed/line_list/latest_line_list.pickle'), 'wb') as fl:
            pickle.dump({names[d]: c for d, c in seed_cases.items()}, fl)
        rates = synth.lockdown_rates(D) if lockdown_rates is None else lockdown_rates
        pd.DataFrame({'pctdif_distance': rates}, index=names).to_csv(
            os.path.join(self.tmp, 'preprocessed/mobility/intra_district_decreased_mobility_rates.csv'))
        shutil.copy(RISK_CSV, os.path.join(self.tmp, 'preprocessed/risk/hw_and_severe_disease_risk.csv'))
        self.census_file = os.path.join(
            self.tmp, f'preprocessed/census/zimbabwe_expanded_census_consolidated_{sample_tag}pct.pickle')
        population_to_dataframe(pop, D).to_pickle(self.census_file)
        self.sample_tag = sample_tag

        pd.read_excel = _fake_read_excel(spec.status_names, spec.interaction_matrix, spec.interaction_sizes)
        import covid19_abm.params as ref_params
        import covid19_abm.base_model as ref_base
        import covid19_abm.scenario_models as ref_scen
        tmp = self.tmp
        ref_params.get_data_dir = lambda *parts: os.path.join(tmp, *parts)
        self.ref_params, self.ref_base, self.ref_scen = ref_params, ref_base, ref_scen

    @property
    def raw_stay_mean(self):
        return self.spec.stay_mean - 0.001

    @property
    def raw_stay_std(self):
        return self.spec.stay_std - 0.001

    def make_params(self, seed_infected):
        spec = self.spec
        p = self.ref_params.ParamsConfig(
            district='new', data_sample_size=self.sample_tag, R0=spec.R0,
            normal_interaction_matrix_file='normal.xlsx', lockdown_interaction_matrix_file='lockdown.xlsx',
            stay_duration_file=self.stay_file, transition_probability_file=self.od_file,
            timestep=timedelta(minutes=spec.step_minutes))
        p.set_new_district_seed(seed_infected=seed_infected)
        return p

    def make_model(self, scenario_name='UnmitigatedScenario', seed_infected=10, log_file=None, load=True,
                   infect=True):
        params = self.make_params(seed_infected)
        cls = getattr(self.ref_scen, scenario_name)
        model = cls(params, model_log_file=log_file, individual_log_file=None)
        params.SIMULATION_START_DATE = self.spec.start_datetime
        model.scheduler.real_time = params.SIMULATION_START_DATE
        if load:
            model.load_agents(params.data_file_name, size=None,
                              infect_num=params.SEED_INFECT_NUM if infect else None)
        return model

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


def spec_from_reference_params(params, step_minutes, start_datetime, n_extra_pools=0):
    """Pull the hot-path inputs back out of a real reference ParamsConfig (what the drop-in does)."""
    D = len(params.DISTRICT_IDS)
    od = params.DAILY_DISTRICT_TRANSITION_PROBABILITY.values.reshape(7, D, D).astype(np.float64)
    stay = params.DISTRICT_WEEKDAY_OD_STAY_COUNT_MATRIX
    names = params.DISTRICT_IDS
    full = pd.MultiIndex.from_product([range(7), names, names])
    stay = stay.reindex(full)
    hw = None
    if params.SCENARIO.startswith('HANDWASHING_RISK'):
        hw = np.array([params.DISTRICT_HW_RISK[n] for n in names])
    return ModelSpec(
        n_districts=D, R0=params.R0, step_minutes=step_minutes, start_datetime=start_datetime,
        status_names=[params.ECON_STAT_ID_TO_NAME[i] for i in range(len(params.ECON_STAT_ID_TO_NAME))],
        interaction_matrix=np.asarray(params.ECONOMIC_STATUS_INTERACTION_MATRIX_VALUES, dtype=np.float64),
        interaction_sizes=np.array([params.ECONOMIC_STATUS_INTERACTION_SIZE_MAP[params.ECON_STAT_ID_TO_NAME[i]]
                                    for i in range(len(params.ECON_STAT_ID_TO_NAME))], dtype=np.float64),
        od_cum=od, stay_mean=stay['avg_duration'].values.reshape(7, D, D).astype(np.float64),
        stay_std=stay['stddev_duration'].values.reshape(7, D, D).astype(np.float64),
        hw_risk=hw, mild_symptom_move_prob=float(params.MILD_SYMPTOM_MOVEMENT_PROBABILITY),
        n_extra_pools=n_extra_pools)


def to_ticks(dates, start):
    """object array of datetime -> int32 minute ticks (datetime.max -> T_NEVER); asserts whole minutes."""
    out = np.empty(len(dates), dtype=np.int64)
    for i, d in enumerate(dates):
        if d == datetime.max:
            out[i] = T_NEVER
        else:
            us = (d - start) // timedelta(microseconds=1)
            assert us % 60_000_000 == 0, (d, start)
            out[i] = us // 60_000_000
    return out.astype(np.int32)


def reference_state(model, start):
    """Reference vectors converted to the oracle's representation."""
    g = lambda name: np.asarray(getattr(model, name))
    st = dict(
        epidemic_state=g('epidemic_state').astype(np.int8),
        clinical_state=g('clinical_state').astype(np.int8),
        infected_symptomatic_status=g('infected_symptomatic_status').astype(np.int8),
        current_location_ids=g('current_location_ids').astype(np.int64),
        current_district_ids=g('current_district_ids').astype(np.int64),
        infected_at_district_ids=g('infected_at_district_ids').astype(np.int64),
        infected_at_location_ids=g('infected_at_location_ids').astype(np.int64),
    )
    for name in ('return_district_at', 'date_infected', 'date_start_contagious', 'date_start_symptomatic',
                 'date_recovered', 'date_start_hospitalized', 'date_end_hospitalized', 'date_start_critical',
                 'date_end_critical', 'date_died'):
        st[name] = to_ticks(g(name), start)
    return st
