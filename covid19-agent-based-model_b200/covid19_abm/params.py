"""`ParamsConfig` -- the model's configuration object, API-compatible with the reference's
`covid19_abm.params.ParamsConfig` (/root/reference/src/covid19_abm/params.py:10-607), and `log_to_file`
(params.py:610-621).

Same constructor keywords, the same public attributes (constants, age -> probability dictionaries, infection
rate tables, OD / stay / interaction tables, id maps) and the same `scenario_*` / `set_*` mutators, so scenario
code written against the reference runs unchanged.  What differs is underneath: nothing here is consulted
during a step.  `Country` boils this object down to an `abm_b200.spec.ModelSpec` once, and
`abm_b200.tables.build_tables` turns that into the integer tables the CUDA kernels read.

The scenario mutators are generated from the declarative tables at the bottom of this file; each entry cites the
reference method it reproduces.
"""
import os
import pickle
import zipfile
from datetime import datetime, timedelta
from xml.etree import ElementTree

import numpy as np
import pandas as pd

from covid19_abm.dir_manager import get_data_dir

_STATUSES = ('Not working, inactive, not in universe', 'In School', 'Homemakers/Housework', 'Office workers',
             'Teachers', 'Service Workers', 'Agriculture Workers', 'Indusrtry Workers', 'In the army',
             'Disabled and not working')


def _bands(pairs, scale=1.0):
    """{inclusive upper age bound: value} as a Series; the open-ended last band is keyed by inf."""
    return pd.Series({k: v for k, v in pairs}) * scale


def _band_value(series, age):
    """Value of the first band whose upper bound is >= age (params.py:223-227, 237, 241)."""
    return series.iloc[int(np.searchsorted(series.index.values, age, side='left'))]


# ------------------------------------------------------------------------------------------------ workbook reader
def _xlsx_sheets(path):
    """Minimal .xlsx reader (no openpyxl needed): {sheet name: list of rows of cell values}."""
    ns = {'m': 'http://schemas.openxmlformats.org/spreadsheetml/2006/main',
          'r': 'http://schemas.openxmlformats.org/officeDocument/2006/relationships'}
    with zipfile.ZipFile(path) as z:
        shared = []
        if 'xl/sharedStrings.xml' in z.namelist():
            for si in ElementTree.fromstring(z.read('xl/sharedStrings.xml')).findall('m:si', ns):
                shared.append(''.join(t.text or '' for t in si.iter('{%s}t' % ns['m'])))
        rels = {r.get('Id'): r.get('Target') for r in ElementTree.fromstring(z.read('xl/_rels/workbook.xml.rels'))}
        out = {}
        wb = ElementTree.fromstring(z.read('xl/workbook.xml'))
        for sh in wb.find('m:sheets', ns):
            target = rels[sh.get('{%s}id' % ns['r'])]
            target = target.lstrip('/') if target.startswith('/') else 'xl/' + target
            rows = []
            for row in ElementTree.fromstring(z.read(target)).iter('{%s}row' % ns['m']):
                cells = {}
                col = 0
                for c in row.findall('m:c', ns):
                    ref = c.get('r')
                    if ref is None:
                        col += 1                      # cells without a reference follow each other
                    else:
                        col = 0
                        for ch in ref:
                            if ch.isalpha():
                                col = col * 26 + (ord(ch.upper()) - 64)
                    v, t = c.find('m:v', ns), c.get('t')
                    if t == 'inlineStr':
                        val = ''.join(x.text or '' for x in c.iter('{%s}t' % ns['m']))
                    elif v is None:
                        val = None
                    elif t == 's':
                        val = shared[int(v.text)]
                    elif t in ('str', 'b'):
                        val = v.text
                    else:
                        val = float(v.text)
                    cells[col - 1] = val
                if cells:
                    rows.append([cells.get(i) for i in range(max(cells) + 1)])
            out[sh.get('name')] = rows
    return out


def read_interaction_workbook(path):
    """(matrix DataFrame, sizes Series) from the `interaction_matrix` / `interactions` sheets (params.py:282-288)."""
    try:
        matrix = pd.read_excel(path, sheet_name='interaction_matrix', index_col=0)
        sizes = pd.read_excel(path, sheet_name='interactions', index_col=0)['interactions']
        return matrix, sizes
    except ImportError:                 # no Excel engine installed: parse the zip ourselves
        pass
    sheets = _xlsx_sheets(path)
    m = sheets['interaction_matrix']
    cols = [c for c in m[0][1:] if c is not None]
    body = [r for r in m[1:] if r and r[0] is not None]
    matrix = pd.DataFrame([[float(v or 0.0) for v in (r[1:1 + len(cols)] + [0.0] * len(cols))[:len(cols)]] for r in body],
                          index=[r[0] for r in body], columns=cols)
    s = sheets['interactions']
    head = s[0]
    j = head.index('interactions')
    body = [r for r in s[1:] if r and r[0] is not None]
    sizes = pd.Series([r[j] for r in body], index=[r[0] for r in body], name='interactions')
    if all(float(v).is_integer() for v in sizes):
        sizes = sizes.astype(int)
    return matrix, sizes


class ParamsConfig:
    # probability of leaving the household on a weekday / any other day, by economic status   (params.py:13-38)
    ECONOMIC_STATUS_WEEKDAY_MOVEMENT_PROBABILITY = dict(zip(_STATUSES, (0.8, 1.0, 0.8, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0)))
    ECONOMIC_STATUS_OTHER_DAY_MOVEMENT_PROBABILITY = dict(zip(_STATUSES, (0.8,) * 9 + (0.0,)))

    # age-band tables (params.py:41-122); bounds are inclusive upper ages
    susceptibility_by_age = _bands([(9, 0.4), (19, 0.38), (29, 0.79), (39, 0.86), (49, 0.8), (59, 0.82), (69, 0.88),
                                    (np.inf, 0.74)])
    hospitalization_rates_by_age = _bands(
        [(14, 0.1), (19, 0.2), (24, 0.5), (29, 1.0), (34, 1.6), (39, 2.3), (44, 2.9), (49, 3.9), (54, 5.8), (59, 7.2),
         (64, 10.2), (69, 11.7), (74, 14.6), (79, 17.7), (np.inf, 18.0)]) / 100
    critical_rates_by_age = _bands(
        [(34, 5.0), (39, 5.3), (44, 6.0), (49, 7.5), (54, 10.4), (59, 14.9), (64, 22.4), (69, 30.7), (74, 38.6),
         (79, 46.1), (np.inf, 70.9)]) / 100
    hospitalized_death_rates_by_age = _bands(
        [(39, 1.3), (44, 1.5), (49, 1.9), (54, 2.7), (59, 4.2), (64, 6.9), (69, 10.5), (74, 14.9), (79, 20.3),
         (np.inf, 58.0)]) / 100
    per_capita_contact_rates = _bands(
        [(5, 7.4709), (10, 10.4079), (15, 12.4579), (20, 11.1542), (25, 10.8716), (30, 9.7019), (35, 11.7569),
         (40, 12.5434), (45, 13.2094), (50, 10.7766), (55, 10.3643), (60, 10.9087), (65, 11.1501), (70, 11.6564),
         (75, 11.8107), (np.inf, 11.5379)])

    ages = list(range(100))

    # clock (params.py:127-136): a step must land on the start / end-of-day hours or nobody moves
    step_timedelta = timedelta(hours=4)
    start_datetime = datetime(2020, 4, 20)
    WEEKDAY_START_DAY_HOUR = 8
    WEEKDAY_END_DAY_HOUR = 16
    OTHER_DAY_START_DAY_HOUR = 8
    OTHER_DAY_END_DAY_HOUR = 16

    INTERACTION_SIZE_MAP = {}
    INTERACTION_SIZE_MAX = 10
    INTERACTION_SIZE_MEAN = 10

    R0 = 3.0

    # contagious periods, days (params.py:145-161)
    ASYMPTOMATIC_CONTAGIOUS_PERIOD_MEAN = 6.5
    ASYMPTOMATIC_CONTAGIOUS_PERIOD_SHAPE = 8.45
    ASYMPTOMATIC_CONTAGIOUS_PERIOD_SCALE = 0.7692307692307693
    SYMPTOMATIC_CONTAGIOUS_PERIOD_MEAN = 7
    SYMPTOMATIC_CONTAGIOUS_PERIOD_SHAPE = 9.8
    SYMPTOMATIC_CONTAGIOUS_PERIOD_SCALE = 0.7142857142857143
    ASYMPTOMATIC_TO_CONTAGIOUS_PERIOD_MEAN = 4.6
    ASYMPTOMATIC_TO_CONTAGIOUS_PERIOD = timedelta(days=ASYMPTOMATIC_TO_CONTAGIOUS_PERIOD_MEAN)
    SYMPTOM_TO_RECOVERY_PERIOD = timedelta(days=SYMPTOMATIC_CONTAGIOUS_PERIOD_MEAN)
    ASYMPTOMATIC_TO_RECOVERY_PERIOD = timedelta(days=ASYMPTOMATIC_CONTAGIOUS_PERIOD_MEAN)

    # clinical course, fixed offsets (params.py:165-177)
    SYMPTOM_TO_HOSPITALIZATION_PERIOD = timedelta(days=5)
    HOSPITALIZATION_PERIOD = timedelta(days=8)
    CRITICAL_PERIOD = timedelta(days=8)
    SYMPTOMATIC_TO_DEATH_PERIOD = timedelta(days=21)

    SYMPTOMATIC_RATE = 0.6
    CRITICAL_FATALITY_RATE = 0.5
    INCUBATION_PERIOD_SHAPE = 6.474197530864197
    INCUBATION_PERIOD_SCALE = 0.7074235807860262

    MILD_SYMPTOM_MOVEMENT_PROBABILITY = 1.0
    DISTRICT_HOSPITALS = {}
    DISTRICT_MOVEMENT_ALLOWED_AGE = 18

    def __init__(
        self, district='old', data_sample_size=5, R0=None,
        normal_interaction_matrix_file='final_close_interaction_matrix_normal.xlsx',
        lockdown_interaction_matrix_file='final_close_interaction_matrix_lockdown.xlsx',
        stay_duration_file='weekday_mobility_duration_count_df.pickle',
        transition_probability_file='daily_region_transition_probability.csv',
        intra_district_decreased_mobility_rates_file='intra_district_decreased_mobility_rates.csv',
        timestep=None,
    ):
        if R0 is not None:
            self.R0 = R0
        if timestep is not None:
            self.step_timedelta = timestep
        self.data_sample_size = data_sample_size
        self.district_type = district
        self.SCENARIO = 'UNMITIGATED'

        # the published rates lump mild and asymptomatic cases together; asymptomatic ones are modelled apart here
        self.hospitalization_rates_by_age = type(self).hospitalization_rates_by_age / self.SYMPTOMATIC_RATE   # params.py:220
        self.MEAN_HOSPITALIZATION_RATE = self.hospitalization_rates_by_age.mean()
        per_age = lambda series: {age: _band_value(series, age) for age in self.ages}
        self.AGE_HOSPITALIZATION_PROBABILITY = per_age(self.hospitalization_rates_by_age)
        self.AGE_CRITICAL_CARE_PROBABILITY = per_age(self.critical_rates_by_age)
        self.AGE_HOSPITALIZATION_FATALITY_PROBABILITY = per_age(self.hospitalized_death_rates_by_age)
        self.AGE_SUSCEPTIBILITY_PROBABILITY = per_age(self.susceptibility_by_age)

        # infection rate per contact and per step (params.py:231-242)
        asym_steps = timedelta(days=self.ASYMPTOMATIC_CONTAGIOUS_PERIOD_MEAN) / self.step_timedelta
        sym_steps = timedelta(days=self.SYMPTOMATIC_CONTAGIOUS_PERIOD_MEAN) / self.step_timedelta
        self.ASYMPTOMATIC_INFECTION_RATE = self.R0 / (asym_steps * self.INTERACTION_SIZE_MEAN)
        self.SYMPTOMATIC_INFECTION_RATE = self.R0 / (sym_steps * self.INTERACTION_SIZE_MEAN)
        contact = [_band_value(self.per_capita_contact_rates, age) for age in self.ages]
        self.AGE_ASYMPTOMATIC_INFECTION_RATE = pd.Series({a: self.R0 / (asym_steps * c) for a, c in zip(self.ages, contact)})
        self.AGE_ASYMPTOMATIC_INFECTION_RATE_VALUES = self.AGE_ASYMPTOMATIC_INFECTION_RATE[sorted(self.ages)].values
        self.AGE_SYMPTOMATIC_INFECTION_RATE = pd.Series({a: self.R0 / (sym_steps * c) for a, c in zip(self.ages, contact)})
        self.AGE_SYMPTOMATIC_INFECTION_RATE_VALUES = self.AGE_SYMPTOMATIC_INFECTION_RATE[sorted(self.ages)].values

        # who may travel between districts: every status that leaves the house, except pupils and teachers (params.py:244-246)
        self.DISTRICT_MOVING_ECONOMIC_STATUS = {s for s, p in self.ECONOMIC_STATUS_WEEKDAY_MOVEMENT_PROBABILITY.items() if p > 0}
        self.DISTRICT_MOVING_ECONOMIC_STATUS -= {'In School', 'Teachers'}

        # length of stay (hours) per (weekday, home, destination); + 0.001 keeps the Gamma parameters positive
        stay = pd.read_pickle(stay_duration_file)
        cols = ['avg_duration', 'stddev_duration']
        stay[cols] = stay[cols] + 0.001
        self.DISTRICT_WEEKDAY_OD_STAY_COUNT_MATRIX = stay

        # cumulative destination probabilities: rows (weekday, home), columns = districts sorted as strings
        od = pd.read_csv(transition_probability_file, index_col=[0, 1])
        od = od.loc[sorted(od.index)]
        self.DISTRICT_IDS = sorted(od.columns)
        self.DISTRICT_ID_TO_NAME = dict(enumerate(self.DISTRICT_IDS))
        self.DISTRICT_NAME_TO_ID = {name: i for i, name in self.DISTRICT_ID_TO_NAME.items()}
        self.DAILY_DISTRICT_TRANSITION_PROBABILITY = od[self.DISTRICT_IDS]

        for name in self.DISTRICT_IDS:      # unused by the step; kept because it consumes the global RNG (params.py:261-262)
            self.DISTRICT_HOSPITALS[name] = [f'c_{name}_{i}' for i in list(np.random.randint(0, 1000, size=10))]

        if self.district_type == 'old':
            self.set_old_district_seed(seed_infected=3)
        elif self.district_type == 'new':
            self.set_new_district_seed(seed_infected=3)

        self.normal_interaction_matrix_file = normal_interaction_matrix_file
        self.lockdown_interaction_matrix_file = lockdown_interaction_matrix_file
        self.set_interaction_parameters(self.normal_interaction_matrix_file)

        self.blocked = False
        self.lockdown = False
        self.intra_district_decreased_mobility_rates_file = intra_district_decreased_mobility_rates_file

    # ------------------------------------------------------------------ interaction structure (params.py:278-297)
    def set_interaction_parameters(self, interaction_matrix_file):
        matrix, sizes = read_interaction_workbook(interaction_matrix_file)
        self.ECONOMIC_STATUS_INTERACTION_MATRIX = matrix
        self.ECONOMIC_STATUS_INTERACTION_SIZE_MAP = sizes
        self.ECONOMIC_STATUS_INTERACTION_MATRIX_CUMSUM = matrix.cumsum(axis=1)
        names = list(self.ECONOMIC_STATUS_INTERACTION_MATRIX_CUMSUM.columns)
        self.ECON_STAT_ID_TO_NAME = dict(enumerate(names))
        self.ECON_STAT_NAME_TO_ID = {n: i for i, n in self.ECON_STAT_ID_TO_NAME.items()}
        self.ECONOMIC_STATUS_INTERACTION_MATRIX_CUMSUM_VALUES = self.ECONOMIC_STATUS_INTERACTION_MATRIX_CUMSUM.values
        self.ECONOMIC_STATUS_INTERACTION_MATRIX_VALUES = matrix[names].values

    # ------------------------------------------------------------------ lockdown / blocking (params.py:299-341)
    def set_intra_district_decreased_mobility_rates(self, intra_district_decreased_mobility_rates_file):
        rates = pd.read_csv(get_data_dir('preprocessed', 'mobility', intra_district_decreased_mobility_rates_file),
                            index_col=0)['pctdif_distance']
        self.LOCKDOWN_DECREASED_MOBILITY_RATE = {self.DISTRICT_NAME_TO_ID[name]: r for name, r in rates.items()}

    _LOCKDOWN_ALLOWED = {'lockdown_empirical': 0.67, 'lockdown_assumed': 0.05, 'lockdown_eased': 0.86}
    _BLOCK_ALLOWED = {'block_empirical': 0.67, 'block_assumed': 0.05, 'block_eased': 0.86}

    def set_lockdown_parameters(self, lockdown_mode='lockdown_empirical'):
        self.set_intra_district_decreased_mobility_rates(self.intra_district_decreased_mobility_rates_file)
        if lockdown_mode not in self._LOCKDOWN_ALLOWED:
            raise ValueError(f'lockdown_mode `{lockdown_mode}` not valid!')
        allowed = self._LOCKDOWN_ALLOWED[lockdown_mode]      # share of inter-district movers that keeps travelling
        self.LOCKDOWN_ALLOWED_PROBABILITY = {d: allowed for d in self.DISTRICT_ID_TO_NAME}
        self.set_interaction_parameters(self.lockdown_interaction_matrix_file)
        self.lockdown = True

    def set_blocked_parameters(self, block_mode='block_empirical'):
        if block_mode not in self._BLOCK_ALLOWED:
            raise ValueError(f'block_mode `{block_mode}` not valid!')
        self.BLOCK_ALLOWED_PROBABILITY = self._BLOCK_ALLOWED[block_mode]
        self.blocked = True

    # ------------------------------------------------------------------ seeding (params.py:343-371)
    def set_old_district_seed(self, seed_infected):
        self.SEED_INFECT_DISTRICT_IDS = np.array([18, 85, 21])
        self.SEED_INFECT_AGE_MIN = 20
        self.SEED_INFECT_NUM = seed_infected
        self.SIMULATION_START_DATE = datetime(2020, 6, 28, 8)

    def set_new_district_seed(self, seed_infected):
        with open(get_data_dir('preprocessed', 'line_list', 'latest_line_list.pickle'), 'rb') as fl:
            reported = pickle.load(fl)
        self.DISTRICT_ID_INFECTED_COUNT = {self.DISTRICT_NAME_TO_ID.get(name): n for name, n in reported.items()}
        share = pd.Series(self.DISTRICT_ID_INFECTED_COUNT)
        self.DISTRICT_ID_INFECTED_PROB = (share / share.sum()).to_dict()
        self.SEED_INFECT_DISTRICT_IDS = np.array(list(self.DISTRICT_ID_INFECTED_COUNT))
        self.SEED_INFECT_AGE_MIN = 20
        self.SEED_INFECT_AGE_MAX = 60
        self.SEED_INFECT_NUM = seed_infected
        self.SIMULATION_START_DATE = datetime(2020, 9, 1, 8)
        self.data_file_name = get_data_dir(
            'preprocessed', 'census', f'zimbabwe_expanded_census_consolidated_{self.data_sample_size}pct.pickle')

    # ------------------------------------------------------------------ hand-washing risk (params.py:373-438)
    def get_effective_R0(self, hw):
        self.HW_MIN_TO_MEAN_INCREASE = 0.05
        self.HW_MIN_TO_MEAN_DELTA = self.MEAN_HW_RISK - self.MIN_HW_RISK
        return 1 + (self.HW_MIN_TO_MEAN_INCREASE * (hw - self.MEAN_HW_RISK) / self.HW_MIN_TO_MEAN_DELTA)

    def _handwashing(self, scenario, severe_column, hw_override):
        self.SCENARIO = scenario
        self.risk_data = pd.read_csv(get_data_dir('preprocessed', 'risk', 'hw_and_severe_disease_risk.csv'))
        hw = self.risk_data[['ID_2', 'mean_hw_risk_pop_weighted']].copy()
        self.MEAN_HW_RISK = hw['mean_hw_risk_pop_weighted'].mean()
        self.MIN_HW_RISK = hw['mean_hw_risk_pop_weighted'].min()
        hw['ID_2'] = hw['ID_2'].map(lambda x: f'd_{x}')
        if hw_override == 'min':
            hw['mean_hw_risk_pop_weighted'] = hw['mean_hw_risk_pop_weighted'].min()
        elif hw_override == 'zero':
            hw['mean_hw_risk_pop_weighted'] = 0
        hw['effective_hw_risk'] = hw['mean_hw_risk_pop_weighted'].map(self.get_effective_R0)
        self.DISTRICT_HW_RISK = hw.set_index('ID_2').to_dict()['effective_hw_risk']
        severe = self.risk_data[['ID_2', severe_column]].copy()
        severe[severe_column] = severe[severe_column] / self.MEAN_HOSPITALIZATION_RATE
        self.DISTRICT_SEVERE_DISEASE_RISK = severe.set_index('ID_2').to_dict()[severe_column]

    def scenario_handwashing_risk(self):
        self._handwashing('HANDWASHING_RISK', 'severe_covid_risk', None)

    def scenario_improved_handwashing_risk_1(self):
        self._handwashing('HANDWASHING_RISK_1', 'severe_covid_risk_improved_1', 'min')

    def scenario_improved_handwashing_risk_2(self):
        self._handwashing('HANDWASHING_RISK_2', 'severe_covid_risk_improved_2', 'zero')

    def scenario_test_interaction_matrix_sensitivity(self):
        self.SCENARIO = 'INTERACTION_MATRIX_SENSITIVITY'
        self.set_interaction_parameters(get_data_dir('raw', 'sensitivity_interaction_matrix.xlsx'))

    # ------------------------------------------------------------------ length of stay (params.py:576-607)
    def get_gamma_shape_scale(self, mean, std):
        """Quirk Q7 of the reference, kept: the resulting Gamma has variance `std`, not `std ** 2`."""
        return mean ** 2 / std, std / mean

    def _stay_mean_std(self, weekday, src, dst):
        m = self.DISTRICT_WEEKDAY_OD_STAY_COUNT_MATRIX
        return m.at[(weekday, src, dst), 'avg_duration'], m.at[(weekday, src, dst), 'stddev_duration']

    def get_district_movement_stay_period(self, weekday, src, dst):
        try:
            shape, scale = self.get_gamma_shape_scale(*self._stay_mean_std(weekday, src, dst))
            return timedelta(hours=np.random.gamma(shape, scale))
        except KeyError:
            return timedelta(hours=24)

    def get_district_movement_stay_parameters(self, weekday, src, dst):
        try:
            return self.get_gamma_shape_scale(*self._stay_mean_std(weekday, src, dst))
        except KeyError:
            return (24, 0.01)


# ---------------------------------------------------------------------------------------------------- scenario tables
_OLD_INBOUND = ['d_901', 'd_921', 'd_922', 'd_302', 'd_406']
_OLD_OUTBOUND = ['d_901', 'd_922', 'd_304', 'd_21', 'd_302']
_OLD_BOTH = ['d_901', 'd_921', 'd_922', 'd_302', 'd_406', 'd_304', 'd_21']
_NEW_GREATEST = ['d_2', 'd_31', 'd_18', 'd_1', 'd_36', 'd_7', 'd_26', 'd_23', 'd_28', 'd_56']

# method -> (SCENARIO, restricted districts)                                   params.py:440-466
_BLOCK_SCENARIOS = {
    'scenario_block_greatest_inbound_movement': ('BLOCK_GREATEST_INBOUND', _OLD_INBOUND),
    'scenario_block_greatest_outbound_movement': ('BLOCK_GREATEST_OUTBOUND', _OLD_OUTBOUND),
    'scenario_block_greatest_movement': ('BLOCK_GREATEST', _OLD_BOTH),
    'scenario_block_new_district_greatest_movement': ('BLOCK_GREATEST_NEW_DIST', _NEW_GREATEST),
}
# method -> (SCENARIO, restricted districts or None = every district, lockdown mode)   params.py:468-494, 526-558
_LOCKDOWN_SCENARIOS = {
    'scenario_lockdown_greatest_inbound_movement': ('LOCKDOWN_GREATEST_INBOUND', _OLD_INBOUND, 'lockdown_empirical'),
    'scenario_lockdown_greatest_outbound_movement': ('LOCKDOWN_GREATEST_OUTBOUND', _OLD_OUTBOUND, 'lockdown_empirical'),
    'scenario_lockdown_greatest_movement': ('LOCKDOWN_GREATEST', _OLD_BOTH, 'lockdown_empirical'),
    'scenario_lockdown_new_district_greatest_movement': ('LOCKDOWN_GREATEST_NEW_DIST', _NEW_GREATEST, 'lockdown_empirical'),
    'scenario_continued_all_lockdown': ('CONTINUED_ALL_LOCKDOWN', None, 'lockdown_empirical'),
    'scenario_eased_all_lockdown': ('EASED_ALL_LOCKDOWN', None, 'lockdown_eased'),
}
# method -> (base lockdown method, SCENARIO, extra attributes)                  params.py:532-574
_DERIVED_SCENARIOS = {
    'scenario_continued_all_lockdown_open_mining': ('scenario_continued_all_lockdown', 'CONTINUED_ALL_LOCKDOWN_MINING', {}),
    'scenario_continued_all_lockdown_open_manufacturing': ('scenario_continued_all_lockdown', 'CONTINUED_ALL_LOCKDOWN_MANUFACTURING', {}),
    'scenario_continued_all_lockdown_open_schools': ('scenario_continued_all_lockdown', 'CONTINUED_ALL_LOCKDOWN_SCHOOLS', {}),
    'scenario_continued_all_lockdown_open_schools_seed_kids': (
        'scenario_continued_all_lockdown', 'CONTINUED_ALL_LOCKDOWN_SCHOOLS_SEED_KIDS',
        {'SEED_INFECT_AGE_MIN': 8, 'SEED_INFECT_AGE_MAX': 18}),
    'scenario_continued_all_lockdown_open_manufacturing_schools': (
        'scenario_continued_all_lockdown', 'CONTINUED_ALL_LOCKDOWN_MANUFACTURING_SCHOOLS', {}),
    'scenario_eased_all_lockdown_open_schools': ('scenario_eased_all_lockdown', 'EASED_ALL_LOCKDOWN_SCHOOLS', {}),
    'scenario_phase1_government_open_schools': ('scenario_eased_all_lockdown', 'PHASE1_GOVERNMENT_OPEN_SCHOOLS', {}),
    'scenario_dynamic_phase1_government_open_schools': (
        'scenario_eased_all_lockdown', 'DYNAMIC_PHASE1_GOVERNMENT_OPEN_SCHOOLS', {}),
    'scenario_accelerated_government_open_schools': ('scenario_eased_all_lockdown', 'ACCELERATED_GOVERNMENT_OPEN_SCHOOLS', {}),
}
# method -> (SCENARIO, attributes)                                              params.py:496-524
_SIMPLE_SCENARIOS = {
    'scenario_block_greatest_reach_movement': ('BLOCK_GREATEST_REACH', {}),
    'scenario_isolate_symptomatic_population': ('ISOLATE_SYMPTOMATIC', {'MILD_SYMPTOM_MOVEMENT_PROBABILITY': 0.05}),
    'scenario_isolate_symptomatic_population_50pct': ('ISOLATE_SYMPTOMATIC_50pct', {'MILD_SYMPTOM_MOVEMENT_PROBABILITY': 0.5}),
    'scenario_isolate_vulnerable_groups_in_house': ('ISOLATE_VULNERABLE_HOUSE', {'VULNERABLE_AGE': 60}),
    'scenario_mask_hygiene': ('MASK_HYGIENE', {'EFFECTIVE_TRANSMISSION': 0.85}),
}


def _install_scenarios(cls):
    def block(name, scenario, districts):
        def method(self):
            self.SCENARIO = scenario
            self.BLOCK_DISTRICTS = list(districts)
            self.BLOCK_DISTRICTS_IDS = [self.DISTRICT_NAME_TO_ID.get(d) for d in self.BLOCK_DISTRICTS]
            self.set_blocked_parameters()
        return method

    def lockdown(name, scenario, districts, mode):
        def method(self):
            self.SCENARIO = scenario
            self.LOCKDOWN_DISTRICTS = list(self.DISTRICT_NAME_TO_ID.keys()) if districts is None else list(districts)
            self.LOCKDOWN_DISTRICTS_IDS = [self.DISTRICT_NAME_TO_ID.get(d) for d in self.LOCKDOWN_DISTRICTS]
            self.set_lockdown_parameters(lockdown_mode=mode)
        return method

    def derived(name, base, scenario, attrs):
        def method(self):
            getattr(self, base)()
            self.SCENARIO = scenario
            for k, v in attrs.items():
                setattr(self, k, v)
        return method

    def simple(name, scenario, attrs):
        def method(self):
            self.SCENARIO = scenario
            for k, v in attrs.items():
                setattr(self, k, v)
        return method

    made = {}
    made.update({n: block(n, *v) for n, v in _BLOCK_SCENARIOS.items()})
    made.update({n: lockdown(n, *v) for n, v in _LOCKDOWN_SCENARIOS.items()})
    made.update({n: derived(n, *v) for n, v in _DERIVED_SCENARIOS.items()})
    made.update({n: simple(n, *v) for n, v in _SIMPLE_SCENARIOS.items()})
    for n, fn in made.items():
        fn.__name__ = n
        fn.__qualname__ = f'{cls.__name__}.{n}'
        setattr(cls, n, fn)

    def scenario_isolate_vulnerable_groups(self):               # params.py:510-514
        self.SCENARIO = 'ISOLATE_VULNERABLE'
        self.VULNERABLE_AGE = 60
        self.VULNERABLE_LOCATION = -1 * (max(self.DISTRICT_ID_TO_NAME) + 1)
    cls.scenario_isolate_vulnerable_groups = scenario_isolate_vulnerable_groups


_install_scenarios(ParamsConfig)


def log_to_file(fname, message, as_log=True, verbose=True, delim='$$'):
    """Append one line to `fname`: '<wall clock> $$ <message>' (the format load_model_logs_df parses,
    /root/reference/src/covid19_abm/data_plotting.py:30-64)."""
    line = f'{datetime.now()} {delim} {message}' if as_log else message
    with open(fname, 'a+') as fl:
        fl.write(line + '\n')
    if verbose:
        print(message)
