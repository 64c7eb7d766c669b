"""Synthetic "Zimbabwe-shaped" inputs (the census / CDR data of the reference are private).

Shapes follow SURVEY.md section 8(d): 60 districts with the 2012 census population shares, household
size 1 + Poisson(3.07) (mean 4.07), census economic-status frequencies, stay-home mass U(0.82, 0.998)
on the OD diagonal, stay duration mean U(24, 48) h / spread U(10, 60).  Everything is seedable and
generated directly as numpy arrays in canonical (district, household) order, so a 15 M-agent
population takes seconds and a 1 B-agent one can be generated shard by shard.
"""
from datetime import datetime

import numpy as np

from . import constants as C
from .spec import ModelSpec, Population


def district_names(n_districts):
    """`d_1..d_D` sorted as strings: district id = position in this list (params.py:255-257)."""
    return sorted(f'd_{i}' for i in range(1, n_districts + 1))


def district_shares(n_districts, seed=0):
    """Population share per district id (string-sorted name order)."""
    base = np.array(C.DISTRICT_POP_60, dtype=np.float64)
    if n_districts <= 60:
        pop_by_number = base[:n_districts]
    else:
        # replicate the 60 real shares and add a Zipf(~1) tail so a few districts stay dominant
        rng = np.random.default_rng(seed + 7919)
        reps = int(np.ceil(n_districts / 60))
        pop_by_number = np.tile(base, reps)[:n_districts] * (1.0 / (1 + np.arange(n_districts) // 60))
        pop_by_number = pop_by_number * rng.uniform(0.7, 1.3, n_districts)
    names = district_names(n_districts)
    number = np.array([int(s[2:]) for s in names])
    shares = pop_by_number[number - 1]
    return shares / shares.sum()


def make_spec(n_districts=60, R0=1.9, seed=0, start_datetime=datetime(2020, 9, 1, 0), step_minutes=240,
              **kw):
    rng = np.random.default_rng(seed + 104729)
    D = n_districts
    shares = district_shares(D, seed)
    stay_home = rng.uniform(0.82, 0.998, size=(7, D))
    w = shares[None, None, :] * rng.uniform(0.0, 1.0, size=(7, D, D))
    idx = np.arange(D)
    w[:, idx, idx] = 0.0
    w = w / np.maximum(w.sum(axis=2, keepdims=True), 1e-300) * (1.0 - stay_home)[:, :, None]
    w[:, idx, idx] = stay_home
    od_cum = np.cumsum(w, axis=2)
    od_cum[:, :, -1] = 1.0
    stay_mean = rng.uniform(24.0, 48.0, size=(7, D, D)) + 0.001
    stay_std = rng.uniform(10.0, 60.0, size=(7, D, D)) + 0.001
    return ModelSpec(n_districts=D, R0=R0, step_minutes=step_minutes, start_datetime=start_datetime,
                     od_cum=od_cum, stay_mean=stay_mean, stay_std=stay_std, **kw)


def _ages_for_status(rng, status):
    n = status.shape[0]
    adult = 18 + np.minimum(rng.exponential(15.0, n), 81.0).astype(np.int32)
    age = adult
    in_school = status == 1
    age = np.where(in_school, rng.integers(5, 19, n), age)
    not_working = status == 0
    small_child = not_working & (rng.random(n) < 0.40)
    age = np.where(small_child, rng.integers(0, 5, n), age)
    disabled = status == 8
    age = np.where(disabled, rng.integers(0, 100, n), age)
    return np.clip(age, 0, 99).astype(np.int32)


def make_population(n_agents, spec: ModelSpec, seed=0, district_range=None, scenario_flags=False):
    """Canonical-order population of ~n_agents over all districts (or the districts in
    `district_range=(lo, hi)`, for shard-by-shard generation: household ids are then local)."""
    D = spec.n_districts
    A = spec.n_status
    shares = district_shares(D, seed)
    counts = np.floor(shares * n_agents).astype(np.int64)
    counts[np.argmax(counts)] += n_agents - counts.sum()
    lo, hi = (0, D) if district_range is None else district_range
    freq = np.array(C.ECON_STATUS_FREQ[:A], dtype=np.float64)
    freq = freq / freq.sum()
    names = spec.status_names
    wk = np.array([C.WEEKDAY_MOVE_PROB[s] for s in names])
    ot = np.array([C.OTHER_DAY_MOVE_PROB[s] for s in names])
    moving_status = (wk > 0) & np.array([s not in ('In School', 'Teachers') for s in names])  # params.py:244-246

    parts = []
    hh_base = 0
    for d in range(lo, hi):
        rng = np.random.default_rng([seed, 15485863, d])
        nd = int(counts[d])
        if nd == 0:
            continue
        est = int(nd / 4.07 * 1.1) + 16
        sizes = 1 + rng.poisson(3.07, est)
        while sizes.sum() < nd:
            sizes = np.concatenate([sizes, 1 + rng.poisson(3.07, est)])
        cut = int(np.searchsorted(np.cumsum(sizes), nd, side='left'))
        sizes = sizes[:cut + 1].copy()
        sizes[-1] -= sizes.sum() - nd
        hh = np.repeat(np.arange(sizes.shape[0], dtype=np.int64), sizes) + hh_base
        hh_base += sizes.shape[0]
        status = rng.choice(A, size=nd, p=freq).astype(np.int32)
        age = _ages_for_status(rng, status)
        parts.append((np.full(nd, d, dtype=np.int32), hh, age, status))

    district = np.concatenate([p[0] for p in parts])
    household = np.concatenate([p[1] for p in parts])
    age = np.concatenate([p[2] for p in parts])
    status = np.concatenate([p[3] for p in parts])
    disabled_id = names.index('Disabled and not working') if 'Disabled and not working' in names else -1
    econ_kind = np.where(status == disabled_id, C.EKIND_HOUSEHOLD, C.EKIND_HOME_DISTRICT).astype(np.int8)
    mover = (moving_status[status] & (age >= C.DISTRICT_MOVEMENT_ALLOWED_AGE)).astype(np.int8)
    pop = Population(district=district, household=household, age=age, status=status, econ_kind=econ_kind,
                     district_mover=mover, move_prob_weekday=wk[status].copy(), move_prob_other=ot[status].copy())
    if scenario_flags:
        rng = np.random.default_rng([seed, 32452843])
        n = pop.n
        school_id = names.index('In School')
        pop.extras['school_goers'] = (status == school_id).astype(np.int8)
        worker = (status >= 3) & (status <= 6)
        pop.extras['manufacturing_workers'] = (worker & (rng.random(n) < 0.016 * 4)).astype(np.int8)
    return pop


def lockdown_rates(n_districts, seed=0):
    """`intra_district_decreased_mobility_rates` stand-in: 1 - |pctdif| ~ U(0.4, 1.0) per district."""
    return np.random.default_rng(seed + 1299709).uniform(0.4, 1.0, n_districts)


def seed_case_counts(pop: Population, spec: ModelSpec, seed=0, n_seed_districts=6):
    """Line-list stand-in: {district id: reported cases} over the most populous districts."""
    rng = np.random.default_rng([seed, 49979687])
    D = spec.n_districts
    order = np.argsort(-np.bincount(pop.district, minlength=D), kind='stable')
    seed_districts = order[:min(n_seed_districts, D)]
    cases = rng.integers(5, 60, seed_districts.shape[0])
    return {int(d): int(c) for d, c in zip(seed_districts, cases)}


def pick_seeds(pop: Population, spec: ModelSpec, infect_num=None, seed=0, cases=None):
    """Seed ids as `Country.load_agents` draws them (base_model.py:780-796): per line-list district
    int(p * infect_num) + 1 agents with age strictly in (20, 60), without replacement."""
    rng = np.random.default_rng([seed, 86028121])
    n = pop.n
    if infect_num is None:
        infect_num = max(1, int(round(0.00015 * n)))
    if cases is None:
        cases = seed_case_counts(pop, spec, seed)
    total = float(sum(cases.values()))
    bounds = np.searchsorted(pop.district, np.arange(spec.n_districts + 1))
    out = []
    for d, c in cases.items():
        lo, hi = bounds[d], bounds[d + 1]
        cand = lo + np.nonzero((pop.age[lo:hi] > 20) & (pop.age[lo:hi] < 60))[0]
        k = min(int(c / total * infect_num) + 1, cand.shape[0])
        out.append(rng.choice(cand, size=k, replace=False))
    return np.sort(np.concatenate(out)).astype(np.int64)


# ---------------------------------------------------------------------------------------------------- data directory
def mining_camp_column(pop: Population, n_districts, seed=0, share=0.04):
    """Census column `mining_district_id` (scenario_models.py:232-236; census notebook cells 41, 44): a share of the
    adult workers of every fifth district live and work in that district's mining camp, 'mining_<district>'; '' for
    everybody else."""
    d_names = np.array(district_names(n_districts))[pop.district]
    miner = (pop.district % 5 == 2) & (pop.status >= 3) & (pop.status <= 6) & (pop.age >= 18) & (
        np.random.default_rng(seed + 911).random(pop.n) < share)
    return np.where(miner, np.char.add('mining_', d_names), '')


def school_census_columns(pop: Population, n_districts, status_names=None):
    """Census columns of the school-phase scenarios (scenario_models.py:356-440): pupils get a school of ~400 in
    their home district (`school_id`, `school_id_district`) and a re-opening phase 1..5 by age band (`phase`)."""
    status_names = list(C.ECONOMIC_STATUS_NAMES if status_names is None else status_names)
    d_names = np.array(district_names(n_districts))[pop.district]
    pupil = pop.status == status_names.index('In School')
    school_no = (np.cumsum(pupil) // 400).astype(str)
    return dict(school_id=np.where(pupil, np.char.add(np.char.add('s_', d_names), np.char.add('_', school_no)), ''),
                school_id_district=np.where(pupil, d_names, ''), phase=np.where(pupil, 1 + (pop.age % 5), 0))


def population_to_dataframe(pop: Population, n_districts, seed=0, status_names=None, school_columns=False,
                            mining_columns=False, mining_share=0.04):
    """Census-style DataFrame with the columns `Country.load_agents` reads.
    Zero-padded household names keep the string sort of base_model.py:577-581 equal to the numeric order."""
    import pandas as pd
    names = np.array(district_names(n_districts))
    rng = np.random.default_rng(seed + 5)
    hh_names = np.char.add('h_', np.char.zfill(pop.household.astype(str), 9))
    d_names = names[pop.district]
    econ_loc = np.where(pop.econ_kind == C.EKIND_HOUSEHOLD, hh_names, d_names)
    status_names = np.array(C.ECONOMIC_STATUS_NAMES if status_names is None else status_names)
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
    if school_columns:
        for column, values in school_census_columns(pop, n_districts, status_names).items():
            df[column] = values
    if mining_columns:
        df['mining_district_id'] = mining_camp_column(pop, n_districts, seed, mining_share)
    return df


def _xlsx_sheet_xml(rows):
    def col_name(j):
        s = ''
        j += 1
        while j:
            j, r = divmod(j - 1, 26)
            s = chr(65 + r) + s
        return s

    def cell(v, i, j):
        ref = f'{col_name(j)}{i + 1}'
        if isinstance(v, str):
            return f'<c r="{ref}" t="inlineStr"><is><t>{v.replace("&", "&amp;").replace("<", "&lt;")}</t></is></c>'
        return f'<c r="{ref}"><v>{v!r}</v></c>'
    body = ''.join(f'<row r="{i + 1}">' + ''.join(cell(v, i, j) for j, v in enumerate(r)) + '</row>' for i, r in enumerate(rows))
    return ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
            '<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"><sheetData>' + body +
            '</sheetData></worksheet>')


def write_interaction_workbook(path, status_names, matrix, sizes):
    """Minimal .xlsx with the two sheets `ParamsConfig.set_interaction_parameters` reads (params.py:282-288)."""
    import zipfile
    matrix = np.asarray(matrix, dtype=float)
    sheet_m = [[''] + list(status_names)] + [[n] + [float(v) for v in matrix[i]] for i, n in enumerate(status_names)]
    sheet_s = [['', 'interactions']] + [[n, int(sizes[i])] for i, n in enumerate(status_names)]
    ct = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
          '<Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">'
          '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
          '<Default Extension="xml" ContentType="application/xml"/>'
          '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/>'
          '<Override PartName="/xl/worksheets/sheet1.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/>'
          '<Override PartName="/xl/worksheets/sheet2.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/>'
          '</Types>')
    rels = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
            '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
            '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/officeDocument" Target="xl/workbook.xml"/>'
            '</Relationships>')
    wb = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
          '<workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
          'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships"><sheets>'
          '<sheet name="interactions" sheetId="1" r:id="rId1"/><sheet name="interaction_matrix" sheetId="2" r:id="rId2"/>'
          '</sheets></workbook>')
    wb_rels = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
               '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
               '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/worksheet" Target="worksheets/sheet1.xml"/>'
               '<Relationship Id="rId2" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/worksheet" Target="worksheets/sheet2.xml"/>'
               '</Relationships>')
    with zipfile.ZipFile(path, 'w', zipfile.ZIP_DEFLATED) as z:
        z.writestr('[Content_Types].xml', ct)
        z.writestr('_rels/.rels', rels)
        z.writestr('xl/workbook.xml', wb)
        z.writestr('xl/_rels/workbook.xml.rels', wb_rels)
        z.writestr('xl/worksheets/sheet1.xml', _xlsx_sheet_xml(sheet_s))
        z.writestr('xl/worksheets/sheet2.xml', _xlsx_sheet_xml(sheet_m))


def write_data_dir(root, spec: ModelSpec, pop: Population, seed_cases=None, sample_tag=100, seed=0,
                   lockdown_interaction_scale=0.5, school_columns=False, mining_columns=False, mining_share=0.04):
    """Synthesise the data directory `run_scenario` expects (the real inputs are private): census pickle, OD csv,
    stay pickle, line list, intra-district mobility rates, interaction workbooks (normal / lockdown / sensitivity)
    and the risk csv.  Point COVID19_ABM_DATA_DIR at `root` to use it.  Returns the census file name."""
    import os
    import pickle
    import pandas as pd
    D = spec.n_districts
    names = district_names(D)
    for sub in ('raw', 'preprocessed/census', 'preprocessed/mobility', 'preprocessed/line_list', 'preprocessed/risk', 'logs'):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    mob = os.path.join(root, 'preprocessed', 'mobility')
    idx = pd.MultiIndex.from_product([range(7), names], names=['weekday', 'home_region'])
    pd.DataFrame(spec.od_cum.reshape(7 * D, D), index=idx, columns=names).to_csv(
        os.path.join(mob, 'daily_region_transition_probability-new-district-pre-lockdown.csv'), float_format='%.17g')
    sidx = pd.MultiIndex.from_product([range(7), names, names], names=['weekday', 'home_region', 'region'])
    pd.DataFrame({'avg_duration': (spec.stay_mean - 0.001).reshape(-1), 'stddev_duration': (spec.stay_std - 0.001).reshape(-1)},
                 index=sidx).to_pickle(os.path.join(mob, 'weekday_mobility_duration_count_df-new-district.pickle'))
    pd.DataFrame({'pctdif_distance': lockdown_rates(D, seed)}, index=names).to_csv(
        os.path.join(mob, 'intra_district_decreased_mobility_rates.csv'))
    if seed_cases is None:
        seed_cases = seed_case_counts(pop, spec, seed)
    with open(os.path.join(root, 'preprocessed', 'line_list', 'latest_line_list.pickle'), 'wb') as fl:
        pickle.dump({names[d]: c for d, c in seed_cases.items()}, fl)
    raw = os.path.join(root, 'raw')
    write_interaction_workbook(os.path.join(raw, 'final_close_interaction_matrix_normal.xlsx'), spec.status_names,
                               spec.interaction_matrix, spec.interaction_sizes)
    # the lockdown workbook is not shipped with the reference: same mixing, fewer contacts (stated choice)
    write_interaction_workbook(os.path.join(raw, 'final_close_interaction_matrix_lockdown.xlsx'), spec.status_names,
                               spec.interaction_matrix, np.maximum(1, np.rint(spec.interaction_sizes * lockdown_interaction_scale)))
    A = spec.n_status
    sens = np.zeros((A, A))
    sens[:A - 1, :A - 1] = 1.0 / (A - 1)
    sens[A - 1, A - 1] = 1.0
    write_interaction_workbook(os.path.join(raw, 'sensitivity_interaction_matrix.xlsx'), spec.status_names, sens, np.full(A, 11))
    rng = np.random.default_rng(seed + 613)
    number = np.array([int(s[2:]) for s in names])
    hw = rng.uniform(0.2, 0.7, D)
    severe = rng.uniform(0.05, 0.16, D)
    pd.DataFrame({'ID_2': number, 'mean_hw_risk_pop_weighted': hw, 'severe_covid_risk': severe,
                  'severe_covid_risk_improved_1': severe * 0.9, 'severe_covid_risk_improved_2': severe * 0.8}).to_csv(
        os.path.join(root, 'preprocessed', 'risk', 'hw_and_severe_disease_risk.csv'), index=False)
    census = os.path.join(root, 'preprocessed', 'census', f'zimbabwe_expanded_census_consolidated_{sample_tag}pct.pickle')
    population_to_dataframe(pop, D, seed, spec.status_names, school_columns=school_columns,
                            mining_columns=mining_columns, mining_share=mining_share).to_pickle(census)
    return census
