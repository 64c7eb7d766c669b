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
