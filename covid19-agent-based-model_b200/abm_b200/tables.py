"""Constant tables the kernels (and, as plain arrays, the oracle) consume.

Every "uniform -> value" mapping whose floating-point evaluation could differ between numpy and CUDA
is tabulated here once, on the host, as integers:
  * dwell-time quantile tables (minutes) for the four sampled periods of `set_epidemic_status`
    (base_model.py:868-874, 1011-1017, 1036-1050; distributions params.py:145-189),
  * half-normal quantile table (fixed point 2^-16) for the length-of-stay transform,
  * fixed-point infection rates (2^-24) per age and symptomatic status (params.py:236-242) and the
    matching household log-survival increments -log(1 - 3 r) (base_model.py:184-187),
  * the hazard -> infection-probability table 1 - exp(-i/256) in 0.32 fixed point (the remainder of the
    hazard goes through an integer Taylor polynomial, see oracle.hazard_threshold_fx),
  * 31-bit Bernoulli thresholds ceil(p * 2^31) for mover draws and OD destination rows
    (base_model.py:316-318, 372-377).
What remains in floating point on the device (the district x status hazard and the length-of-stay
transform) is a fixed sequence of IEEE-754 double +, *, / evaluated without FMA contraction, mirrored op
for op by the oracle.
"""
import numpy as np
from scipy import stats

from . import constants as C
from .spec import ModelSpec, Population

QBITS = 12
QN = 1 << QBITS                 # 4096 bins in the main table and in the tail table
TWO31 = 2147483648.0
TWO32 = 4294967296.0
RATE_SCALE = 16777216.0         # 2^24: fixed-point unit of the rate / log-survival tables
OMEXP_N = 8192                  # entries of the 1 - exp(-i/256) table: hazards up to 32
DWELL_INCUBATION, DWELL_ASYM_TO_CONTAGIOUS, DWELL_SYM_RECOVERY, DWELL_ASYM_RECOVERY = 0, 1, 2, 3


def _quantile_knots(ppf):
    """[2, QN+1] knots: row 0 = Q(i/QN); row 1 refines the last bin down to 1 - 2^-25."""
    i = np.arange(QN + 1, dtype=np.float64)
    main_u = i / QN
    main_u[QN] = (QN - 1) / QN          # unused knot, keep finite
    tail_u = 1.0 - (QN - i) / QN / QN
    tail_u[QN] = 1.0 - 0.5 / QN / QN
    return np.stack([ppf(main_u), ppf(tail_u)])


def dwell_tables():
    """uint32 [4, 2, QN+1] minutes."""
    day = C.TICKS_PER_DAY
    dists = [
        stats.gamma(C.INCUBATION_GAMMA[0], scale=C.INCUBATION_GAMMA[1]),
        stats.expon(scale=C.ASYM_TO_CONTAGIOUS_MEAN_DAYS),
        stats.gamma(C.SYM_RECOVERY_GAMMA[0], scale=C.SYM_RECOVERY_GAMMA[1]),
        stats.gamma(C.ASYM_RECOVERY_GAMMA[0], scale=C.ASYM_RECOVERY_GAMMA[1]),
    ]
    out = np.stack([np.rint(_quantile_knots(d.ppf) * day) for d in dists])
    return np.ascontiguousarray(out.astype(np.uint32))


def znorm_table():
    """int32 [2, QN+1]: half-normal quantile * 2^16."""
    knots = _quantile_knots(lambda u: stats.norm.ppf(0.5 + 0.5 * u))
    return np.ascontiguousarray(np.rint(knots * 65536.0).astype(np.int32))


def omexp_table():
    """uint32 [OMEXP_N]: round((1 - exp(-i / 256)) * 2^32)."""
    i = np.arange(OMEXP_N, dtype=np.float64)
    return np.minimum(np.rint(-np.expm1(-i / 256.0) * TWO32), TWO32 - 1).astype(np.uint32)


def bern_threshold(p):
    """T with  (x >> 1) < T  <=>  (x >> 1) / 2^31 < p  for every 32-bit x."""
    p = np.clip(np.asarray(p, dtype=np.float64), 0.0, 1.0)
    return np.ceil(p * TWO31).astype(np.uint32)


def move_classes(pop: Population, mild_prob: float):
    """Dictionary-encode the per-agent (weekday, other-day) movement probabilities.

    Returns (move_class uint16 [N], move_thr uint32 [n_classes, 4]) with columns
    (weekday healthy, weekday mild-symptom, other-day healthy, other-day mild-symptom); each
    threshold is ceil((p * mild) * 2^31), the product taken in float64 exactly as
    base_model.py:318 does."""
    # few distinct values per column: encode each column by itself (1-D), then the pairs of small codes
    uw, iw = np.unique(pop.move_prob_weekday, return_inverse=True)
    uo, io = np.unique(pop.move_prob_other, return_inverse=True)
    pair_code = iw.astype(np.int64) * uo.shape[0] + io
    codes, inv = np.unique(pair_code, return_inverse=True)
    uniq = np.stack([uw[codes // uo.shape[0]], uo[codes % uo.shape[0]]], axis=1)
    if uniq.shape[0] > 65535:
        raise ValueError('more than 65535 distinct movement-probability pairs')
    thr = np.stack([bern_threshold(uniq[:, 0] * 1.0), bern_threshold(uniq[:, 0] * mild_prob),
                    bern_threshold(uniq[:, 1] * 1.0), bern_threshold(uniq[:, 1] * mild_prob)], axis=1)
    return inv.reshape(-1).astype(np.uint16), np.ascontiguousarray(thr), uniq


def build_tables(spec: ModelSpec):
    """All population-independent tables, as a dict of contiguous numpy arrays."""
    D, A, P = spec.n_districts, spec.n_status, spec.n_pools
    if A > C.MAX_STATUS:
        raise ValueError(f'at most {C.MAX_STATUS} economic statuses supported')
    at = spec.age_tables()
    pad = lambda v: np.concatenate([v, np.repeat(v[-1], 128 - v.shape[0])])
    r = np.stack([pad(at['r_asym']), pad(at['r_sym'])])                      # [2, 128]
    rfx = np.rint(r * RATE_SCALE).astype(np.uint32)
    hw = np.ones(1) if spec.hw_risk is None else np.asarray(spec.hw_risk, dtype=np.float64)
    hr = C.HOUSEHOLD_RATE_FACTOR * r[None, :, :] * hw[:, None, None]         # [hw_rows, 2, 128]
    if np.any(hr >= 1.0):
        raise ValueError('household infection probability >= 1')
    qfx = np.rint(-np.log1p(-hr) * RATE_SCALE).astype(np.uint32)
    if qfx.max() * C.BIG_HOUSEHOLD >= 2 ** 32:
        raise ValueError('household hazard sum would overflow 32 bits')
    W = (spec.interaction_sizes[:, None] * spec.interaction_matrix) * (1.0 / RATE_SCALE)   # [A(infector), A(contact)]
    pool_district = np.concatenate([np.arange(D, dtype=np.int32),
                                    np.asarray(spec.extra_pool_district, dtype=np.int32)])
    pool_hw = np.ones(P) if spec.hw_risk is None else hw[pool_district]
    od_thr = bern_threshold(spec.od_cum)
    od_thr = np.maximum.accumulate(od_thr, axis=2)        # rows are cumulative; enforce monotone
    shape = spec.stay_mean ** 2 / spec.stay_std           # params.py:576-580 (quirk Q7: variance = std)
    c9 = 1.0 / (9.0 * shape)
    stay_par = np.stack([spec.stay_mean, 1.0 - c9, np.sqrt(c9)], axis=3)      # Wilson-Hilferty (mean, a, b)
    return dict(
        n_districts=D, n_status=A, n_pools=P, hw_rows=int(hw.shape[0]),
        rfx=np.ascontiguousarray(rfx), qfx=np.ascontiguousarray(qfx),
        W=np.ascontiguousarray(W), pool_hw=np.ascontiguousarray(pool_hw, dtype=np.float64),
        pool_district=pool_district,
        od_thr=np.ascontiguousarray(od_thr), stay_par=np.ascontiguousarray(stay_par),
        dwell=dwell_tables(), znorm=znorm_table(), omexp=omexp_table(),
        p_hosp=np.ascontiguousarray(pad(at['p_hosp'])), p_crit=np.ascontiguousarray(pad(at['p_crit'])),
        p_hfat=np.ascontiguousarray(pad(at['p_hfat'])),
        severe_risk=np.ascontiguousarray(spec.severe_risk, dtype=np.float64),
        symptomatic_rate=float(spec.symptomatic_rate), critical_fatality_rate=float(spec.critical_fatality_rate),
        mild_active=bool(spec.mild_symptom_move_prob < 1.0),
        step_ticks=int(spec.step_minutes),
        start_minute_of_day=spec.start_datetime.hour * 60 + spec.start_datetime.minute,
        start_weekday=spec.start_datetime.weekday(),
    )
