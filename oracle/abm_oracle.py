"""CPU oracle of the covid19_abm agent step -- TEST INFRASTRUCTURE, not a product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module; the product (covid19-agent-based-model_b200/) never does and fails loudly without its CUDA
library.

What it is: a numpy restatement of one sub-step of the reference,
    EpidemicScheduler.step            /root/reference/src/covid19_abm/base_model.py:33-179
    update_agent_states               base_model.py:226-272
    get_active_ids / go_home / go_out base_model.py:195-224, 274-340
    return_to_home_district           base_model.py:342-360
    move_to_other_district            base_model.py:362-441
    Country.set_epidemic_status       base_model.py:838-1055
    Country.log_model_output          base_model.py:1082-1106
operating on reference-style vectors (epidemic_state, clinical_state, current_location_ids, date_* ...)
with three deliberate, documented differences (SURVEY.md sections 0, 8(c), A.4):
  1. time is an int32 minute tick since simulation start instead of a datetime object; `datetime.max`
     is T_NEVER.  Every reference comparison is `date < now`, kept strict.
  2. every np.random call site is replaced by a word of a counter-based Philox4x32-10 block keyed
     (seed; agent or agent-group, step, site) (oracle/philox.py); dwell times come from host-built quantile tables,
     the length of stay from a Wilson-Hilferty transform of a table-based normal quantile.
  3. the explicit contact sampling of base_model.py:76-175 is restated in its mean-field
     "pressure table" form: an outside location L infects a susceptible of status b with probability
     1 - exp(-hw[L] * sum_a k_a M[a,b] R[L,a] / Ncnt[L,b]); a household infects each susceptible
     member present with probability 1 - prod_i (1 - 3 r_i) over its contagious members present.
Parity pins: (i) tests/golden/ref_functions_*.npz -- outputs of the REAL reference functions
(update_agent_states, go_home, go_out, set_epidemic_status, log_model_output) run in the build
container with these same draws injected (generator: tests/golden/make_ref_function_fixtures.py);
(ii) tests/golden/ref_ensemble_*.json -- ensemble curves of the real reference with its own RNG, which
bound difference 3 statistically (generator: tests/golden/make_ref_ensemble.py).  The reference ships
no tests or golden vectors of its own (tests/conftest.py:1-10 is empty).
"""
import numpy as np

from . import philox

T_NEVER = np.int32(2**31 - 1)
TWO31 = 2147483648.0
TWO32 = 4294967296.0
QBITS = 12
QN = 1 << QBITS
DAY = 1440
OMEXP_N = 8192


def u31(x):
    """Uniform in [0, 1) with 31 bits: the float the reference would have drawn."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(1)).astype(np.float64) * (1.0 / TWO31)


def sample_quantile(table2, x):
    """Integer inverse-CDF lookup with linear interpolation; the last bin uses the refined tail row."""
    x = np.asarray(x, dtype=np.uint32).astype(np.uint64)
    t = table2.astype(np.int64)
    i = (x >> np.uint64(32 - QBITS)).astype(np.int64)
    f = (x & np.uint64((1 << (32 - QBITS)) - 1)).astype(np.int64)
    im = np.minimum(i, QN - 2)
    main = t[0][im] + (((t[0][im + 1] - t[0][im]) * f) >> (32 - QBITS))
    sub = 32 - 2 * QBITS
    j = f >> sub
    g = f & ((1 << sub) - 1)
    tail = t[1][j] + (((t[1][j + 1] - t[1][j]) * g) >> sub)
    return np.where(i < QN - 1, main, tail)


def hazard_threshold_fx(omexp, h):
    """31-bit Bernoulli threshold of the infection probability 1 - exp(-h / 2^32) for a hazard h given
    in units of 2^-32 (uint64), in integer arithmetic -- mirrored op for op in CUDA
    (abm_kernels.cuh hazard_threshold_fx): with h = i / 256 + l, l < 2^-8,
        1 - exp(-h) = Q[i] + u - Q[i] u,   Q = omexp table,   u = l - l^2/2 + l^3/6."""
    h = np.asarray(h, dtype=np.uint64)
    big = h >= np.uint64(OMEXP_N << 24)
    hh = np.where(big, np.uint64(0), h)
    i = (hh >> np.uint64(24)).astype(np.int64)
    l = hh & np.uint64(0xFFFFFF)
    q = omexp[i].astype(np.uint64)
    l2 = (l * l) >> np.uint64(33)
    l3 = (((l2 * l) >> np.uint64(32)) * np.uint64(171)) >> np.uint64(9)     # == // 3 for operands < 256
    u = l - l2 + l3
    p = np.minimum(q + u - ((q * u) >> np.uint64(32)), np.uint64(0xFFFFFFFF))
    thr = ((p >> np.uint64(1)) + (p & np.uint64(1))).astype(np.uint32)
    return np.where(big, np.uint32(0x80000000), thr).astype(np.uint32)


def hazard_to_fx(lam):
    """Hazard (float64 >= 0) -> units of 2^-32, truncated; saturates at the end of the omexp table."""
    lam = np.asarray(lam, dtype=np.float64)
    cap = float(OMEXP_N >> 8)
    return np.where(lam >= cap, np.uint64(OMEXP_N << 24),
                    np.floor(np.minimum(lam, cap) * TWO32).astype(np.uint64)).astype(np.uint64)


class OracleModel:
    """Reference-style agent vectors + the restated step.

    pop : object with arrays district, household, age, status, econ_kind, econ_pool (or None),
          district_mover, move_prob_weekday, move_prob_other   (canonical order)
    par : dict with od_cum [7,D,D], mild_prob, and optional hw flag
    tab : dict of host-built tables (see covid19-agent-based-model_b200/abm_b200/tables.py)
    agent_id_base : global id of local agent 0 (Philox key) for household-aligned shards
    agent_ids : global id of every local agent, for shards that are not one contiguous range (overrides the base)
    reduce_fn : callable(np.ndarray uint64) -> summed over ranks (None = single rank)
    """

    def __init__(self, pop, par, tab, seed, agent_id_base=0, reduce_fn=None, agent_ids=None):
        self.tab, self.par, self.seed = tab, par, int(seed)
        self.reduce_fn = reduce_fn
        n = pop.district.shape[0]
        self.n = n
        self.D, self.A, self.P = tab['n_districts'], tab['n_status'], tab['n_pools']
        # global canonical ids (the Philox keys): one contiguous range, or given per agent for a shard made of several
        self.agent_ids = (np.arange(n, dtype=np.int64) + int(agent_id_base) if agent_ids is None
                          else np.ascontiguousarray(agent_ids, dtype=np.int64))
        self.step_ticks = tab['step_ticks']
        self.k = 0
        # --- agent vectors (base_model.py:572-634)
        self.district_ids = pop.district.astype(np.int64)
        self.H = int(pop.household.max()) + 1 if n else 0
        self.household_ids = pop.household.astype(np.int64) + self.D            # :577-581
        self.age = pop.age.astype(np.int64)
        self.economic_status_ids = pop.status.astype(np.int64)
        econ_loc = np.where(pop.econ_kind == 0, self.household_ids, self.district_ids)
        if pop.econ_pool is not None:
            econ_loc = np.where(pop.econ_kind == 2, self.D + self.H + (pop.econ_pool.astype(np.int64) - self.D),
                                econ_loc)
        self.economic_activity_location_ids = econ_loc
        self.economic_status_weekday_movement_probability = pop.move_prob_weekday.astype(np.float64).copy()
        self.economic_status_other_day_movement_probability = pop.move_prob_other.astype(np.float64).copy()
        self.mild_symptom_movement_probability = np.ones(n)
        self.current_location_ids = self.household_ids.copy()                   # :621
        self.current_district_ids = self.district_ids.copy()
        self.return_district_at = np.full(n, T_NEVER, dtype=np.int32)
        self.district_mover = pop.district_mover.astype(np.int64).copy()
        # --- epidemic / clinical vectors (base_model.py:527-570)
        self.epidemic_state = np.zeros(n, dtype=np.int8)
        self.clinical_state = np.zeros(n, dtype=np.int8)
        self.infected_symptomatic_status = np.full(n, -1, dtype=np.int8)
        for name in ('date_infected', 'date_start_contagious', 'date_start_symptomatic', 'date_recovered',
                     'date_start_hospitalized', 'date_end_hospitalized', 'date_start_critical',
                     'date_end_critical', 'date_died'):
            setattr(self, name, np.full(n, T_NEVER, dtype=np.int32))
        self.infected_at_district_ids = np.full(n, -1, dtype=np.int64)
        self.infected_at_location_ids = np.full(n, -1, dtype=np.int64)
        self.log_rows = []
        self.last_tables = None

    # ------------------------------------------------------------------ clock
    def now(self):
        return self.k * self.step_ticks

    def _abs_minute(self):
        return self.tab['start_minute_of_day'] + self.now()

    def hour(self):
        return (self._abs_minute() // 60) % 24

    def weekday(self):
        return (self.tab['start_weekday'] + self._abs_minute() // DAY) % 7

    def infection_rate_fx(self, ids):
        sym = (self.infected_symptomatic_status[ids] == 1).astype(np.int64)
        return self.tab['rfx'][sym, self.age[ids]].astype(np.uint64)

    # ------------------------------------------------------------------ base_model.py:1082-1106
    def counters(self, now):
        return dict(
            current_contagious_count=int((self.date_start_contagious < now).sum()),
            infected_count=int((self.date_infected < now).sum()),
            current_exposed_cases=int((self.epidemic_state == 1).sum()),
            current_contagious_cases=int((self.epidemic_state == 2).sum()),
            current_hospitalized_cases=int((self.clinical_state == 1).sum()),
            current_critical_cases=int((self.clinical_state == 2).sum()),
            asymptomatic_count=int((self.infected_symptomatic_status == 0).sum()),
            symptomatic_count=int((self.infected_symptomatic_status == 1).sum()),
            hospitalized_count=int((self.date_start_hospitalized < now).sum()),
            critical_count=int((self.date_start_critical < now).sum()),
            died_count=int((self.date_died < now).sum()),
            recovered_count=int((self.date_recovered < now).sum()),
        )

    # ------------------------------------------------------------------ base_model.py:226-272
    def update_agent_states(self, now):
        dead = (self.date_died < now) & (self.epidemic_state < 4)
        self.epidemic_state[dead] = 4
        self.current_location_ids[dead] = -1
        self.clinical_state[dead] = 3

        recovered = (self.date_recovered < now) & (self.epidemic_state < 3)
        self.epidemic_state[recovered] = 3
        back_home = recovered & (self.date_start_hospitalized < now)
        self.current_location_ids[back_home] = self.household_ids[back_home]
        self.clinical_state[recovered] = 3
        self.mild_symptom_movement_probability[recovered] = 1.0

        critical = (self.date_start_critical < now) & (self.clinical_state < 2)
        self.clinical_state[critical] = 2
        hospitalised = (self.date_start_hospitalized < now) & (self.clinical_state < 1)
        self.clinical_state[hospitalised] = 1
        self.current_location_ids[hospitalised] = -self.current_district_ids[hospitalised]

    # ------------------------------------------------------------------ base_model.py:195-224
    def get_active(self, now):
        active = (self.clinical_state < 1) | (self.epidemic_state == 3)
        active &= self.current_location_ids >= 0
        if self.par['mild_prob'] < 1:
            sick = active & (self.date_start_symptomatic < now) & (self.epidemic_state < 3)
            self.mild_symptom_movement_probability[sick] = self.par['mild_prob']
        return active

    # ------------------------------------------------------------------ base_model.py:274-302
    def go_home(self, now):
        active = self.get_active(now)
        at_home_district = active & (self.return_district_at == T_NEVER)
        self.current_location_ids[at_home_district] = self.household_ids[at_home_district]
        due = active & (self.return_district_at < now)
        self.current_location_ids[due] = self.household_ids[due]
        self.return_district_at[due] = T_NEVER
        self.current_district_ids[due] = self.district_ids[due]

    # ------------------------------------------------------------------ base_model.py:304-441
    def go_out(self, now):
        x1 = philox.group_draws(self.seed, self.k, self.agent_ids, philox.SITE_MOVER)
        x2 = philox.group_draws(self.seed, self.k, self.agent_ids, philox.SITE_OD)
        active = self.get_active(now)
        if self.weekday() <= 4:
            econ_move_prob = self.economic_status_weekday_movement_probability
        else:
            econ_move_prob = self.economic_status_other_day_movement_probability
        mover = active & (u31(x1) < econ_move_prob * self.mild_symptom_movement_probability)     # :316-318

        returning = mover & (self.return_district_at < now)                                        # :342-360
        self.current_location_ids[returning] = self.economic_activity_location_ids[returning]
        self.return_district_at[returning] = T_NEVER
        self.current_district_ids[returning] = self.district_ids[returning]

        cand = np.nonzero(mover & (self.district_mover == 1) & (self.return_district_at == T_NEVER))[0]  # :366-370
        cum = self.par['od_cum'][self.weekday()]
        target = np.empty(cand.shape[0], dtype=np.int64)
        for lo in range(0, cand.shape[0], 1 << 18):                                               # :372-377
            c = cand[lo:lo + (1 << 18)]
            target[lo:lo + c.shape[0]] = (u31(x2[c])[:, None] < cum[self.district_ids[c]]).argmax(axis=1)
        other = target != self.district_ids[cand]                                                 # :410
        leaving, dst = cand[other], target[other]
        src = self.district_ids[leaving]
        par = self.tab['stay_par'][self.weekday(), src, dst]                                       # :418-428 (Q7)
        xs = philox.group_draws(self.seed, self.k, self.agent_ids[leaving], philox.SITE_STAY)
        mag = sample_quantile(self.tab['znorm'], (xs.astype(np.uint64) << np.uint64(1)) & np.uint64(0xFFFFFFFF))
        z = np.where(xs >> np.uint32(31) == 1, -mag, mag).astype(np.float64) * (1.0 / 65536.0)
        t = par[:, 1] + par[:, 2] * z
        hours = par[:, 0] * ((t * t) * t)
        minutes = np.floor(np.clip(hours * 60.0, 0.0, 1.0e9)).astype(np.int64)
        self.return_district_at[leaving] = np.minimum(now + minutes, int(T_NEVER) - 1).astype(np.int32)   # :430-434
        self.current_district_ids[leaving] = dst
        self.current_location_ids[leaving] = dst

        normal = mover & ~returning
        normal[leaving] = False                                                                    # :331-339
        self.current_location_ids[normal] = self.economic_activity_location_ids[normal]

    # ------------------------------------------------------------------ base_model.py:60-175 (mean field)
    def pool_of(self, loc):
        """Outside-pool row of a location id, -1 for households / non-zero hospital / dead."""
        pool = np.full(loc.shape, -1, dtype=np.int64)
        d = (loc >= 0) & (loc < self.D)        # district pools; hospital of district 0 is location 0 (Q3)
        pool[d] = loc[d]
        s = loc >= self.D + self.H             # school / extra pools
        pool[s] = self.D + (loc[s] - self.D - self.H)
        return pool

    def pressure_tables(self):
        loc = self.current_location_ids
        pool = self.pool_of(loc)
        outside = pool >= 0
        key = pool[outside] * self.A + self.economic_status_ids[outside]
        ncnt = np.bincount(key, minlength=self.P * self.A).astype(np.uint64)
        cont = np.nonzero((self.epidemic_state == 2) & outside)[0]
        ckey = pool[cont] * self.A + self.economic_status_ids[cont]
        rsum = np.zeros(self.P * self.A, dtype=np.uint64)
        np.add.at(rsum, ckey, self.infection_rate_fx(cont))
        buf = np.concatenate([rsum, ncnt])
        if self.reduce_fn is not None:
            buf = self.reduce_fn(buf)
        return buf[:self.P * self.A].reshape(self.P, self.A), buf[self.P * self.A:].reshape(self.P, self.A)

    def outside_thresholds(self, rsum, ncnt):
        W, hw = self.tab['W'], self.tab['pool_hw']
        rd = rsum.astype(np.float64)
        acc = np.zeros((self.P, self.A))
        for a in range(self.A):
            acc = acc + W[a][None, :] * rd[:, a][:, None]
        with np.errstate(divide='ignore', invalid='ignore'):
            lam = hw[:, None] * (acc / ncnt.astype(np.float64))
        thr = hazard_threshold_fx(self.tab['omexp'], hazard_to_fx(np.where(ncnt > 0, lam, 0.0)))
        return np.where(ncnt > 0, thr, 0).astype(np.uint32)

    def household_hazard(self):
        loc = self.current_location_ids
        at_home = (loc >= self.D) & (loc < self.D + self.H)
        cont = np.nonzero((self.epidemic_state == 2) & at_home)[0]
        sym = (self.infected_symptomatic_status[cont] == 1).astype(np.int64)
        row = self.district_ids[cont] if self.tab['hw_rows'] > 1 else np.zeros(cont.shape[0], dtype=np.int64)
        q = self.tab['qfx'][row, sym, self.age[cont]].astype(np.uint64)
        hsum = np.zeros(self.H, dtype=np.uint64)
        np.add.at(hsum, loc[cont] - self.D, q)
        return at_home, hsum

    def infect(self, now):
        rsum, ncnt = self.pressure_tables()
        thr_out = self.outside_thresholds(rsum, ncnt)
        at_home, hsum = self.household_hazard()
        self.last_tables = dict(rsum=rsum, ncnt=ncnt, thr_out=thr_out, hsum=hsum)
        sus = self.epidemic_state < 1
        loc = self.current_location_ids
        pool = self.pool_of(loc)
        thr = np.zeros(self.n, dtype=np.uint32)
        o = sus & (pool >= 0)
        thr[o] = thr_out[pool[o], self.economic_status_ids[o]]
        h = np.nonzero(sus & at_home)[0]
        hs = hsum[loc[h] - self.D]
        exposed = hs > 0
        thr[h[exposed]] = hazard_threshold_fx(self.tab['omexp'], hs[exposed] << np.uint64(8))
        cand = np.nonzero(thr > 0)[0]
        x0 = philox.group_draws(self.seed, self.k, self.agent_ids[cand], philox.SITE_INFECT)
        newly = cand[(x0 >> np.uint32(1)) < thr[cand]]
        self.set_epidemic_status(newly, now)
        return newly

    # ------------------------------------------------------------------ base_model.py:838-1055
    def set_epidemic_status(self, ids, now=None):
        ids = np.asarray(ids, dtype=np.int64)
        if ids.size == 0:
            return
        now = self.now() if now is None else now
        tab = self.tab
        a0, a1, a2, a3 = philox.agent_draws(self.seed, self.k, self.agent_ids[ids], philox.SITE_COURSE_A)
        b0, b1, b2, _ = philox.agent_draws(self.seed, self.k, self.agent_ids[ids], philox.SITE_COURSE_B)
        self.infected_at_district_ids[ids] = self.current_district_ids[ids]
        self.infected_at_location_ids[ids] = self.current_location_ids[ids]
        self.epidemic_state[ids] = 1
        self.date_infected[ids] = now
        sym = u31(a0) < tab['symptomatic_rate']                                                   # :853-857
        self.infected_symptomatic_status[ids] = np.where(sym, 1, 0)

        s = ids[sym]
        age = self.age[s]
        onset = now + sample_quantile(tab['dwell'][0], a1[sym])                                   # :868-878
        self.date_start_symptomatic[s] = onset
        self.date_start_contagious[s] = onset - 12 * 60                                           # :880-882
        p_h = tab['p_hosp'][age] * tab['severe_risk'][self.current_district_ids[s]]               # :885-896
        hosp = u31(a2[sym]) < p_h
        crit = hosp & (u31(a3[sym]) < tab['p_crit'][age])                                         # :903-912
        dies = np.where(crit, u31(b0[sym]) < tab['critical_fatality_rate'],                       # :919-933
                        hosp & (u31(b1[sym]) < tab['p_hfat'][age]))
        hosp_start = onset + 5 * DAY
        hosp_end = hosp_start + 8 * DAY
        crit_end = hosp_end + 8 * DAY
        self.date_start_hospitalized[s[hosp]] = hosp_start[hosp]                                  # :940-949
        self.date_end_hospitalized[s[hosp]] = hosp_end[hosp]
        self.date_start_critical[s[crit]] = hosp_end[crit]                                        # :951-959
        self.date_end_critical[s[crit]] = crit_end[crit]
        died = onset + 21 * DAY                                                                   # :967-994
        died = np.where(crit, np.maximum(died, crit_end), died)
        self.date_died[s[dies]] = died[dies]
        rec = onset + sample_quantile(tab['dwell'][2], b2[sym])                                   # :1008-1022
        rec = np.where(crit, crit_end, np.where(hosp, hosp_end, rec))                             # :996-1006
        self.date_recovered[s[~dies]] = rec[~dies]

        a = ids[~sym]                                                                             # :1024-1055
        cont = now + sample_quantile(tab['dwell'][1], a1[~sym])
        self.date_start_contagious[a] = cont
        self.date_recovered[a] = cont + sample_quantile(tab['dwell'][3], b2[~sym])

    # ------------------------------------------------------------------ base_model.py:33-179
    def step(self):
        now, hour = self.now(), self.hour()
        if hour == 8:                                               # :36, :1083
            row = self.counters(now)
            row['tick'] = now
            self.log_rows.append(row)
        self.update_agent_states(now)                               # :38
        if hour == 16:                                              # :42-43
            self.go_home(now)
        if hour == 8:                                               # :46-47
            self.go_out(now)
        newc = (self.date_start_contagious < now) & (self.epidemic_state < 2)   # :54-58
        self.epidemic_state[newc] = 2
        self.infect(now)                                            # :60-175
        self.k += 1                                                 # :177-179

    def state_dict(self):
        names = ['epidemic_state', 'clinical_state', 'infected_symptomatic_status', 'current_location_ids',
                 'current_district_ids', 'return_district_at', 'date_infected', 'date_start_contagious',
                 'date_start_symptomatic', 'date_recovered', 'date_start_hospitalized', 'date_end_hospitalized',
                 'date_start_critical', 'date_end_critical', 'date_died', 'infected_at_district_ids',
                 'infected_at_location_ids']
        return {k: getattr(self, k).copy() for k in names}
