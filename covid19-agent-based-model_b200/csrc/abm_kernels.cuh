// abm_kernels.cuh -- sm_100a kernels of the covid19_abm agent step.
//
// One launch of abm_substep_kernel per 4-hour sub-step does, for every agent slot, in this order:
//   I(k-1)  infection draw of the previous sub-step against the district x status threshold table and the
//           household hazard (reference: EpidemicScheduler.step location loop, base_model.py:60-175, in
//           its pressure-table form) and, on infection, the whole disease course
//           (Country.set_epidemic_status, base_model.py:838-1055)
//   log     the four "current" counters of Country.log_model_output (base_model.py:1092-1095) at hour 8
//   T(k)    time-to-next-event transitions (update_agent_states, base_model.py:226-272) and the
//           contagious promotion (base_model.py:54-58)
//   M(k)    go_home at hour 16 / go_out + OD mobility at hour 8 (base_model.py:274-441)
//   A(k)    accumulation of the district x status infectious-pressure and headcount tables
// Fusing I(k-1) into the launch of sub-step k means the 8-byte hot record of an agent is read once
// per sub-step.  abm_hazard_table_kernel then turns the accumulated tables into 31-bit Bernoulli
// thresholds for the next launch.
//
// All random numbers are Philox4x32-10 keyed (seed; agent id, sub-step, site).  Everything that is
// compared against the oracle is integer arithmetic or IEEE double +,*,/ through the __d*_rn
// intrinsics (never contracted to FMA), so results are bit-identical to oracle/abm_oracle.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abm_b200.h"

namespace abm {

// ---- flags word (constants.py F_*)
constexpr uint32_t F_EPI_MASK = 0x7u, F_CLIN_SHIFT = 3, F_CLIN_MASK = 0x3u, F_SYM_SHIFT = 5, F_SYM_MASK = 0x3u;
constexpr uint32_t F_LOC_SHIFT = 7, F_LOC_MASK = 0x7u;
constexpr uint32_t F_AWAY = 1u << 10, F_DMOVER = 1u << 11;
constexpr uint32_t F_STATUS_SHIFT = 12, F_STATUS_MASK = 0xFu, F_AGE_SHIFT = 16, F_AGE_MASK = 0x7Fu;
constexpr uint32_t F_HEAD = 1u << 23, F_OUT_H = 1u << 24, F_OUT_C = 1u << 25, F_OUT_D = 1u << 26, F_ONSET = 1u << 27;
constexpr uint32_t F_EKIND_SHIFT = 28, F_EKIND_MASK = 0x3u, F_BIGHH = 1u << 30, F_FILLER = 1u << 31;
constexpr uint32_t LOC_HOME = 0, LOC_ECON = 1, LOC_DIST = 2, LOC_HOSP = 3, LOC_DEAD = 4;
constexpr uint32_t EKIND_HOUSEHOLD = 0, EKIND_HOME_DISTRICT = 1, EKIND_POOL = 2;
constexpr uint32_t EPI_S = 0, EPI_I = 1, EPI_C = 2, EPI_R = 3, EPI_D = 4;
constexpr int32_t T_NEVER = 0x7FFFFFFF;
constexpr uint32_t FIRE_NEVER = 0xFFFFu;
constexpr int32_t DAY = 1440;

// ---- step modes
constexpr uint32_t M_INFECT = 1, M_STEP = 2, M_GO_OUT = 4, M_GO_HOME = 8, M_LOG = 16, M_WEEKDAY = 32;

// ---- cumulative counters (device array cum[])
enum { CUM_CONTAGIOUS = 0, CUM_INFECTED, CUM_HOSPITALIZED, CUM_CRITICAL, CUM_DIED, CUM_RECOVERED, CUM_ASYM, CUM_SYM,
       CUR_EXPOSED, CUR_CONTAGIOUS, CUR_HOSPITALIZED, CUR_CRITICAL, N_TALLY };

constexpr int THREADS = 256;
constexpr int PER_THREAD = 4;
static_assert(THREADS * PER_THREAD == ABM_TILE, "one tile per block");

struct DevTables {
    const uint32_t* rfx; const uint32_t* qfx; const double* W; const double* pool_hw;
    const uint32_t* od_thr; const double* stay_par; const uint32_t* dwell; const int32_t* znorm;
    const double* p_hosp; const double* p_crit; const double* p_hfat; const double* severe_risk;
    const uint32_t* move_thr; const int64_t* district_start; const int64_t* tile_base;
    const uint16_t* tile_home; const int64_t* big_start;
    int32_t D, A, P, hw_rows, n_big, step_ticks, mild_active;
    uint32_t seed_lo, seed_hi;
    unsigned long long step_inv;    // floor(2^64 / step_ticks) + 1
    double symptomatic_rate, critical_fatality_rate;
};

struct DevAgents {
    uint32_t* flags; uint32_t* aux;
    int32_t* t_infected; int32_t* t_contagious; int32_t* t_onset; int32_t* t_recovered; int32_t* t_return;
    uint32_t* inf_at; const uint16_t* move_class; const uint32_t* econ_pool;
};

struct DevWork {
    unsigned long long* acc;        // [2][P][A]: R fixed-point sums, then head counts
    const uint32_t* thr_out;        // [P][A] thresholds from the previous sub-step
    unsigned long long* hbig_read;  // [n_big] household hazard of big households, previous sub-step
    unsigned long long* hbig_write; // [n_big] accumulated this sub-step
    unsigned long long* cum;        // [N_TALLY]
    long long* log_row;             // day-log row written when M_LOG
};

struct StepArgs {
    int32_t k;        // sub-step being executed (T/M/A phases)
    int32_t now;      // its minute tick
    int32_t k_inf;    // sub-step whose infection draw is applied (k - 1)
    int32_t now_inf;
    uint32_t mode;
    int32_t weekday;  // 0..6 of `now`
};

// ------------------------------------------------------------------------------------------------ Philox
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}

__device__ __forceinline__ U4 agent_draws(const DevTables& T, int64_t cid, int32_t step, uint32_t site) {
    return philox4x32_10((uint32_t)cid, (uint32_t)((uint64_t)cid >> 32), (uint32_t)step, site, T.seed_lo, T.seed_hi);
}

// Bernoulli with probability p on a 32-bit draw: (x >> 1) / 2^31 < p, as an exact double comparison.
__device__ __forceinline__ bool bernoulli(uint32_t x, double p) {
    return (double)(x >> 1) < __dmul_rn(p, 2147483648.0);
}

// Integer inverse-CDF lookup (oracle.sample_quantile): [2][QN+1] knots, last bin refined by row 1.
__device__ __forceinline__ uint32_t sample_quantile_u32(const uint32_t* __restrict__ t2, uint32_t x) {
    const uint32_t i = x >> 20, f = x & 0xFFFFFu;
    if (i < ABM_QN - 1) {
        const uint64_t lo = __ldg(t2 + i), hi = __ldg(t2 + i + 1);
        return (uint32_t)(lo + (((hi - lo) * f) >> 20));
    }
    const uint32_t j = f >> 8, g = f & 0xFFu;
    const uint64_t lo = __ldg(t2 + (ABM_QN + 1) + j), hi = __ldg(t2 + (ABM_QN + 1) + j + 1);
    return (uint32_t)(lo + (((hi - lo) * g) >> 8));
}

// 1 - exp(-x), x >= 0: fixed sequence of IEEE double ops (oracle.one_minus_exp_neg).
template <int N>
__device__ __forceinline__ double horner_omexp(double y) {
    double s = 1.0;
#pragma unroll
    for (int n = N; n > 1; --n) s = __dsub_rn(1.0, __dmul_rn(__dmul_rn(y, 1.0 / (double)n), s));
    return __dmul_rn(y, s);
}
// Taylor order by range: 9 terms below 2^-5, 13 below 2^-2, 20 below 2^-1 (all to < 1 ulp of the series)
__device__ __forceinline__ double one_minus_exp_neg(double x) {
    x = fmin(x, 40.0);
    if (x < 0.03125) return horner_omexp<9>(x);
    if (x < 0.25) return horner_omexp<13>(x);
    if (x < 0.5) return horner_omexp<20>(x);
    double e = __dsub_rn(1.0, horner_omexp<20>(__dmul_rn(x, 0.015625)));
#pragma unroll
    for (int i = 0; i < 6; ++i) e = __dmul_rn(e, e);
    return __dsub_rn(1.0, e);
}
__device__ __noinline__ uint32_t hazard_threshold(double x) {
    return (uint32_t)ceil(__dmul_rn(one_minus_exp_neg(x), 2147483648.0));
}

// sub-step index at which `date < now` first holds: floor(t / step) + 1, by multiply-high with
// inv = floor(2^64 / step) + 1 (exact for every 32-bit t)
__device__ __forceinline__ uint32_t fire_index(int32_t t, unsigned long long step_inv) {
    if (t == T_NEVER) return FIRE_NEVER;
    if (t < 0) return 0;
    const uint32_t f = (uint32_t)__umul64hi((unsigned long long)(uint32_t)t, step_inv) + 1u;
    return f < 0xFFFEu ? f : 0xFFFEu;
}

// home district from the canonical id (agents are sorted by home district)
__device__ __forceinline__ uint32_t district_of(const DevTables& T, int64_t cid) {
    int lo = 0, hi = T.D;           // district_start[lo] <= cid < district_start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(T.district_start + mid) <= cid) lo = mid; else hi = mid;
    }
    return (uint32_t)lo;
}
// `tile_home` = home district of the tile's first agent: almost always the answer for an away agent
__device__ __forceinline__ uint32_t home_district(const DevTables& T, uint32_t f, uint32_t a, int64_t cid, int tile_home = -1) {
    if (!(f & F_AWAY)) return a >> 16;
    if (tile_home >= 0 && cid < __ldg(T.district_start + tile_home + 1)) return (uint32_t)tile_home;
    return district_of(T, cid);
}
__device__ __forceinline__ int big_slot(const DevTables& T, int64_t cid) {
    int lo = 0, hi = T.n_big;       // big_start[lo] <= cid < big_start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(T.big_start + mid) <= cid) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool at_own_household(uint32_t f) {
    const uint32_t kind = (f >> F_LOC_SHIFT) & F_LOC_MASK;
    return kind == LOC_HOME || (kind == LOC_ECON && ((f >> F_EKIND_SHIFT) & F_EKIND_MASK) == EKIND_HOUSEHOLD);
}

// Outside pool (row of the district x status tables) the agent is currently in, or -1.
// Hospital is location -current_district (base_model.py:270-272): district 0 maps onto its own outside
// pool (quirk Q3); every other hospital / the dead location holds no susceptible agent.
__device__ __forceinline__ int pool_of(const DevTables& T, const DevAgents& G, uint32_t f, uint32_t a, int64_t slot, int64_t cid,
                                       int tile_home = -1) {
    const uint32_t kind = (f >> F_LOC_SHIFT) & F_LOC_MASK;
    if (kind == LOC_ECON) {
        const uint32_t ek = (f >> F_EKIND_SHIFT) & F_EKIND_MASK;
        if (ek == EKIND_HOME_DISTRICT) return (int)home_district(T, f, a, cid, tile_home);
        if (ek == EKIND_POOL) return (int)__ldg(G.econ_pool + slot);
        return -1;
    }
    if (kind == LOC_DIST) return (int)(a >> 16);
    if (kind == LOC_HOSP) return (a >> 16) == 0 ? 0 : -1;
    return -1;
}

// ------------------------------------------------------------------------------------------------ infection event
// Country.set_epidemic_status for one agent (base_model.py:838-1055).  Returns the new flags; writes the
// cold vectors; *t_next gets the sub-step at which the first scheduled event fires (>= k_first).
__device__ __forceinline__ uint32_t infect_agent(const DevTables& T, const DevAgents& G, uint32_t f, uint32_t& a, int64_t slot,
                                                 int64_t cid, int32_t k_inf, int32_t now_inf, uint32_t k_first,
                                                 unsigned int* s_tally) {
    const U4 A4 = agent_draws(T, cid, k_inf, 1u);
    const U4 B4 = agent_draws(T, cid, k_inf, 2u);
    const uint32_t age = (f >> F_AGE_SHIFT) & F_AGE_MASK;
    const uint32_t cur = a >> 16;
    const bool sym = bernoulli(A4.x, T.symptomatic_rate);                                  // :853-857
    int32_t t_c, t_o, t_r;
    uint32_t out = 0;
    if (sym) {
        t_o = now_inf + (int32_t)sample_quantile_u32(T.dwell + 0 * 2 * (ABM_QN + 1), A4.y); // :868-878
        t_c = t_o - 12 * 60;                                                                // :880-882
        const double p_h = __dmul_rn(__ldg(T.p_hosp + age), __ldg(T.severe_risk + cur));    // :885-896
        const bool hosp = bernoulli(A4.z, p_h);                                             // :898-901
        const bool crit = hosp && bernoulli(A4.w, __ldg(T.p_crit + age));                   // :903-912
        const bool dies = crit ? bernoulli(B4.x, T.critical_fatality_rate)                  // :919-921
                               : (hosp && bernoulli(B4.y, __ldg(T.p_hfat + age)));          // :923-933
        if (hosp) out |= F_OUT_H;
        if (crit) out |= F_OUT_C;
        if (dies) out |= F_OUT_D;
        if (dies) t_r = T_NEVER;                                                            // :967-994
        else if (crit) t_r = t_o + 21 * DAY;                                                // :996-1000
        else if (hosp) t_r = t_o + 13 * DAY;                                                // :1002-1006
        else t_r = t_o + (int32_t)sample_quantile_u32(T.dwell + 2 * 2 * (ABM_QN + 1), B4.z); // :1008-1022
        atomicAdd(s_tally + CUM_SYM, 1u);
    } else {                                                                                // :1024-1055
        t_o = T_NEVER;
        t_c = now_inf + (int32_t)sample_quantile_u32(T.dwell + 1 * 2 * (ABM_QN + 1), A4.y);
        t_r = t_c + (int32_t)sample_quantile_u32(T.dwell + 3 * 2 * (ABM_QN + 1), B4.z);
        atomicAdd(s_tally + CUM_ASYM, 1u);
    }
    G.t_infected[slot] = now_inf;                                                           // :851
    G.t_contagious[slot] = t_c;
    G.t_onset[slot] = t_o;
    G.t_recovered[slot] = t_r;
    G.inf_at[slot] = (((f >> F_LOC_SHIFT) & F_LOC_MASK) << 16) | cur;                       // :846-847
    f &= ~(F_EPI_MASK | (F_SYM_MASK << F_SYM_SHIFT) | F_OUT_H | F_OUT_C | F_OUT_D | F_ONSET);
    f |= EPI_I | ((sym ? 2u : 1u) << F_SYM_SHIFT) | out;                                    // :850
    // earliest scheduled event (contagious always precedes the others except when onset < 12 h)
    uint32_t fi = fire_index(t_c, T.step_inv);
    const uint32_t fo = fire_index(t_o, T.step_inv), fr = fire_index(t_r, T.step_inv);
    fi = fi < fo ? fi : fo;
    fi = fi < fr ? fi : fr;
    fi = fi > k_first ? fi : k_first;
    a = (a & 0xFFFF0000u) | fi;
    return f;
}

// ------------------------------------------------------------------------------------------------ transitions
// update_agent_states (base_model.py:226-272) + contagious promotion (base_model.py:54-58) for one agent
// whose next-event index is due.  Cumulative date counters are tallied the first time a date is passed.
__device__ __forceinline__ uint32_t transition_agent(const DevTables& T, const DevAgents& G, uint32_t f, uint32_t& a, int64_t slot,
                                                     int32_t now, unsigned int* s_tally) {
    const int32_t t_c = G.t_contagious[slot], t_o = G.t_onset[slot], t_r = G.t_recovered[slot];
    const int32_t t_h = (f & F_OUT_H) ? t_o + 5 * DAY : T_NEVER;    // params.py:165
    const int32_t t_cs = (f & F_OUT_C) ? t_o + 13 * DAY : T_NEVER;  // hospital end = critical start (base_model.py:953)
    const int32_t t_d = (f & F_OUT_D) ? t_o + 21 * DAY : T_NEVER;   // params.py:177 (== end of critical care when critical)
    uint32_t epi = f & F_EPI_MASK, clin = (f >> F_CLIN_SHIFT) & F_CLIN_MASK, loc = (f >> F_LOC_SHIFT) & F_LOC_MASK;
    if (t_c < now && epi == EPI_I) atomicAdd(s_tally + CUM_CONTAGIOUS, 1u);
    if (t_d < now && epi < EPI_D) {                                  // :230-237
        epi = EPI_D; loc = LOC_DEAD; clin = 3;
        atomicAdd(s_tally + CUM_DIED, 1u);
    }
    if (t_r < now && epi < EPI_R) {                                  // :240-254
        epi = EPI_R;
        if (t_h < now) loc = LOC_HOME;
        clin = 3;
        atomicAdd(s_tally + CUM_RECOVERED, 1u);
    }
    if (t_cs < now && clin < 2) { clin = 2; atomicAdd(s_tally + CUM_CRITICAL, 1u); }      // :257-261
    if (t_h < now && clin < 1) { clin = 1; loc = LOC_HOSP; atomicAdd(s_tally + CUM_HOSPITALIZED, 1u); }  // :263-272
    if (t_o < now) f |= F_ONSET;                                     // :212-219 (symptom onset passed)
    if (t_c < now && epi < EPI_C) epi = EPI_C;                       // :54-58
    f = (f & ~(F_EPI_MASK | (F_CLIN_MASK << F_CLIN_SHIFT) | (F_LOC_MASK << F_LOC_SHIFT))) | epi | (clin << F_CLIN_SHIFT) |
        (loc << F_LOC_SHIFT);
    uint32_t nf = FIRE_NEVER;
    const int32_t ts[6] = {t_c, t_o, t_h, t_cs, t_d, t_r};
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (ts[i] >= now) { const uint32_t x = fire_index(ts[i], T.step_inv); nf = x < nf ? x : nf; }
    a = (a & 0xFFFF0000u) | nf;
    return f;
}

__device__ __forceinline__ bool is_active(uint32_t f, uint32_t a) {
    // get_active_ids (base_model.py:198-209): not hospitalised or recovered, at a non-negative location
    const uint32_t epi = f & F_EPI_MASK, clin = (f >> F_CLIN_SHIFT) & F_CLIN_MASK, loc = (f >> F_LOC_SHIFT) & F_LOC_MASK;
    const bool loc_ok = !(loc == LOC_DEAD || (loc == LOC_HOSP && (a >> 16) != 0));
    return (clin < 1 || epi == EPI_R) && loc_ok;
}

// Length of stay for a traveller leaving `home` for `dest` on `weekday` (base_model.py:418-434): Gamma(mean^2/std,
// std/mean) hours (params.py:576-580, quirk Q7) through the Wilson-Hilferty transform of a table normal draw.
__device__ __forceinline__ int32_t stay_return_tick(const DevTables& T, uint32_t x3, int weekday, uint32_t home, uint32_t dest, int32_t now) {
    const double* par = T.stay_par + (((size_t)weekday * T.D + home) * T.D + dest) * 3;
    const uint32_t mag_x = x3 << 1;
    const uint32_t i = mag_x >> 20, fr = mag_x & 0xFFFFFu;
    int64_t mag;
    if (i < ABM_QN - 1) {
        const int64_t zl = __ldg(T.znorm + i), zh = __ldg(T.znorm + i + 1);
        mag = zl + (((zh - zl) * (int64_t)fr) >> 20);
    } else {
        const uint32_t j = fr >> 8, g = fr & 0xFFu;
        const int64_t zl = __ldg(T.znorm + (ABM_QN + 1) + j), zh = __ldg(T.znorm + (ABM_QN + 1) + j + 1);
        mag = zl + (((zh - zl) * (int64_t)g) >> 8);
    }
    const double z = __dmul_rn((double)((x3 >> 31) ? -mag : mag), 1.0 / 65536.0);
    const double t = __dadd_rn(__ldg(par + 1), __dmul_rn(__ldg(par + 2), z));
    const double hours = __dmul_rn(__ldg(par + 0), __dmul_rn(__dmul_rn(t, t), t));
    const double m = floor(fmin(fmax(__dmul_rn(hours, 60.0), 0.0), 1.0e9));
    const int64_t ret = (int64_t)now + (int64_t)m;
    return (int32_t)(ret < (int64_t)T_NEVER - 1 ? ret : (int64_t)T_NEVER - 1);
}

// first column j with u < thr[j] of a non-decreasing OD row; none -> column 0 (quirk Q8)   base_model.py:372-377
__device__ __forceinline__ uint32_t od_destination(const DevTables& T, uint32_t u, int weekday, uint32_t home) {
    const uint32_t* row = T.od_thr + ((size_t)weekday * T.D + home) * T.D;
    int lo = 0, hi = T.D;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u < __ldg(row + mid)) hi = mid; else lo = mid + 1;
    }
    return lo < T.D ? (uint32_t)lo : 0u;
}

// ------------------------------------------------------------------------------------------------ main kernel
// One block = one tile of 1024 slots staged in shared memory; one thread owns 4 consecutive slots (one
// 128-bit load of each hot word).  Work that only a few agents need in a given sub-step (the infection
// draw at night, the disease-course sampling, due transitions, length-of-stay sampling) is not executed
// in place -- a warp would run it for one or two active lanes -- but appended to per-tile queues in
// shared memory and then processed with all lanes busy.  MODE is the compile-time copy of StepArgs.mode
// for the four sub-step kinds that make up a simulated day (RUNTIME_MODE = read it from the argument).
constexpr uint32_t RUNTIME_MODE = 0xFFFFFFFFu;
constexpr int Q_DRAW = 0, Q_INFECT = 1, Q_TRANS = 2, Q_STAY = 3, N_QUEUES = 4;

__device__ __forceinline__ void enqueue_one(unsigned int* s_n, uint16_t* s_q, int q, int o) {
    const unsigned int pos = atomicAdd(s_n + q, 1u);
    s_q[q * ABM_TILE + pos] = (uint16_t)o;
}

template <uint32_t MODE>
__global__ void __launch_bounds__(THREADS, 4)
abm_substep_kernel(const DevTables T, const DevAgents G, const DevWork Wk, const StepArgs s) {
    __shared__ __align__(16) uint32_t s_f[ABM_TILE];     // staged flags
    __shared__ __align__(16) uint32_t s_a[ABM_TILE];     // staged aux
    __shared__ __align__(16) uint32_t s_thr[ABM_TILE];   // infection threshold per slot; later the stay draw
    __shared__ uint16_t s_q[N_QUEUES * ABM_TILE];
    __shared__ unsigned int s_n[N_QUEUES];
    __shared__ unsigned int s_tally[N_TALLY];
    __shared__ unsigned int s_cnt[ABM_MAX_STATUS];
    __shared__ unsigned int s_rlo[ABM_MAX_STATUS], s_rhi[ABM_MAX_STATUS];

    const uint32_t mode = MODE == RUNTIME_MODE ? s.mode : ((MODE & ~M_WEEKDAY) | (s.mode & M_WEEKDAY));
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int o0 = tid * PER_THREAD;
    const int64_t slot0 = tile * ABM_TILE + o0;
    const int64_t cid_tile = __ldg(T.tile_base + tile);
    const int L0 = (int)__ldg(T.tile_home + tile);

    const uint4 f4 = *reinterpret_cast<const uint4*>(G.flags + slot0);
    const uint4 a4 = *reinterpret_cast<const uint4*>(G.aux + slot0);
    uint32_t f[PER_THREAD] = {f4.x, f4.y, f4.z, f4.w};
    uint32_t a[PER_THREAD] = {a4.x, a4.y, a4.z, a4.w};

    if (tid < N_TALLY) s_tally[tid] = 0;
    if (tid < N_QUEUES) s_n[tid] = 0;
    if (tid < ABM_MAX_STATUS) { s_cnt[tid] = 0; s_rlo[tid] = 0; s_rhi[tid] = 0; }
    *reinterpret_cast<uint4*>(s_f + o0) = f4;
    *reinterpret_cast<uint4*>(s_a + o0) = a4;
    *reinterpret_cast<uint4*>(s_thr + o0) = make_uint4(0u, 0u, 0u, 0u);

    // ---------------------------------------------------------------- household thresholds for I(k-1)
    int tile_has_cont_home = 0;
    if (mode & M_INFECT) {
        bool any_cont_home = false;
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j)
            any_cont_home |= (f[j] & (F_EPI_MASK | F_BIGHH)) == EPI_C && at_own_household(f[j]);
        tile_has_cont_home = __syncthreads_or(any_cont_home);
        if (tile_has_cont_home) {
            // every household head sums the log-survival increments of its contagious members present
            // (1 - prod(1 - 3 r_i), base_model.py:156-175, 184-187) and publishes the threshold to all members
#pragma unroll 1
            for (int j = 0; j < PER_THREAD; ++j) {
                const int o = o0 + j;
                uint32_t c = s_f[o];
                if ((c & (F_HEAD | F_BIGHH | F_FILLER)) != F_HEAD) continue;
                int e = o;
                unsigned long long hsum = 0ull;
                const uint32_t* q = T.qfx;
                if (T.hw_rows > 1) q += (size_t)home_district(T, c, s_a[o], cid_tile + o, L0) * 256;
                do {
                    if ((c & F_EPI_MASK) == EPI_C && at_own_household(c))
                        hsum += __ldg(q + ((((c >> F_SYM_SHIFT) & F_SYM_MASK) == 2u) ? 128 : 0) + ((c >> F_AGE_SHIFT) & F_AGE_MASK));
                    ++e;
                    if (e >= ABM_TILE) break;
                    c = s_f[e];
                } while (!(c & F_HEAD));
                if (hsum) {
                    const uint32_t thr = hazard_threshold(__dmul_rn((double)hsum, 1.0 / 4294967296.0));
                    for (int m = o; m < e; ++m) s_thr[m] = thr;
                }
            }
            __syncthreads();
        }
    } else {
        __syncthreads();
    }

    // ---------------------------------------------------------------- P1: classify own agents, build the queues
    uint32_t logc = 0;     // packed "current" counters for the day log (state after I(k-1), before T(k))
    {
        int n_draw = 0;
        uint32_t draw_mask = 0;
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) {
            const uint32_t fj = f[j], aj = a[j];
            const int o = o0 + j;
            if ((mode & M_INFECT) && (fj & F_EPI_MASK) == EPI_S) {
                uint32_t thr = 0;
                if (at_own_household(fj)) {
                    if (fj & F_BIGHH) {
                        const unsigned long long hz = Wk.hbig_read[big_slot(T, cid_tile + o)];
                        if (hz) { thr = hazard_threshold(__dmul_rn((double)hz, 1.0 / 4294967296.0)); s_thr[o] = thr; }
                    } else if (tile_has_cont_home) {
                        thr = s_thr[o];
                    }
                } else {
                    const int pool = pool_of(T, G, fj, aj, slot0 + j, cid_tile + o, L0);
                    if (pool >= 0) {
                        thr = __ldg(Wk.thr_out + pool * T.A + ((fj >> F_STATUS_SHIFT) & F_STATUS_MASK));
                        if (thr) s_thr[o] = thr;
                    }
                }
                if (thr) { draw_mask |= 1u << j; ++n_draw; }
            }
            if (mode & M_STEP) {
                if (mode & M_LOG) {
                    const uint32_t epi = fj & F_EPI_MASK, clin = (fj >> F_CLIN_SHIFT) & F_CLIN_MASK;
                    logc += (epi == EPI_I ? 1u : 0u) | (epi == EPI_C ? 0x100u : 0u) | (clin == 1 ? 0x10000u : 0u) | (clin == 2 ? 0x1000000u : 0u);
                }
                if ((aj & 0xFFFFu) <= (uint32_t)s.k) enqueue_one(s_n, s_q, Q_TRANS, o);
            }
        }
        if (mode & M_INFECT) {
            // warp-aggregated append of the draw queue
            int incl = n_draw;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            int base = 0;
            if (lane == 31 && total) base = (int)atomicAdd(s_n + Q_DRAW, (unsigned int)total);
            base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - n_draw;
#pragma unroll
            for (int j = 0; j < PER_THREAD; ++j)
                if (draw_mask & (1u << j)) s_q[Q_DRAW * ABM_TILE + base++] = (uint16_t)(o0 + j);
        }
    }
    __syncthreads();

    if (mode & M_INFECT) {
        // ------------------------------------------------------------ P2: the infection draws, all lanes busy
        const int n_draw = (int)s_n[Q_DRAW];
        for (int i = tid; i < n_draw; i += THREADS) {
            const int o = s_q[Q_DRAW * ABM_TILE + i];
            const U4 X = agent_draws(T, cid_tile + o, s.k_inf, 0u);
            if ((X.x >> 1) < s_thr[o]) enqueue_one(s_n, s_q, Q_INFECT, o);
        }
        __syncthreads();
        // ------------------------------------------------------------ P3: disease course of the newly infected
        const int n_inf = (int)s_n[Q_INFECT];
        for (int i = tid; i < n_inf; i += THREADS) {
            const int o = s_q[Q_INFECT * ABM_TILE + i];
            const int64_t slot = tile * ABM_TILE + o;
            uint32_t aj = s_a[o];
            const uint32_t fj = infect_agent(T, G, s_f[o], aj, slot, cid_tile + o, s.k_inf, s.now_inf, (uint32_t)s.k_inf + 1u, s_tally);
            s_f[o] = fj; s_a[o] = aj;
            G.flags[slot] = fj; G.aux[slot] = aj;
            atomicAdd(s_tally + CUM_INFECTED, 1u);
            if (mode & M_STEP) {
                if (mode & M_LOG) atomicAdd(s_tally + CUR_EXPOSED, 1u);
                if ((aj & 0xFFFFu) <= (uint32_t)s.k) enqueue_one(s_n, s_q, Q_TRANS, o);
            }
        }
        if (mode & M_STEP) __syncthreads();
    }

    if (mode & M_STEP) {
        // ------------------------------------------------------------ T(k): due transitions
        const int n_trans = (int)s_n[Q_TRANS];
        if (n_trans) {
            for (int i = tid; i < n_trans; i += THREADS) {
                const int o = s_q[Q_TRANS * ABM_TILE + i];
                const int64_t slot = tile * ABM_TILE + o;
                uint32_t aj = s_a[o];
                const uint32_t fj = transition_agent(T, G, s_f[o], aj, slot, s.now, s_tally);
                s_f[o] = fj; s_a[o] = aj;
                G.flags[slot] = fj; G.aux[slot] = aj;
            }
        }
        if (n_trans || (s_n[Q_INFECT] && (mode & M_INFECT))) {
            __syncthreads();
            const uint4 nf = *reinterpret_cast<const uint4*>(s_f + o0);
            const uint4 na = *reinterpret_cast<const uint4*>(s_a + o0);
            f[0] = nf.x; f[1] = nf.y; f[2] = nf.z; f[3] = nf.w;
            a[0] = na.x; a[1] = na.y; a[2] = na.z; a[3] = na.w;
        }
        const uint32_t fb[PER_THREAD] = {f[0], f[1], f[2], f[3]};   // what global memory holds now
        const uint32_t ab[PER_THREAD] = {a[0], a[1], a[2], a[3]};

        // ------------------------------------------------------------ M(k): movement (base_model.py:274-441)
        if (mode & (M_GO_OUT | M_GO_HOME)) {
#pragma unroll
            for (int j = 0; j < PER_THREAD; ++j) {
                uint32_t fj = f[j], aj = a[j];
                if ((fj & F_FILLER) || !is_active(fj, aj)) continue;
                const int o = o0 + j;
                const int64_t slot = slot0 + j, cid = cid_tile + o;
                if (mode & M_GO_HOME) {                                                    // :274-302
                    if (!(fj & F_AWAY)) {
                        fj = (fj & ~(F_LOC_MASK << F_LOC_SHIFT)) | (LOC_HOME << F_LOC_SHIFT);
                    } else if (G.t_return[slot] < s.now) {
                        aj = (aj & 0xFFFFu) | (home_district(T, fj, aj, cid, L0) << 16);
                        fj = (fj & ~(F_AWAY | (F_LOC_MASK << F_LOC_SHIFT))) | (LOC_HOME << F_LOC_SHIFT);
                    }
                }
                if (mode & M_GO_OUT) {                                                     // :304-340
                    const bool mild = T.mild_active && (fj & F_ONSET) && (fj & F_EPI_MASK) < EPI_R;   // :212-219, :316
                    const uint32_t thr = __ldg(T.move_thr + 4 * (uint32_t)__ldg(G.move_class + slot) + ((mode & M_WEEKDAY) ? 0 : 2) + (mild ? 1 : 0));
                    if (thr) {
                        const bool needs_od = (fj & F_DMOVER) != 0;
                        U4 X = U4{0, 0, 0, 0};
                        if (thr < 0x80000000u || needs_od) X = agent_draws(T, cid, s.k, 0u);
                        if ((X.y >> 1) < thr) {                                            // :317-318 a mover
                            if ((fj & F_AWAY) && G.t_return[slot] < s.now) {               // :342-360
                                aj = (aj & 0xFFFFu) | (home_district(T, fj, aj, cid, L0) << 16);
                                fj &= ~F_AWAY;
                            }
                            bool left = false;
                            if (needs_od && !(fj & F_AWAY)) {                              // :366-370
                                const uint32_t home = aj >> 16;
                                const uint32_t dest = od_destination(T, X.z >> 1, s.weekday, home);
                                if (dest != home) {                                        // :410-439
                                    fj = (fj & ~(F_LOC_MASK << F_LOC_SHIFT)) | F_AWAY | (LOC_DIST << F_LOC_SHIFT);
                                    aj = (aj & 0xFFFFu) | (dest << 16);
                                    s_thr[o] = X.w;            // the stay draw, consumed by the Q_STAY pass
                                    s_a[o] = (home << 16) | dest;
                                    enqueue_one(s_n, s_q, Q_STAY, o);
                                    left = true;
                                }
                            }
                            if (!left) fj = (fj & ~(F_LOC_MASK << F_LOC_SHIFT)) | (LOC_ECON << F_LOC_SHIFT);   // :331-339, :346-352
                        }
                    }
                }
                f[j] = fj; a[j] = aj;
            }
        }

        // ------------------------------------------------------------ A(k): pressure / headcount tables
        uint32_t cw0 = 0, cw1 = 0, cw2 = 0;    // packed per-status head counts of this thread (8 bits each)
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) {
            const uint32_t fj = f[j], aj = a[j];
            const int64_t slot = slot0 + j, cid = cid_tile + o0 + j;
            const uint32_t st = (fj >> F_STATUS_SHIFT) & F_STATUS_MASK;
            const int pool = pool_of(T, G, fj, aj, slot, cid, L0);
            const bool cont = (fj & F_EPI_MASK) == EPI_C;
            if (pool >= 0) {
                const uint32_t r = cont ? __ldg(T.rfx + ((((fj >> F_SYM_SHIFT) & F_SYM_MASK) == 2u) ? 128 : 0) + ((fj >> F_AGE_SHIFT) & F_AGE_MASK)) : 0u;
                if (pool == L0) {
                    const uint32_t inc = 1u << (8 * (st & 3));
                    if ((st >> 2) == 0) cw0 += inc; else if ((st >> 2) == 1) cw1 += inc; else cw2 += inc;
                    if (cont) { atomicAdd(s_rlo + st, r & 0xFFFFu); atomicAdd(s_rhi + st, r >> 16); }
                } else {
                    atomicAdd(Wk.acc + (size_t)T.P * T.A + (size_t)pool * T.A + st, 1ull);
                    if (cont) atomicAdd(Wk.acc + (size_t)pool * T.A + st, (unsigned long long)r);
                }
            } else if (cont && (fj & F_BIGHH) && at_own_household(fj)) {
                const uint32_t qrow = T.hw_rows > 1 ? home_district(T, fj, aj, cid, L0) : 0u;
                const uint32_t q = __ldg(T.qfx + (size_t)qrow * 256 + ((((fj >> F_SYM_SHIFT) & F_SYM_MASK) == 2u) ? 128 : 0) + ((fj >> F_AGE_SHIFT) & F_AGE_MASK));
                atomicAdd(Wk.hbig_write + big_slot(T, cid), (unsigned long long)q);
            }
        }

        // ------------------------------------------------------------ write back what movement changed
        if (f[0] != fb[0] || f[1] != fb[1] || f[2] != fb[2] || f[3] != fb[3])
            *reinterpret_cast<uint4*>(G.flags + slot0) = make_uint4(f[0], f[1], f[2], f[3]);
        if (a[0] != ab[0] || a[1] != ab[1] || a[2] != ab[2] || a[3] != ab[3])
            *reinterpret_cast<uint4*>(G.aux + slot0) = make_uint4(a[0], a[1], a[2], a[3]);

        if (mode & M_LOG) {
            logc = __reduce_add_sync(0xFFFFFFFFu, logc);   // <= 128 per 8-bit field
            if (lane < 4) {
                const unsigned int v = (logc >> (8 * lane)) & 0xFFu;
                if (v) atomicAdd(s_tally + CUR_EXPOSED + lane, v);
            }
        }
        // head counts: 3 packed warp reductions, lane b keeps status b
        const uint32_t w0 = __reduce_add_sync(0xFFFFFFFFu, cw0);
        const uint32_t w1 = __reduce_add_sync(0xFFFFFFFFu, cw1);
        const uint32_t w2 = __reduce_add_sync(0xFFFFFFFFu, cw2);
        if (lane < T.A) {
            const uint32_t w = (lane >> 2) == 0 ? w0 : ((lane >> 2) == 1 ? w1 : w2);
            const unsigned int v = (w >> (8 * (lane & 3))) & 0xFFu;
            if (v) atomicAdd(s_cnt + lane, v);
        }
    }
    __syncthreads();
    if (mode & M_GO_OUT) {
        // ------------------------------------------------------------ length of stay of the departing travellers
        const int n_stay = (int)s_n[Q_STAY];
        for (int i = tid; i < n_stay; i += THREADS) {
            const int o = s_q[Q_STAY * ABM_TILE + i];
            const uint32_t hd = s_a[o];
            G.t_return[tile * ABM_TILE + o] = stay_return_tick(T, s_thr[o], s.weekday, hd >> 16, hd & 0xFFFFu, s.now);
        }
    }
    if (tid < N_TALLY) {
        const unsigned int v = s_tally[tid];
        if (v) {
            if (tid < CUR_EXPOSED) atomicAdd(Wk.cum + tid, (unsigned long long)v);
            else atomicAdd(reinterpret_cast<unsigned long long*>(Wk.log_row) + (ABM_C_CURRENT_EXPOSED_CASES + tid - CUR_EXPOSED),
                           (unsigned long long)v);
        }
    }
    if ((mode & M_STEP) && tid >= 32 && tid < 32 + T.A) {
        const int b = tid - 32;
        if (s_cnt[b]) atomicAdd(Wk.acc + (size_t)T.P * T.A + (size_t)L0 * T.A + b, (unsigned long long)s_cnt[b]);
        const unsigned long long r = ((unsigned long long)s_rhi[b] << 16) + s_rlo[b];
        if (r) atomicAdd(Wk.acc + (size_t)L0 * T.A + b, r);
    }
}

// ------------------------------------------------------------------------------------------------ table kernel
// District x status tables -> infection thresholds:
//   lambda[L][b] = hw[L] * (sum_a W[a][b] * R[L][a]) / Ncnt[L][b],  thr = ceil((1 - exp(-lambda)) * 2^31)
// then clears the accumulators for the next sub-step, snapshots them for inspection, and at log
// sub-steps copies the cumulative counters into the day-log row.
struct TableArgs {
    int32_t write_log;
    int32_t tick;
    long long seeds_before_now;
};

__global__ void abm_hazard_table_kernel(const DevTables T, unsigned long long* acc, unsigned long long* acc_snapshot,
                                        uint32_t* thr_out, unsigned long long* hbig_clear, const unsigned long long* cum,
                                        long long* log_row, const TableArgs ta) {
    const int n = T.P * T.A;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int L = i / T.A, b = i - L * T.A;
        const unsigned long long cnt = acc[n + i];
        uint32_t thr = 0;
        if (cnt) {
            double sum = 0.0;
            for (int a = 0; a < T.A; ++a)
                sum = __dadd_rn(sum, __dmul_rn(__ldg(T.W + a * T.A + b), (double)acc[L * T.A + a]));
            const double lam = __dmul_rn(__ldg(T.pool_hw + L), __ddiv_rn(sum, (double)cnt));
            thr = hazard_threshold(lam);
        }
        thr_out[i] = thr;
    }
    if (i < T.n_big) hbig_clear[i] = 0ull;
    if (blockIdx.x == 0 && ta.write_log && threadIdx.x < 8) {
        const int t = threadIdx.x;
        // cumulative date counters (base_model.py:1089-1090, 1097-1103)
        if (t < 6) log_row[t] = (long long)cum[t] + (t == CUM_INFECTED ? ta.seeds_before_now : 0);
        if (t == 6) log_row[ABM_C_ASYMPTOMATIC_COUNT] = (long long)cum[CUM_ASYM];
        if (t == 7) { log_row[ABM_C_SYMPTOMATIC_COUNT] = (long long)cum[CUM_SYM]; log_row[ABM_C_TICK] = ta.tick; }
    }
}

// second pass of the table kernel: snapshot + clear (separate launch so every thread of the first
// pass has read the full row of R before it is zeroed)
__global__ void abm_clear_tables_kernel(unsigned long long* acc, unsigned long long* acc_snapshot, int n2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) { acc_snapshot[i] = acc[i]; acc[i] = 0ull; }
}

// ------------------------------------------------------------------------------------------------ seeding
__global__ void abm_seed_kernel(const DevTables T, const DevAgents G, unsigned long long* cum, const int64_t* slots,
                                const int64_t* ids, int64_t n, int32_t k, int32_t now) {
    __shared__ unsigned int s_tally[N_TALLY];
    if (threadIdx.x < N_TALLY) s_tally[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t slot = slots[i];
        uint32_t a = G.aux[slot];
        const uint32_t f = infect_agent(T, G, G.flags[slot], a, slot, ids[i], k, now, (uint32_t)k, s_tally);
        G.flags[slot] = f;
        G.aux[slot] = a;
    }
    __syncthreads();
    if (threadIdx.x == CUM_ASYM || threadIdx.x == CUM_SYM) {
        const unsigned int v = s_tally[threadIdx.x];
        if (v) atomicAdd(cum + threadIdx.x, (unsigned long long)v);
    }
}

// ------------------------------------------------------------------------------------------------ district statistics
// get_safe_and_unsafe_districts (base_model.py:733-758): symptomatic infected-or-contagious per home district.
__global__ void abm_district_sym_kernel(const DevTables T, const DevAgents G, int64_t n_slots, unsigned long long* sym,
                                        unsigned long long* pop) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const uint32_t f = G.flags[slot];
    if (f & F_FILLER) return;
    const int64_t tile = slot / ABM_TILE;
    const int64_t cid = __ldg(T.tile_base + tile) + (slot - tile * ABM_TILE);
    const uint32_t d = home_district(T, f, G.aux[slot], cid);
    atomicAdd(pop + d, 1ull);
    const uint32_t epi = f & F_EPI_MASK;
    if ((epi == EPI_I || epi == EPI_C) && ((f >> F_SYM_SHIFT) & F_SYM_MASK) == 2u) atomicAdd(sym + d, 1ull);
}

// current-state counters over all slots (abm_counters)
__global__ void abm_count_state_kernel(const DevAgents G, int64_t n_slots, unsigned long long* out4) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (slot < n_slots) {
        const uint32_t f = G.flags[slot];
        const uint32_t epi = f & F_EPI_MASK, clin = (f >> F_CLIN_SHIFT) & F_CLIN_MASK;
        c = (epi == EPI_I ? 1u : 0u) | (epi == EPI_C ? 0x100u : 0u) | (clin == 1 ? 0x10000u : 0u) | (clin == 2 ? 0x1000000u : 0u);
    }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    const int lane = threadIdx.x & 31;
    if (lane < 4) {
        const unsigned int v = (c >> (8 * lane)) & 0xFFu;
        if (v) atomicAdd(out4 + lane, (unsigned long long)v);
    }
}

}  // namespace abm
