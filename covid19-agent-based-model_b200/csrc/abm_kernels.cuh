// abm_kernels.cuh -- sm_100a kernels of the covid19_abm agent step.
//
// One 4-hour sub-step does, for every agent, in this order:
//   I(k-1)  infection draw of the previous sub-step against the district x status threshold table and the
//           household hazard (reference: EpidemicScheduler.step location loop, base_model.py:60-175, in
//           its pressure-table form) and, on infection, the whole disease course
//           (Country.set_epidemic_status, base_model.py:838-1055)
//   T(k)    time-to-next-event transitions (update_agent_states, base_model.py:226-272) and the
//           contagious promotion (base_model.py:54-58)
//   M(k)    go_home at hour 16 / go_out + OD mobility at hour 8 (base_model.py:274-441)
//   A(k)    accumulation of the district x status infectious-pressure and headcount tables
// Fusing I(k-1) into the launch of sub-step k means the 8-byte hot record of an agent is read once per sub-step.
// The work is split by KIND OF CONTROL FLOW over three launches:
//   abm_sweep_kernel         every slot, the dense per-agent work; agents that need a rare code path are deferred
//   abm_event_kernel         one thread per deferred agent: infection events, due transitions, departures
//   abm_hazard_table_kernel  accumulated tables -> 31-bit Bernoulli thresholds for the next launch, the sum over ranks,
//                            the day log (Country.log_model_output, base_model.py:1082-1106), the clock
//
// Sweep execution model: ONE WARP OWNS ONE SUB-TILE of 128 consecutive agent slots (lane l holds slots
// 4l..4l+3, one 128-bit load per hot word).  The host layout guarantees that a sub-tile holds whole
// households of a single home district and that slot offset o of a sub-tile is canonical agent id
// base + o with base a multiple of 4, so (a) the household hazard is a warp-local segmented sum,
// (b) the home district is a per-warp scalar and (c) one Philox4x32 block serves the four agents of a
// lane.  Warps never wait for each other.  The sweep is instruction-issue bound, not HBM bound (profiles/), so it is
// organised to issue few instructions: flag fields that are tested together are adjacent, the four agents of a lane run
// through predicated straight-line code, and the hazard -> probability map is integer arithmetic.
//
// All random numbers are Philox4x32-10 keyed by the seed.  Counter = (id, id >> 32, sub-step, site):
// id = agent_id >> 2 for the group sites (word agent_id & 3 belongs to the agent), id = agent_id for
// the disease-course sites.  Everything that is compared against the oracle is integer arithmetic or
// IEEE double +,*,/ through the __d*_rn intrinsics (never contracted to FMA), so results are
// bit-identical to oracle/abm_oracle.py -- whichever kernel handles an agent.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abm_b200.h"

namespace abm {

// ---- flags word (constants.py F_*)
constexpr uint32_t F_EPI_MASK = 0x7u, F_CLIN_SHIFT = 3, F_CLIN_MASK = 0x3u, F_LOC_SHIFT = 5, F_LOC_MASK = 0x7u;
constexpr uint32_t F_AWAY = 1u << 8, F_DMOVER = 1u << 9, F_HEAD = 1u << 10, F_ONSET = 1u << 11;
constexpr uint32_t F_STATUS_SHIFT = 12, F_STATUS_MASK = 0xFu, F_AGE_SHIFT = 16, F_AGE_MASK = 0x7Fu;
constexpr uint32_t F_SYMPT = 1u << 23, F_ASYMPT = 1u << 24, F_OUT_H = 1u << 25, F_OUT_C = 1u << 26, F_OUT_D = 1u << 27;
constexpr uint32_t F_EKIND_SHIFT = 28, F_EKIND_MASK = 0x3u, F_BIGHH = 1u << 30, F_FILLER = 1u << 31;
// where the agent is: own household / outside pool of its HOME district (also while away on paper,
// quirk Q4) / outside pool of the destination of a trip / extra outside pool (school) / hospital
// (-current_district) / dead
constexpr uint32_t LOC_HOME = 0, LOC_HOMEPOOL = 1, LOC_DEST = 2, LOC_POOL = 3, LOC_HOSP = 4, LOC_DEAD = 5;
constexpr uint32_t EKIND_HOUSEHOLD = 0, EKIND_HOME_DISTRICT = 1, EKIND_POOL = 2;
constexpr uint32_t EPI_S = 0, EPI_I = 1, EPI_C = 2, EPI_R = 3, EPI_D = 4;
constexpr int32_t T_NEVER = 0x7FFFFFFF;
constexpr uint32_t FIRE_NEVER = 0xFFFFu;
constexpr int32_t DAY = 1440;
constexpr uint32_t LOCF = F_LOC_MASK << F_LOC_SHIFT;
constexpr uint32_t CLINF = F_CLIN_MASK << F_CLIN_SHIFT;
constexpr uint32_t IN_HOMEPOOL = LOC_HOMEPOOL << F_LOC_SHIFT;

// ---- Philox sites (oracle/philox.py)
constexpr uint32_t SITE_INFECT = 0, SITE_COURSE_A = 1, SITE_COURSE_B = 2, SITE_MOVER = 3, SITE_OD = 4, SITE_STAY = 5;

// ---- step modes.  DENSE_I / DENSE_A: most agents were / are in an outside pool (daytime) at the sub-step
// whose infection draw is applied / whose tables are accumulated; any data is handled correctly either way
constexpr uint32_t M_INFECT = 1, M_STEP = 2, M_GO_OUT = 4, M_GO_HOME = 8, M_LOG = 16, M_WEEKDAY = 32, M_DENSE_I = 64,
                   M_DENSE_A = 128;
// INCR: nobody moves in this sub-step, so the district x status tables (and the big-household hazards) are those of
// the previous sub-step plus the changes made by the few transitions that fire: the A(k) sweep is skipped, the
// transitions emit signed deltas, and the table kernel adds them to the sums it kept.  All sums are integers modulo
// 2^64, so the result is identical to a full recount.
constexpr uint32_t M_INCR = 256;
// run-time only: do not let the next kernel start early (set when the next kernel in the stream may be
// another sub-step kernel, which reads the agent store in its prologue: abm_flush)
constexpr uint32_t M_NO_EARLY_DEPENDENTS = 1u << 16;

// ---- counters.  cum[]: events whose date has passed (cumulative log columns); cur deltas are split in
// "pre" (infection phase: part of the state the day log sees) and "post" (transition phase: after it)
enum { CUM_CONTAGIOUS = 0, CUM_INFECTED, CUM_HOSPITALIZED, CUM_CRITICAL, CUM_DIED, CUM_RECOVERED, CUM_ASYM, CUM_SYM,
       PRE_EXPOSED, PRE_CONTAGIOUS, PRE_HOSPITALIZED, PRE_CRITICAL,
       POST_EXPOSED, POST_CONTAGIOUS, POST_HOSPITALIZED, POST_CRITICAL, N_TALLY };

#ifndef ABM_THREADS
#define ABM_THREADS 256          // threads per CTA of the sub-step kernel: one warp per sub-tile, ABM_THREADS / 32 sub-tiles per CTA
#endif
constexpr int THREADS = ABM_THREADS;
constexpr int WARPS = THREADS / 32;
constexpr int PER_LANE = 4;
constexpr int SUB = ABM_SUBTILE;
static_assert(ABM_TILE % (WARPS * SUB) == 0 && PER_LANE * 32 == SUB, "one sub-tile per warp, whole CTAs per tile");
constexpr uint32_t FULL = 0xFFFFFFFFu;

struct DevTables {
    const uint32_t* rfx; const uint32_t* qfx; const double* W; const double* pool_hw;
    const uint32_t* od_thr; const double* stay_par; const uint32_t* dwell; const int32_t* znorm;
    const double* p_hosp; const double* p_crit; const double* p_hfat; const double* severe_risk;
    const uint32_t* move_thr; const unsigned long long* sub_info; const int64_t* big_start; const uint32_t* omexp;
    int32_t D, A, P, hw_rows, n_big, step_ticks, mild_active;
    unsigned long long step_inv;    // floor(2^64 / step_ticks) + 1
    double symptomatic_rate, critical_fatality_rate;
    uint32_t rk[20];                // Philox round keys (k0, k1) x 10
};

struct DevAgents {
    uint32_t* flags; uint32_t* aux;
    int32_t* cold;      // [n_slots][ABM_COLD_WORDS]: one 32-byte sector per agent (ABM_COLD_* word indices)
    const uint16_t* move_class; const uint32_t* econ_pool; const uint32_t* hh_index;
};
// the cold record as two 16-byte halves
struct Cold { int4 t; int4 u; };    // t = (t_infected, t_contagious, t_onset, t_recovered), u = (t_return, inf_at, 0, 0)
__device__ __forceinline__ Cold load_cold(const DevAgents& G, uint32_t slot) {
    const int4* p = reinterpret_cast<const int4*>(G.cold) + (size_t)slot * 2;
    Cold c; c.t = p[0]; c.u = p[1];
    return c;
}

// Accumulators that every CTA adds to exist in `rep_mask + 1` replicas (a CTA uses replica
// blockIdx & rep_mask): atomics on one address serialise in L2, and the hottest rows (the capital's
// pool, the event tallies) would otherwise receive tens of thousands of them per launch.  The table
// kernel sums the replicas.
#ifndef ABM_MAX_REPLICAS
#define ABM_MAX_REPLICAS 16      // 64 in the first builds: 16 costs the sweeps nothing and the table kernel reads a quarter (-1.4 us per sub-step, profiles/r02c_variants_*.txt)
#endif
struct DevWork {
    unsigned long long* acc2[2];    // [replicas][2][P][A] each: R fixed-point sums, then head counts; sub-step k uses acc2[k & 1]
    uint32_t* thr_out;              // [P][A] thresholds from the previous sub-step
    unsigned long long* hbig2[2];   // [n_big] household hazard of big households: written at k into hbig2[k & 1], read from the other
    unsigned long long* tally;      // [replicas][N_TALLY]
    unsigned long long* acc_sum;    // [2][P][A] replicas (and ranks) summed: tables of the last sub-step
    unsigned long long* acc_local;  // [2][P][A] replicas of THIS rank summed (kept across incremental sub-steps)
    long long* counters;            // [12] running cumulative (0..7) and current (8..11) counters
    long long* daylog;              // [max_log_rows + 1][ABM_LOG_STRIDE]
    uint32_t rep_mask;
    int32_t max_log_rows;
    // multi-GPU exchange through peer memory (one node, NVLink / NVSwitch): every rank owns an exchange buffer
    // that its peers WRITE their rows into: [2 parities][n_ranks sources][2][P][A] elements of 16 bytes (see the table kernel)
    unsigned long long* xchg[ABM_MAX_RANKS];   // xchg[r] = rank r's buffer as mapped here (own buffer at index `rank`)
    unsigned int* rows_done;        // table-kernel rows finished in the current launch
    // event queues: sweep CTA b appends to queue b & q_mask; queue q owns entries [q * q_cap, (q + 1) * q_cap)
    uint4* q_entries;               // (slot | newly-infected << 31, flags, aux, 0)
    uint32_t* q_count;              // [q_mask + 1][2] entries used from the front (generic) / from the back (decided leavers); reset by the event kernel
    uint32_t q_mask, q_cap;
    uint32_t n_sub;                 // sub-tiles (warps' worth of slots) of this rank's store
    // optional timeline (abm_set_trace): [sub-step k][kernel 0 sweep / 1 events / 2 table][0 first CTA start, 1 last CTA end], ns
    unsigned long long* trace;
    int32_t trace_rows;
    int32_t n_ranks, rank;          // n_ranks > 1 only when the peer exchange is active
};


// The simulation clock lives in device memory and is advanced by the table kernel, so a whole simulated
// day (6 x [sub-step kernel, table kernel]) is one replayable CUDA graph with no per-launch arguments.
struct __align__(16) DevClock {
    int32_t k;            // next sub-step to execute
    int32_t now;          // its minute tick
    int32_t weekday;      // 0..6 of `now`
    int32_t minute_abs;   // start_minute_of_day + now
    int32_t weekday0;     // weekday of tick 0's day
    int32_t n_log_rows;
    long long seeds_before;   // seeded infections with tick < now     (base_model.py:796 + :1090)
    long long seeds_at_now;   // seeded at the current tick
    uint32_t xchg_gen;        // table-kernel launches so far: generation tag of the peer exchange (never reset)
    int32_t xchg_error;       // a peer did not publish its rows in time
};

// Loads of what the PRECEDING grid wrote (thresholds, clock), issued after grid_dependency_wait: the default cached
// load (ld.global.ca).  The wait makes the predecessor's stores visible to these; ld.global.nc (__ldg) would be wrong
// here (the data is not read-only for the lifetime of this grid) and ld.global.cg would send every look-up of the
// small threshold table to L2.
__device__ __forceinline__ uint32_t ld_ca(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int4 ld_ca(const int4* p) {
    int4 v;
    asm volatile("ld.global.ca.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

struct StepArgs {
    int32_t k;        // sub-step being executed (T/M/A phases)
    int32_t now;      // its minute tick
    int32_t k_inf;    // sub-step whose infection draw is applied (k - 1)
    int32_t now_inf;
    uint32_t mode;
    int32_t weekday;  // 0..6 of `now`
};
__device__ __forceinline__ StepArgs step_args(const DevTables& T, const DevClock* clk, uint32_t mode) {
    const int4 c = ld_ca(reinterpret_cast<const int4*>(clk));           // k, now, weekday, minute_abs (written by the table kernel)
    StepArgs s;
    s.k = c.x; s.now = c.y; s.k_inf = s.k - 1; s.now_inf = s.now - T.step_ticks;
    s.weekday = c.z;
    s.mode = mode | (s.weekday <= 4 ? M_WEEKDAY : 0u);       // base_model.py:309
    return s;
}

// ------------------------------------------------------------------------------------------------ launch overlap
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute
// may start once every CTA of its predecessor has executed launch_dependents (or exited); everything it
// reads that the predecessor writes must come after grid_dependency_wait, which returns when the
// predecessor has completed and its memory is visible.  Both are no-ops for a normally serialised launch.
__device__ __forceinline__ void launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// extra stamps (debug): word w >= 6 of the sub-step's row, min or max over CTAs
__device__ __forceinline__ void trace_point(const DevWork& Wk, int k, int w, bool is_max) {
    if (Wk.trace && threadIdx.x == 0 && k >= 0 && k < Wk.trace_rows) {
        unsigned long long* p = Wk.trace + (size_t)k * ABM_TRACE_WORDS + w;
        if (is_max) atomicMax(p, global_timer_ns()); else atomicMin(p, global_timer_ns());
    }
}
// one thread per CTA stamps the launch's [first start, last end] window
__device__ __forceinline__ void trace_stamp(const DevWork& Wk, int k, int kernel, unsigned long long t_start) {
    if (Wk.trace && threadIdx.x == 0 && k >= 0 && k < Wk.trace_rows) {
        unsigned long long* row = Wk.trace + (size_t)k * ABM_TRACE_WORDS + kernel * 2;
        atomicMin(row, t_start);
        atomicMax(row + 1, global_timer_ns());
    }
}

// Ask L2 to fetch a sector and keep it (evict-last): the sweep issues this for the sectors the event kernel is going to
// touch for a deferred agent, so that kernel's scattered reads / partial writes meet L2 instead of a burst of random
// DRAM transactions.
__device__ __forceinline__ void prefetch_l2_keep(const void* p) {
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}

// system-scope accesses for memory shared with peer GPUs
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 16-byte accesses that always go to memory (peer or local), never to a cached copy
__device__ __forceinline__ void st_volatile_v4(uint4* p, const uint4& v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}


// ------------------------------------------------------------------------------------------------ Philox
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(const DevTables& T, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ T.rk[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ T.rk[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    return U4{c0, c1, c2, c3};
}
// one block per agent (disease-course sites)
__device__ __forceinline__ U4 agent_draws(const DevTables& T, int64_t cid, int32_t step, uint32_t site) {
    return philox4x32_10(T, (uint32_t)cid, (uint32_t)((uint64_t)cid >> 32), (uint32_t)step, site);
}
// one block per group of four consecutive agent ids; word (cid & 3) belongs to agent cid
__device__ __forceinline__ U4 group_draws(const DevTables& T, int64_t cid4, int32_t step, uint32_t site) {
    const uint64_t g = (uint64_t)cid4 >> 2;
    return philox4x32_10(T, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, site);
}
__device__ __forceinline__ uint32_t word_of(const U4& X, int j) { return j == 0 ? X.x : (j == 1 ? X.y : (j == 2 ? X.z : X.w)); }

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

// 31-bit Bernoulli threshold round((1 - exp(-h)) * 2^31) of a hazard h = i / 256 + l / 2^32 (l < 2^24), in
// 32-bit integer arithmetic (oracle.hazard_threshold_fx):
//   1 - exp(-h) = Q[i] + u - Q[i] u,  Q[i] = 1 - exp(-i / 256) (table, 0.32 fixed point),
//   u = 1 - exp(-l) = l - l^2/2 + l^3/6  (remainder < 2^-36).
__device__ __forceinline__ uint32_t hazard_threshold_core(const uint32_t* __restrict__ omexp, uint32_t i, uint32_t l) {
    const uint32_t q = __ldg(omexp + i);
    const uint32_t l2 = (uint32_t)(((unsigned long long)l * l) >> 33);
    const uint32_t l3 = (__umulhi(l2, l) * 171u) >> 9;          // (l^3 / 2 >> 32) / 3; the operand is < 128
    const uint32_t u = l - l2 + l3;
    uint32_t p = q + u - __umulhi(q, u);
    p = p < q ? 0xFFFFFFFFu : p;                                // the sum can reach 2^32 at the end of the table
    return (p >> 1) + (p & 1u);
}
// hazard in units of 2^-32
__device__ __forceinline__ uint32_t hazard_threshold_fx(const uint32_t* __restrict__ omexp, unsigned long long h) {
    if (h >= ((unsigned long long)ABM_OMEXP_N << 24)) return 0x80000000u;     // exp(-32) < 2^-46
    return hazard_threshold_core(omexp, (uint32_t)(h >> 24), (uint32_t)h & 0xFFFFFFu);
}
// hazard in units of 2^-24 (household sums of the qfx table)
__device__ __forceinline__ uint32_t hazard_threshold_q24(const uint32_t* __restrict__ omexp, uint32_t hq) {
    const uint32_t i = hq >> 16;
    return i >= (uint32_t)ABM_OMEXP_N ? 0x80000000u : hazard_threshold_core(omexp, i, (hq & 0xFFFFu) << 8);
}

// sub-step index at which `date < now` first holds: floor(t / step) + 1, by multiply-high with
// inv = floor(2^64 / step) + 1 (exact for every 32-bit t).  Monotone in t.
__device__ __forceinline__ uint32_t fire_index(int32_t t, unsigned long long step_inv) {
    if (t == T_NEVER) return FIRE_NEVER;
    if (t < 0) return 0;
    const uint32_t f = (uint32_t)__umul64hi((unsigned long long)(uint32_t)t, step_inv) + 1u;
    return f < 0xFFFEu ? f : 0xFFFEu;
}

__device__ __forceinline__ int big_slot(const DevTables& T, int64_t cid) {
    int lo = 0, hi = T.n_big;       // big_start[lo] <= cid < big_start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(T.big_start + mid) <= cid) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t loc_of(uint32_t f) { return (f >> F_LOC_SHIFT) & F_LOC_MASK; }
__device__ __forceinline__ uint32_t set_loc(uint32_t f, uint32_t loc) { return (f & ~LOCF) | (loc << F_LOC_SHIFT); }
__device__ __forceinline__ uint32_t status_of(uint32_t f) { return (f >> F_STATUS_SHIFT) & F_STATUS_MASK; }
__device__ __forceinline__ uint32_t cur_of(uint32_t a) { return a & 0xFFFFu; }
// index into the [2][128] rate tables: byte 2 of the flags = age | symptomatic << 7
__device__ __forceinline__ uint32_t rate_index(uint32_t f) { return __byte_perm(f, 0u, 0x4442u); }

// Outside pool (row of the district x status tables) the agent is currently in, or -1.
// Hospital is location -current_district (base_model.py:270-272): district 0 maps onto its own outside
// pool (quirk Q3); every other hospital / the dead location holds no susceptible agent.
__device__ __forceinline__ int pool_of(const DevAgents& G, uint32_t f, uint32_t a, uint32_t slot, int L0) {
    const uint32_t kind = loc_of(f);
    if (kind == LOC_HOMEPOOL) return L0;
    if (kind == LOC_DEST) return (int)cur_of(a);
    if (kind == LOC_POOL) return (int)__ldg(G.econ_pool + slot);
    if (kind == LOC_HOSP) return cur_of(a) == 0u ? 0 : -1;
    return -1;
}
// location an agent that goes about its economic activity ends up in (base_model.py:331-339, 346-352)
__device__ __forceinline__ uint32_t econ_loc(uint32_t f) {
    const uint32_t ek = (f >> F_EKIND_SHIFT) & F_EKIND_MASK;
    static_assert(EKIND_HOUSEHOLD == 0 && LOC_HOME == 0 && EKIND_HOME_DISTRICT == 1 && LOC_HOMEPOOL == 1 && EKIND_POOL == 2 && LOC_POOL == 3, "econ_loc");
    return ek + (ek >> 1);          // branch-free: no jump table in the middle of divergent code
}
// get_active_ids (base_model.py:198-209): (not hospitalised, or recovered) and at a non-negative location.
// The location test is implied: hospital / dead locations only occur with clinical state >= 1 and an
// epidemic state other than RECOVERED (recovery from hospital sends the agent home, base_model.py:246-250).
__device__ __forceinline__ bool is_active(uint32_t f) {
    return (f & CLINF) == 0u || (f & F_EPI_MASK) == EPI_R;
}

// ------------------------------------------------------------------------------------------------ infection event
// Country.set_epidemic_status for one agent (base_model.py:838-1055).  Returns the new flags; writes the
// cold vectors; the high half of `a` gets the sub-step at which the first scheduled event fires (>= k_first).
// tally: shared (int) or global (unsigned long long) array indexed by the enum above.
template <typename TallyT>
__device__ __forceinline__ uint32_t infect_agent(const DevTables& T, uint32_t f, uint32_t& a, Cold& c, int64_t cid, int32_t k_inf,
                                                 int32_t now_inf, uint32_t k_first, TallyT* tally) {
    const U4 A4 = agent_draws(T, cid, k_inf, SITE_COURSE_A);
    const U4 B4 = agent_draws(T, cid, k_inf, SITE_COURSE_B);
    const uint32_t age = (f >> F_AGE_SHIFT) & F_AGE_MASK;
    const uint32_t cur = cur_of(a);
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
        atomicAdd(tally + CUM_SYM, (TallyT)1);
    } else {                                                                                // :1024-1055
        t_o = T_NEVER;
        t_c = now_inf + (int32_t)sample_quantile_u32(T.dwell + 1 * 2 * (ABM_QN + 1), A4.y);
        t_r = t_c + (int32_t)sample_quantile_u32(T.dwell + 3 * 2 * (ABM_QN + 1), B4.z);
        atomicAdd(tally + CUM_ASYM, (TallyT)1);
    }
    atomicAdd(tally + PRE_EXPOSED, (TallyT)1);
    c.t = make_int4(now_inf, t_c, t_o, t_r);                                                // :851
    c.u.y = (int32_t)((loc_of(f) << 16) | cur);                                             // :846-847
    f &= ~(F_EPI_MASK | F_SYMPT | F_ASYMPT | F_OUT_H | F_OUT_C | F_OUT_D | F_ONSET);
    f |= EPI_I | (sym ? F_SYMPT : F_ASYMPT) | out;                                          // :850
    // earliest scheduled event: the one with the smallest date (fire_index is monotone)
    int32_t first = t_c < t_o ? t_c : t_o;
    first = first < t_r ? first : t_r;
    uint32_t fi = fire_index(first, T.step_inv);
    fi = fi > k_first ? fi : k_first;
    a = (a & 0xFFFFu) | (fi << 16);
    return f;
}

// ------------------------------------------------------------------------------------------------ transitions
// update_agent_states (base_model.py:226-272) + contagious promotion (base_model.py:54-58) for one agent
// whose next-event index is due.  Cumulative date counters are tallied the first time a date is passed;
// the "current" counters of the day log move by the change of epidemic / clinical state.
// What an agent contributes to the accumulated tables: used to emit the difference a transition makes.
struct Contribution { int pool; bool cont; bool big_home; };
__device__ __forceinline__ Contribution contribution_of(const DevAgents& G, uint32_t f, uint32_t a, uint32_t slot, int L0) {
    Contribution c;
    c.pool = pool_of(G, f, a, slot, L0);
    c.cont = (f & F_EPI_MASK) == EPI_C;
    c.big_home = c.cont && (f & (LOCF | F_BIGHH)) == F_BIGHH;
    return c;
}

struct DeltaSink {          // where incremental sub-steps send their table changes (nullptr = full recount follows)
    unsigned long long* acc_r; unsigned long long* acc_n; unsigned long long* hbig; int L0; int64_t cid;
};

__device__ __forceinline__ uint32_t transition_agent(const DevTables& T, const DevAgents& G, uint32_t f, uint32_t& a, uint32_t slot,
                                                     const Cold& c, int32_t now, int* s_tally, const DeltaSink* sink = nullptr) {
    const uint32_t f_before = f;
    const int32_t t_c = c.t.y, t_o = c.t.z, t_r = c.t.w;
    const int32_t t_h = (f & F_OUT_H) ? t_o + 5 * DAY : T_NEVER;    // params.py:165
    const int32_t t_cs = (f & F_OUT_C) ? t_o + 13 * DAY : T_NEVER;  // hospital end = critical start (base_model.py:953)
    const int32_t t_d = (f & F_OUT_D) ? t_o + 21 * DAY : T_NEVER;   // params.py:177 (== end of critical care when critical)
    const uint32_t epi0 = f & F_EPI_MASK, clin0 = (f >> F_CLIN_SHIFT) & F_CLIN_MASK;
    uint32_t epi = epi0, clin = clin0, loc = loc_of(f);
    if (t_c < now && epi == EPI_I) atomicAdd(s_tally + CUM_CONTAGIOUS, 1);
    if (t_d < now && epi < EPI_D) {                                  // :230-237
        epi = EPI_D; loc = LOC_DEAD; clin = 3;
        atomicAdd(s_tally + CUM_DIED, 1);
    }
    if (t_r < now && epi < EPI_R) {                                  // :240-254
        epi = EPI_R;
        if (t_h < now) loc = LOC_HOME;
        clin = 3;
        atomicAdd(s_tally + CUM_RECOVERED, 1);
    }
    if (t_cs < now && clin < 2) { clin = 2; atomicAdd(s_tally + CUM_CRITICAL, 1); }      // :257-261
    if (t_h < now && clin < 1) { clin = 1; loc = LOC_HOSP; atomicAdd(s_tally + CUM_HOSPITALIZED, 1); }  // :263-272
    if (t_o < now) f |= F_ONSET;                                     // :212-219 (symptom onset passed)
    if (t_c < now && epi < EPI_C) epi = EPI_C;                       // :54-58
    if (epi != epi0) {
        if (epi0 == EPI_I) atomicAdd(s_tally + POST_EXPOSED, -1);
        if (epi0 == EPI_C) atomicAdd(s_tally + POST_CONTAGIOUS, -1);
        if (epi == EPI_C) atomicAdd(s_tally + POST_CONTAGIOUS, 1);
    }
    if (clin != clin0) {
        if (clin0 == 1) atomicAdd(s_tally + POST_HOSPITALIZED, -1);
        if (clin0 == 2) atomicAdd(s_tally + POST_CRITICAL, -1);
        if (clin == 1) atomicAdd(s_tally + POST_HOSPITALIZED, 1);
        if (clin == 2) atomicAdd(s_tally + POST_CRITICAL, 1);
    }
    f = (f & ~(F_EPI_MASK | CLINF | LOCF)) | epi | (clin << F_CLIN_SHIFT) | (loc << F_LOC_SHIFT);
    if (sink && ((f ^ f_before) & (F_EPI_MASK | LOCF))) {
        const Contribution c0 = contribution_of(G, f_before, a, slot, sink->L0), c1 = contribution_of(G, f, a, slot, sink->L0);
        const uint32_t st = (f >> F_STATUS_SHIFT) & F_STATUS_MASK;
        const unsigned long long rate = __ldg(T.rfx + __byte_perm(f, 0u, 0x4442u));
        if (c0.pool != c1.pool) {
            if (c0.pool >= 0) atomicAdd(sink->acc_n + c0.pool * T.A + st, ~0ull);           // - 1
            if (c1.pool >= 0) atomicAdd(sink->acc_n + c1.pool * T.A + st, 1ull);
        }
        if (c0.pool != c1.pool || c0.cont != c1.cont) {
            if (c0.pool >= 0 && c0.cont) atomicAdd(sink->acc_r + c0.pool * T.A + st, 0ull - rate);
            if (c1.pool >= 0 && c1.cont) atomicAdd(sink->acc_r + c1.pool * T.A + st, rate);
        }
        if (c0.big_home != c1.big_home) {
            const unsigned long long qv = __ldg(T.qfx + (T.hw_rows > 1 ? (size_t)sink->L0 * 256 : 0) + __byte_perm(f, 0u, 0x4442u));
            atomicAdd(sink->hbig + big_slot(T, sink->cid), c1.big_home ? qv : 0ull - qv);
        }
    }
    // next event = smallest date not yet passed (fire_index is monotone, so one evaluation suffices)
    int32_t nxt = T_NEVER;
    const int32_t ts[6] = {t_c, t_o, t_h, t_cs, t_d, t_r};
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (ts[i] >= now && ts[i] < nxt) nxt = ts[i];
    a = (a & 0xFFFFu) | (fire_index(nxt, T.step_inv) << 16);
    return f;
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
// 8-ary search: the pivots of a level are independent loads, so a row of 60 columns costs two memory round trips
// (500 columns: three) instead of the six to nine dependent ones of a binary search.
__device__ __forceinline__ uint32_t od_destination(const uint32_t* __restrict__ row, int D, uint32_t u) {
    if (u >= __ldg(row + D - 1)) return 0u;
    int lo = 0, len = D;                    // invariant: the answer is in [lo, lo + len) and row[lo + len - 1] > u
    while (len > 8) {
        const int S = (len + 7) >> 3;
        int cnt = 0;
#pragma unroll
        for (int i = 1; i <= 7; ++i) {
            const int idx = min(lo + i * S - 1, lo + len - 1);
            cnt += u >= __ldg(row + idx) ? 1 : 0;       // pivots <= u form a prefix; a clamped pivot is > u
        }
        lo += cnt * S;
        len = min(S, len - cnt * S);
    }
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < len) cnt += u >= __ldg(row + lo + i) ? 1 : 0;
    return (uint32_t)(lo + cnt);
}

// ------------------------------------------------------------------------------------------------ sweep + event kernels
// One sub-step is three launches:
//   abm_sweep_kernel    every agent slot, one warp per sub-tile: the DENSE work -- household hazards, the infection draw
//                       I(k-1), detection of due transitions, the movement of everybody who simply goes to work / goes
//                       home, and the pressure / head-count accumulation A(k).  It never touches a cold vector and has
//                       no rare code path: an agent that needs one -- newly infected, a scheduled event due, away on a
//                       trip at a movement hour, leaving the district -- is DEFERRED: its slot is appended to an event
//                       queue and the sweep leaves it alone (no movement, no accumulation, words written back as read).
//   abm_event_kernel    one THREAD per deferred agent (full warps instead of the 1-3 active lanes these bodies had inside
//                       the sweep): infection event, due transitions, its movement, its table contribution -- the
//                       reference's order of operations for that one agent -- and the store of its hot words.
//   abm_hazard_table_kernel   tables -> thresholds, day log, clock (unchanged).
// All table sums are integers, every draw is keyed by (agent, sub-step, site), so splitting the agents between the two
// kernels cannot change a result.
constexpr uint32_t RUNTIME_MODE = 0xFFFFFFFFu;
// Event kernel shape: 128 threads x 8 CTAs per SM = 64 registers per thread, no spills, and all of the <= 1024 queues
// resident in one wave (1184 CTA slots).  The 32-register shape of the first split (256 x 8) spilled 236 bytes per thread;
// its local-memory traffic (3.7 M sectors at 08:00) cost 14 % of the kernel at the peak, 4.6 % of a simulated day
// (profiles/r02b_variants_*.txt: 256 x 4 at 64 registers is as fast at the peak and slower on quiet nights).
#ifndef ABM_EVENT_THREADS
#define ABM_EVENT_THREADS 128
#endif
#ifndef ABM_EVENT_MIN_BLOCKS
#define ABM_EVENT_MIN_BLOCKS (1024 / ABM_EVENT_THREADS)
#endif
constexpr int EVENT_THREADS = ABM_EVENT_THREADS;
constexpr uint32_t EV_LEAVER = 1u << 31;  // entry word 3: the sweep decided "mover, leaves the district"; low 31 bits = the OD draw

// Funnel loop (rare, irregular locations only): iterate a body over the agents whose bit is set in the per-lane `mask`.
// Inside, J is the agent's index 0..3 within the lane and fj / aj are its flags / aux words.
#define ABM_FUNNEL_RO(mask, ...)                                                     \
    while (__any_sync(FULL, (mask) != 0u)) {                                         \
        if (mask) {                                                                  \
            const int J = __ffs(mask) - 1;                                           \
            (mask) &= (mask) - 1u;                                                   \
            const uint32_t fj = J == 0 ? f0 : (J == 1 ? f1 : (J == 2 ? f2 : f3));    \
            const uint32_t aj = J == 0 ? a0 : (J == 1 ? a1 : (J == 2 ? a2 : a3));    \
            __VA_ARGS__                                                              \
        }                                                                            \
    }

// Append the agents whose bit is set in `ev` (per-lane 4-bit mask; bit j = slot0 + j) to the event queue of this CTA's
// queue group.  Entry = (slot | newly-infected << 31, flags, aux, 0): the hot words travel with the entry, so the event
// kernel reads its queue coalesced and touches only the agent's cold sector.  One atomic per LANE that has anything to
// append (few lanes do; the queues spread the atomics over up to 1024 addresses).
#ifndef ABM_PUSH_PREFETCH
#define ABM_PUSH_PREFETCH 15     // bit 0 / 1: cold sector / hot words of a generic event; bit 2 / 3: the same for a decided leaver
#endif
#define ABM_PUSH_EVENTS(ev, newly, lvm, XO)                                                          \
    if (ev) {                                                                                        \
        const uint32_t q = cta & Wk.q_mask;                                                          \
        uint4* seg = Wk.q_entries + (size_t)q * Wk.q_cap;                                            \
        uint32_t m = (ev) & ~(lvm);                 /* generic events: from the front */             \
        if (m) {                                                                                     \
            uint4* dst = seg + atomicAdd(Wk.q_count + 2u * q, (uint32_t)__popc(m));                  \
            do {                                                                                     \
                const uint32_t j = __ffs(m) - 1u;                                                    \
                m &= m - 1u;                                                                         \
                const uint32_t fj = j == 0 ? f0 : (j == 1 ? f1 : (j == 2 ? f2 : f3));    /* untouched */ \
                const uint32_t aj = j == 0 ? a0 : (j == 1 ? a1 : (j == 2 ? a2 : a3));                \
                *dst++ = make_uint4((slot0 + j) | ((((newly) >> j) & 1u) << 31), fj, aj, 0u);        \
                if (ABM_PUSH_PREFETCH & 1) prefetch_l2_keep(G.cold + (size_t)(slot0 + j) * ABM_COLD_WORDS); \
                if (ABM_PUSH_PREFETCH & 2) prefetch_l2_keep(G.flags + slot0 + j);                    \
                if (ABM_PUSH_PREFETCH & 2) prefetch_l2_keep(G.aux + slot0 + j);                      \
            } while (m);                                                                             \
        }                                                                                            \
        m = (ev) & (lvm);                           /* decided leavers: from the back */             \
        if (m) {                                                                                     \
            uint4* dst = seg + (Wk.q_cap - 1u) - atomicAdd(Wk.q_count + 2u * q + 1u, (uint32_t)__popc(m)); \
            do {                                                                                     \
                const uint32_t j = __ffs(m) - 1u;                                                    \
                m &= m - 1u;                                                                         \
                const uint32_t fj = j == 0 ? f0 : (j == 1 ? f1 : (j == 2 ? f2 : f3));                \
                const uint32_t aj = j == 0 ? a0 : (j == 1 ? a1 : (j == 2 ? a2 : a3));                \
                *dst-- = make_uint4(slot0 + j, fj, aj, EV_LEAVER | (word_of(XO, (int)j) >> 1));      \
                if ((ABM_PUSH_PREFETCH & 4) && (fj & F_EPI_MASK) != EPI_S) prefetch_l2_keep(G.cold + (size_t)(slot0 + j) * ABM_COLD_WORDS); \
                if (ABM_PUSH_PREFETCH & 8) prefetch_l2_keep(G.flags + slot0 + j);                    \
                if (ABM_PUSH_PREFETCH & 8) prefetch_l2_keep(G.aux + slot0 + j);                      \
            } while (m);                                                                             \
        }                                                                                            \
    }

// resident CTAs per SM the register allocation aims at: 8 (32 registers, full occupancy) for the read-only night / noon
// sweeps, 5 (48 registers) for the movement sweeps
// (measured: 6 instead of 8 CTAs on the night sweeps, and 4 or 6 instead of 5 on the movement sweeps, change nothing
// beyond noise)
#ifndef ABM_SWEEP_CTAS_INCR
#define ABM_SWEEP_CTAS_INCR (2048 / THREADS)
#endif
#ifndef ABM_SWEEP_CTAS_MOVE
#define ABM_SWEEP_CTAS_MOVE (1280 / THREADS)
#endif
constexpr int sweep_min_blocks(uint32_t mode) { return (mode != RUNTIME_MODE && (mode & M_INCR)) ? ABM_SWEEP_CTAS_INCR : ABM_SWEEP_CTAS_MOVE; }

// TRACE: the timeline stamps of abm_set_trace are compiled in only for the instantiation that is launched while a trace
// buffer is attached.  In the production instantiation they cost the night sweep 24 instructions per warp and, at its
// 32-register budget, two spilled registers: 3.7 % of the mean sweep launch (profiles/r02b_variants_*.txt).
template <uint32_t MODE, bool TRACE>
__global__ void __launch_bounds__(THREADS, sweep_min_blocks(MODE))
abm_sweep_kernel(const DevTables T, const DevAgents G, const DevWork Wk, const DevClock* clk, const uint32_t mode_rt) {
    __shared__ uint32_t s_hz[WARPS][SUB];             // per-warp household hazard sums, then thresholds
    __shared__ unsigned int s_acc[WARPS][2 * ABM_MAX_STATUS];   // per-warp home-pool pressure sums / head counts

    const uint32_t mode_static = MODE == RUNTIME_MODE ? (mode_rt & 0xFFFFu) : MODE;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned long long t_start = TRACE ? global_timer_ns() : 0ull;
    // The event kernel may be scheduled as soon as SM resources free up (i.e. while this grid drains): it waits for
    // this grid's completion before it touches anything.
    launch_dependents();
    // One CTA per tile.  (A persistent grid walking the tiles was measured: 40 instead of 32 registers on the night
    // kernels, 62 instead of 48 at go-out, no faster on a quiet night and 10 % slower at the peak.)
    const uint32_t cta = blockIdx.x;
    const uint32_t sub = cta * WARPS + wid;
    const unsigned long long info = __ldg(T.sub_info + sub);
    const uint32_t L0 = (uint32_t)(info >> 48);                              // home district of every agent here
    const uint32_t slot0 = sub * SUB + lane * PER_LANE;
    const uint32_t A = (uint32_t)T.A;

    // streaming loads / stores (evict-first): the hot words are touched once per launch and must not push the small
    // tables every warp re-reads out of L1
    const uint4 f4 = __ldcs(reinterpret_cast<const uint4*>(G.flags + slot0));
    const uint4 a4 = __ldcs(reinterpret_cast<const uint4*>(G.aux + slot0));
    // household index of each slot within the sub-tile (static, one byte per slot): issued with the hot
    // words so the household phase does not wait for a dependent DRAM round trip
    uint32_t ix = 0;
    if (mode_static & M_INFECT) ix = __ldcs(G.hh_index + (slot0 >> 2));
    // movement classes (go-out only): requested now, used after the infection phase
    uint2 mc = make_uint2(0u, 0u);
    if (mode_static & M_GO_OUT) mc = __ldcs(reinterpret_cast<const uint2*>(G.move_class + slot0));

    unsigned int* acc_w = s_acc[wid];
    if ((mode_static & M_DENSE_A) && !(mode_static & M_INCR)) {
        if (lane < 2 * ABM_MAX_STATUS) acc_w[lane] = 0u;
        __syncwarp();
    }

    uint32_t f0 = f4.x, f1 = f4.y, f2 = f4.z, f3 = f4.w;
    uint32_t a0 = a4.x, a1 = a4.y, a2 = a4.z, a3 = a4.w;
    // return ticks of the travellers among the four agents (movement sub-steps only): requested now so the DRAM round
    // trip overlaps the infection phase instead of stalling the movement phase
    int32_t tr0 = T_NEVER, tr1 = T_NEVER, tr2 = T_NEVER, tr3 = T_NEVER;
    if (mode_static & (M_GO_OUT | M_GO_HOME)) {
        if (f0 & F_AWAY) tr0 = __ldcs(G.cold + (size_t)(slot0 + 0u) * ABM_COLD_WORDS + ABM_COLD_T_RETURN);
        if (f1 & F_AWAY) tr1 = __ldcs(G.cold + (size_t)(slot0 + 1u) * ABM_COLD_WORDS + ABM_COLD_T_RETURN);
        if (f2 & F_AWAY) tr2 = __ldcs(G.cold + (size_t)(slot0 + 2u) * ABM_COLD_WORDS + ABM_COLD_T_RETURN);
        if (f3 & F_AWAY) tr3 = __ldcs(G.cold + (size_t)(slot0 + 3u) * ABM_COLD_WORDS + ABM_COLD_T_RETURN);
    }
    uint32_t* hz = s_hz[wid];

    // ================================================================ I(k-1): infection draws
    uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;          // 31-bit infection thresholds of the four agents
    if (mode_static & M_INFECT) {
        // ---- household term: 1 - prod(1 - 3 r_i) over contagious members present (base_model.py:156-175, 184-187)
        constexpr uint32_t CH_MASK = F_EPI_MASK | LOCF | F_BIGHH, CH_VAL = EPI_C | (LOC_HOME << F_LOC_SHIFT);
        constexpr uint32_t SH_VAL = EPI_S | (LOC_HOME << F_LOC_SHIFT);
        const bool c0 = (f0 & CH_MASK) == CH_VAL, c1 = (f1 & CH_MASK) == CH_VAL, c2 = (f2 & CH_MASK) == CH_VAL,
                   c3 = (f3 & CH_MASK) == CH_VAL;
        const bool h0 = (f0 & CH_MASK) == SH_VAL, h1 = (f1 & CH_MASK) == SH_VAL, h2 = (f2 & CH_MASK) == SH_VAL,
                   h3 = (f3 & CH_MASK) == SH_VAL;
        // somebody contagious at home AND somebody susceptible at home in this sub-tile, else no household can infect
        if (__any_sync(FULL, c0 | c1 | c2 | c3) && __any_sync(FULL, h0 | h1 | h2 | h3)) {
            const uint32_t i0 = ix & 0xFFu, i1 = (ix >> 8) & 0xFFu, i2 = (ix >> 16) & 0xFFu, i3 = ix >> 24;
            const uint32_t n_hh = (__shfl_sync(FULL, i3, 31) + 1u) & 0xFFu;   // 255 = headless run of a big household
            *reinterpret_cast<uint4*>(hz + lane * 4) = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            const uint32_t qoff = T.hw_rows > 1 ? L0 * 256u : 0u;
            if (c0) atomicAdd(hz + i0, __ldg(T.qfx + (qoff + rate_index(f0))));
            if (c1) atomicAdd(hz + i1, __ldg(T.qfx + (qoff + rate_index(f1))));
            if (c2) atomicAdd(hz + i2, __ldg(T.qfx + (qoff + rate_index(f2))));
            if (c3) atomicAdd(hz + i3, __ldg(T.qfx + (qoff + rate_index(f3))));
            __syncwarp();
#pragma unroll 1
            for (uint32_t h = lane; h < n_hh; h += 32) {     // one threshold per household with somebody contagious
                const uint32_t hq = hz[h];
                if (hq) hz[h] = hazard_threshold_q24(T.omexp, hq);
            }
            __syncwarp();
            if (h0) t0 = hz[i0];
            if (h1) t1 = hz[i1];
            if (h2) t2 = hz[i2];
            if (h3) t3 = hz[i3];
        }
    }

    // Everything above depends only on the agent store, which the previous event kernel finished writing
    // before the table kernel started.  From here on the outputs of the table kernel are needed: thresholds,
    // cleared accumulators, the advanced clock.
    grid_dependency_wait();
    const StepArgs s = step_args(T, clk, mode_static);
    const uint32_t mode = s.mode;

    uint32_t newly = 0;
    if (mode & M_INFECT) {
        // ---- outside pools and big households: susceptible and not (at home in a small household).
        // Home-district pool and trip destinations are one table look-up (predicated, all four agents);
        // schools, the hospital of district 0 and big households are rare and go through the funnel.
        constexpr uint32_t IN_DEST = LOC_DEST << F_LOC_SHIFT;
        uint32_t odd = 0;
        if (mode & M_DENSE_I) {
            // daytime: almost everybody is in the outside pool of the home district -> row in registers
            const uint32_t row_thr = lane < (int)A ? ld_ca(Wk.thr_out + L0 * A + lane) : 0u;
            const uint32_t r0 = __shfl_sync(FULL, row_thr, status_of(f0)), r1 = __shfl_sync(FULL, row_thr, status_of(f1));
            const uint32_t r2 = __shfl_sync(FULL, row_thr, status_of(f2)), r3 = __shfl_sync(FULL, row_thr, status_of(f3));
#define ABM_OUT_THR(fj, aj, rj, tj, bit)                                                             \
            {                                                                                        \
                const uint32_t lf = (fj) & LOCF;                                                     \
                const bool so = ((fj) & F_EPI_MASK) == EPI_S && ((fj) & (LOCF | F_BIGHH)) != 0u;     \
                const bool dst = so && lf == IN_DEST;                                                \
                if (so && lf == IN_HOMEPOOL) tj = rj;                                                \
                if (dst) tj = ld_ca(Wk.thr_out + (cur_of(aj) * A + status_of(fj)));                  \
                odd |= (so && lf != IN_HOMEPOOL && !dst) ? (bit) : 0u;                               \
            }
            ABM_OUT_THR(f0, a0, r0, t0, 1u)
            ABM_OUT_THR(f1, a1, r1, t1, 2u)
            ABM_OUT_THR(f2, a2, r2, t2, 4u)
            ABM_OUT_THR(f3, a3, r3, t3, 8u)
#undef ABM_OUT_THR
        } else {
            // night: only travellers and the few agents of quirk Q4 are in an outside pool.  Straight-line,
            // predicated look-ups (the four loads are in flight together; no divergent loop to resolve).
#define ABM_OUT_THR(fj, aj, tj, bit)                                                                 \
            {                                                                                        \
                const uint32_t lf = (fj) & LOCF;                                                     \
                const bool so = ((fj) & F_EPI_MASK) == EPI_S && ((fj) & (LOCF | F_BIGHH)) != 0u;     \
                const bool reg = so && (lf - IN_HOMEPOOL) <= (IN_DEST - IN_HOMEPOOL);                \
                if (reg) tj = ld_ca(Wk.thr_out + ((lf == IN_DEST ? cur_of(aj) : L0) * A + status_of(fj))); \
                odd |= (so && !reg) ? (bit) : 0u;                                                    \
            }
            ABM_OUT_THR(f0, a0, t0, 1u)
            ABM_OUT_THR(f1, a1, t1, 2u)
            ABM_OUT_THR(f2, a2, t2, 4u)
            ABM_OUT_THR(f3, a3, t3, 8u)
#undef ABM_OUT_THR
        }
        ABM_FUNNEL_RO(odd, {
            uint32_t thr = 0;
            if (loc_of(fj) == LOC_HOME) {        // big household: hazard accumulated by the previous launch
                const int64_t cid = (int64_t)(info & 0xFFFFFFFFFFFFull) + lane * PER_LANE + J;
                thr = hazard_threshold_fx(T.omexp, __ldcg(((s.k & 1) ? Wk.hbig2[0] : Wk.hbig2[1]) + big_slot(T, cid)) << 8);
            } else {
                const int pool = pool_of(G, fj, aj, slot0 + J, (int)L0);
                if (pool >= 0) thr = ld_ca(Wk.thr_out + ((uint32_t)pool * A + status_of(fj)));
            }
            if (J == 0) t0 = thr; else if (J == 1) t1 = thr; else if (J == 2) t2 = thr; else t3 = thr;
        })
        // ---- the draw: one Philox block for the four agents of the lane
        if (t0 | t1 | t2 | t3) {
            const int64_t cid0 = (int64_t)(info & 0xFFFFFFFFFFFFull) + lane * PER_LANE;
            const U4 X = group_draws(T, cid0, s.k_inf, SITE_INFECT);
            newly = ((X.x >> 1) < t0 ? 1u : 0u) | ((X.y >> 1) < t1 ? 2u : 0u) | ((X.z >> 1) < t2 ? 4u : 0u) |
                    ((X.w >> 1) < t3 ? 8u : 0u);
        }
    }

    uint32_t defer = newly;           // agents handed to the event kernel
    uint32_t leaver = 0;              // ... of which: nothing but "mover, leaves the home district" (decided here)
    U4 XO = U4{0u, 0u, 0u, 0u};       // OD draws of the four agents (go-out only)
    if (mode & M_STEP) {
        // ============================================================ T(k): due transitions -> event kernel
        {
            const uint32_t lim = ((uint32_t)s.k << 16) | 0xFFFFu;      // fire index (high half of aux) <= k
            defer |= (a0 <= lim ? 1u : 0u) | (a1 <= lim ? 2u : 0u) | (a2 <= lim ? 4u : 0u) | (a3 <= lim ? 8u : 0u);
        }
        bool dirty = false, adirty = false;

        // ============================================================ M(k): movement (base_model.py:274-441)
        if (mode & M_GO_HOME) {                                                         // :274-302
#define ABM_GO_HOME(fj, aj, trj, bit)                                                                \
            {                                                                                        \
                const bool act = !(defer & (bit)) && is_active(fj);                                  \
                const bool away = ((fj) & F_AWAY) != 0u;                                             \
                const bool back = act && away && (trj) < s.now;                         /* :285-293 */  \
                const bool home = back || (act && !away);                               /* LOC_HOME == 0 */ \
                const uint32_t nf = home ? ((fj) & ~(LOCF | (back ? F_AWAY : 0u))) : (fj);           \
                aj = back ? (((aj) & 0xFFFF0000u) | L0) : (aj);                                      \
                dirty |= nf != (fj); adirty |= back;                                                 \
                fj = nf;                                                                             \
            }
            ABM_GO_HOME(f0, a0, tr0, 1u)
            ABM_GO_HOME(f1, a1, tr1, 2u)
            ABM_GO_HOME(f2, a2, tr2, 4u)
            ABM_GO_HOME(f3, a3, tr3, 8u)
#undef ABM_GO_HOME
        }
        if (mode & M_GO_OUT) {                                                          // :304-340
            const uint32_t thr_col = (mode & M_WEEKDAY) ? 0u : 2u;
            const uint32_t od_base = ((uint32_t)s.weekday * (uint32_t)T.D + L0) * (uint32_t)T.D + L0;
            const uint32_t od_lo = L0 > 0 ? __ldg(T.od_thr + (od_base - 1u)) : 0u, od_hi = __ldg(T.od_thr + od_base);
            const uint32_t od_span = od_hi - od_lo;       // stay-home interval [od_lo, od_hi): u - od_lo < od_span (rows are non-decreasing)
            const int64_t cid0 = (int64_t)(info & 0xFFFFFFFFFFFFull) + lane * PER_LANE;
            const U4 XM = group_draws(T, cid0, s.k, SITE_MOVER);
            uint32_t mover = 0, want_od = 0;
            const bool mild_on = T.mild_active != 0;
#define ABM_MOVER(fj, cls, xm, bit)                                                                  \
            {                                                                                        \
                const bool act = !(defer & (bit)) && is_active(fj);                                  \
                const bool mild = mild_on && ((fj) & F_ONSET) && ((fj) & F_EPI_MASK) < EPI_R;         /* :212-219, :316 */ \
                const uint32_t thr = __ldg(T.move_thr + (thr_col + 4u * (cls) + (mild ? 1u : 0u)));   \
                const bool mv = act && ((xm) >> 1) < thr;                               /* :317-318 a mover */ \
                mover |= mv ? (bit) : 0u;                                                            \
                want_od |= (mv && ((fj) & F_DMOVER)) ? (bit) : 0u;                                   \
            }
            ABM_MOVER(f0, mc.x & 0xFFFFu, XM.x, 1u)
            ABM_MOVER(f1, mc.x >> 16, XM.y, 2u)
            ABM_MOVER(f2, mc.y & 0xFFFFu, XM.z, 4u)
            ABM_MOVER(f3, mc.y >> 16, XM.w, 8u)
#undef ABM_MOVER
            // District movers whose OD draw falls outside the stay-home interval of their row leave the district: the
            // destination search and the length of stay are the event kernel's.  Travellers are back in the home district
            // when the stay is over (:342-360), then like everybody else.  All other movers go about their economic activity.
            if (want_od) XO = group_draws(T, cid0, s.k, SITE_OD);
#define ABM_GO_OUT(fj, aj, trj, xo, bit)                                                             \
            {                                                                                        \
                const bool mv = (mover & (bit)) != 0u;                                               \
                const bool away = ((fj) & F_AWAY) != 0u;                                             \
                const bool q4 = away && !((trj) < s.now);   /* away on paper, not due back: quirk Q4 */ \
                const bool leaves = (want_od & (bit)) && (((xo) >> 1) - od_lo) >= od_span;           \
                const bool lv = mv && !q4 && leaves;        /* incl. a traveller who is due back and leaves again */ \
                const bool upd = mv && !lv;                                                          \
                const bool ret = upd && away && !q4;        /* traveller back in the home district (:342-360) */ \
                defer |= lv ? (bit) : 0u;                                                            \
                leaver |= (lv && !away) ? (bit) : 0u;       /* at home, untouched by I / T: fully decided */ \
                const uint32_t nf = upd ? (set_loc(fj, econ_loc(fj)) & ~(ret ? F_AWAY : 0u)) : (fj); /* :331-339, :346-352 */ \
                aj = ret ? (((aj) & 0xFFFF0000u) | L0) : (aj);                                       \
                dirty |= nf != (fj); adirty |= ret;                                                  \
                fj = nf;                                                                             \
            }
            ABM_GO_OUT(f0, a0, tr0, XO.x, 1u)
            ABM_GO_OUT(f1, a1, tr1, XO.y, 2u)
            ABM_GO_OUT(f2, a2, tr2, XO.z, 4u)
            ABM_GO_OUT(f3, a3, tr3, XO.w, 8u)
#undef ABM_GO_OUT
        }

        // ============================================================ A(k): pressure / headcount tables
        if (!(mode & M_INCR)) {
            unsigned long long* const acc_k = (s.k & 1) ? Wk.acc2[1] : Wk.acc2[0];
            unsigned long long* acc_r = acc_k + (size_t)(cta & Wk.rep_mask) * (2u * T.P * A);
            unsigned long long* acc_n = acc_r + (size_t)T.P * A;
            // Home-district pool: per-warp shared accumulators (daytime) or global adds (night).  Trip
            // destinations: predicated global adds.  Schools, the hospital of district 0 (quirk Q3) and
            // contagious members of big households at home are rare and go through the funnel.  Deferred agents
            // are counted by the event kernel, after their events.
            constexpr uint32_t IN_DEST = LOC_DEST << F_LOC_SHIFT, IN_POOL = LOC_POOL << F_LOC_SHIFT, IN_HOSP = LOC_HOSP << F_LOC_SHIFT;
            constexpr uint32_t BH_MASK = F_EPI_MASK | LOCF | F_BIGHH, BH_VAL = EPI_C | F_BIGHH;
            uint32_t work = 0;
            if (mode & M_DENSE_A) {
                // per-lane head counts of the home pool in 4-bit fields (status 0..7 | 8..11), widened to bytes for
                // three warp reductions; everybody else who is in a pool (trip destination, school, hospital of
                // district 0) goes through the funnel
                uint32_t cw_lo = 0, cw_hi = 0;
#define ABM_COUNT(fj, bit)                                                                           \
                {                                                                                    \
                    const uint32_t lf = (fj) & LOCF;                                                 \
                    const bool cnt = !(defer & (bit));                                               \
                    const bool hp = cnt && lf == IN_HOMEPOOL;                                        \
                    const uint32_t sh = ((fj) >> (F_STATUS_SHIFT - 2)) & 0x3Cu;         /* 4 * status */ \
                    const uint32_t one = hp ? (1u << (sh & 31u)) : 0u;                               \
                    cw_lo += (sh & 32u) ? 0u : one;                                                  \
                    cw_hi += (sh & 32u) ? one : 0u;                                                  \
                    if (hp && ((fj) & F_EPI_MASK) == EPI_C) atomicAdd(acc_w + (sh >> 2), __ldg(T.rfx + rate_index(fj))); \
                    work |= (cnt && (lf == IN_DEST || lf == IN_POOL || lf == IN_HOSP || ((fj) & BH_MASK) == BH_VAL)) ? (bit) : 0u; \
                }
                ABM_COUNT(f0, 1u)
                ABM_COUNT(f1, 2u)
                ABM_COUNT(f2, 4u)
                ABM_COUNT(f3, 8u)
#undef ABM_COUNT
                const uint32_t ev = __reduce_add_sync(FULL, cw_lo & 0x0F0F0F0Fu);            // statuses 0,2,4,6
                const uint32_t od = __reduce_add_sync(FULL, (cw_lo >> 4) & 0x0F0F0F0Fu);     // statuses 1,3,5,7
                const uint32_t hi = __reduce_add_sync(FULL, (cw_hi & 0x0F0Fu) | (((cw_hi >> 4) & 0x0F0Fu) << 16));   // 8,10 | 9,11
                if (lane < (int)A) {
                    const uint32_t w = lane < 8 ? ((lane & 1) ? od : ev) : hi;
                    const uint32_t sh = lane < 8 ? 8u * (lane >> 1) : 8u * (((lane - 8) >> 1) + 2 * ((lane - 8) & 1));
                    acc_w[ABM_MAX_STATUS + lane] += (w >> sh) & 0xFFu;       // this warp's own accumulator row
                }
            } else {
                // night: hardly anybody is in an outside pool
                uint32_t pooled = 0;
#define ABM_COUNT(fj, bit)                                                                           \
                {                                                                                    \
                    const uint32_t lf = (fj) & LOCF;                                                 \
                    const bool cnt = !(defer & (bit));                                               \
                    const bool reg = cnt && (lf - IN_HOMEPOOL) <= (IN_DEST - IN_HOMEPOOL);           \
                    pooled |= reg ? (bit) : 0u;                                                      \
                    work |= (cnt && (lf == IN_POOL || lf == IN_HOSP || ((fj) & BH_MASK) == BH_VAL)) ? (bit) : 0u; \
                }
                ABM_COUNT(f0, 1u)
                ABM_COUNT(f1, 2u)
                ABM_COUNT(f2, 4u)
                ABM_COUNT(f3, 8u)
#undef ABM_COUNT
#define ABM_ADD_POOLED(fj, aj, bit)                                                                  \
                if (pooled & (bit)) {                                                                \
                    const uint32_t e = (((fj) & LOCF) == IN_DEST ? cur_of(aj) : L0) * A + status_of(fj); \
                    atomicAdd(acc_n + e, 1ull);                                                      \
                    if (((fj) & F_EPI_MASK) == EPI_C) atomicAdd(acc_r + e, (unsigned long long)__ldg(T.rfx + rate_index(fj))); \
                }
                ABM_ADD_POOLED(f0, a0, 1u)
                ABM_ADD_POOLED(f1, a1, 2u)
                ABM_ADD_POOLED(f2, a2, 4u)
                ABM_ADD_POOLED(f3, a3, 8u)
#undef ABM_ADD_POOLED
            }
            ABM_FUNNEL_RO(work, {
                const int pool = pool_of(G, fj, aj, slot0 + J, (int)L0);
                const bool cont = (fj & F_EPI_MASK) == EPI_C;
                if (pool >= 0) {
                    const uint32_t e = (uint32_t)pool * A + status_of(fj);
                    atomicAdd(acc_n + e, 1ull);
                    if (cont) atomicAdd(acc_r + e, (unsigned long long)__ldg(T.rfx + rate_index(fj)));
                } else if (cont && (fj & (LOCF | F_BIGHH)) == F_BIGHH) {
                    const int64_t cid = (int64_t)(info & 0xFFFFFFFFFFFFull) + lane * PER_LANE + J;
                    const uint32_t qv = __ldg(T.qfx + ((T.hw_rows > 1 ? L0 * 256u : 0u) + rate_index(fj)));
                    atomicAdd(((s.k & 1) ? Wk.hbig2[1] : Wk.hbig2[0]) + big_slot(T, cid), (unsigned long long)qv);
                }
            })
        }   // full recount

        // ============================================================ write back what changed
        // (aux changes only when a traveller returned; deferred agents keep the words they were read with)
        if (dirty) {
            __stcs(reinterpret_cast<uint4*>(G.flags + slot0), make_uint4(f0, f1, f2, f3));
            if (mode & (M_GO_OUT | M_GO_HOME))
                if (adirty) __stcs(reinterpret_cast<uint4*>(G.aux + slot0), make_uint4(a0, a1, a2, a3));
        }
    }

    // ================================================================ hand the deferred agents to the event kernel
    ABM_PUSH_EVENTS(defer, newly, leaver, XO)
    if (TRACE) trace_stamp(Wk, s.k, 0, t_start);

    // ================================================================ flush of the warp's home-pool accumulators
    if ((mode & M_DENSE_A) && !(mode & M_INCR)) {
        __syncwarp();
        if (lane < 2 * ABM_MAX_STATUS) {
            // lanes 0..11: pressure sums, lanes 12..23: head counts of the warp's home district
            const uint32_t b = lane < ABM_MAX_STATUS ? lane : lane - ABM_MAX_STATUS;
            const unsigned int v = acc_w[lane];
            if (v && b < A) {
                unsigned long long* const acc_k = (s.k & 1) ? Wk.acc2[1] : Wk.acc2[0];
                atomicAdd(acc_k + (size_t)(sub & Wk.rep_mask) * (2u * T.P * A) + (lane < ABM_MAX_STATUS ? 0 : (size_t)T.P * A) + (size_t)L0 * A + b,
                          (unsigned long long)v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ event kernel
// One thread per deferred agent.  What happens to it, in the reference's order for one agent: infection event of the
// previous sub-step's draw (Country.set_epidemic_status, base_model.py:838-1055), due transitions
// (update_agent_states + contagious promotion, :226-272, :54-58), go_home / go_out incl. the return from and the
// departure on a trip (:274-441), and its contribution to the district x status tables.  The hot words come with the
// queue entry (coalesced); the only scattered accesses are the agent's 32-byte cold sector and the final stores.
__global__ void __launch_bounds__(EVENT_THREADS, ABM_EVENT_MIN_BLOCKS)
abm_event_kernel(const DevTables T, const DevAgents G, const DevWork Wk, const DevClock* clk, const uint32_t mode_rt) {
    __shared__ int s_tally[N_TALLY];
    const unsigned long long t_start = Wk.trace ? global_timer_ns() : 0ull;
    if (threadIdx.x < N_TALLY) s_tally[threadIdx.x] = 0;
    grid_dependency_wait();       // the sweep has completed: queue and agent store are visible
    // first entry of this thread, requested together with the count and the clock (one round trip instead of three);
    // beyond the count it is stale data inside the queue's allocation and is not used
    uint4 e_first = __ldcs(Wk.q_entries + (size_t)blockIdx.x * Wk.q_cap + threadIdx.x);
    // the table kernel may become resident (it waits for this grid before it reads what it writes); a flush is
    // followed by a sweep, which reads the agent store in its prologue and must not start early
    if (!(mode_rt & M_NO_EARLY_DEPENDENTS)) launch_dependents();
    __syncthreads();
    const StepArgs s = step_args(T, clk, mode_rt & 0xFFFFu);
    const uint32_t mode = s.mode;
    const uint32_t A = (uint32_t)T.A;

    for (uint32_t q = blockIdx.x; q <= Wk.q_mask; q += gridDim.x) {
        const uint32_t n_gen = ld_ca(Wk.q_count + 2u * q), n = n_gen + ld_ca(Wk.q_count + 2u * q + 1u);
        const uint4* entries = Wk.q_entries + (size_t)q * Wk.q_cap;
        // generic events first, then the decided leavers: warps run one kind of body (but for the one at the boundary)
        for (uint32_t i = threadIdx.x; i < n; i += EVENT_THREADS) {
            const uint4 e = (q == blockIdx.x && i == threadIdx.x && i < n_gen) ? e_first
                                                                               : __ldcs(entries + (i < n_gen ? i : Wk.q_cap - 1u - (i - n_gen)));
            const uint32_t slot = e.x & 0x7FFFFFFFu;
            uint32_t f = e.y, a = e.z;
            const unsigned long long info = __ldg(T.sub_info + (slot >> 7));
            const int L0 = (int)(info >> 48);
            const int64_t cid = (int64_t)(info & 0xFFFFFFFFFFFFull) + (slot & (SUB - 1));
            const int J = (int)(slot & 3u);                     // word of the group Philox blocks that is this agent's
            if (e.w & EV_LEAVER) {
                // The sweep already decided: an active mover at home whose OD draw leaves the district (:366-441), no
                // infection, no due event.  Destination, length of stay; the cold sector is only written (one word).
                const uint32_t* od_row = T.od_thr + ((size_t)s.weekday * T.D + L0) * T.D;
                const uint32_t dest = od_destination(od_row, T.D, e.w & 0x7FFFFFFFu);
                if (dest != (uint32_t)L0) {                                             // :410-439
                    const uint32_t xs = word_of(group_draws(T, cid, s.k, SITE_STAY), J);
                    const int32_t t_ret = stay_return_tick(T, xs, s.weekday, (uint32_t)L0, dest, s.now);
                    if ((f & F_EPI_MASK) == EPI_S) {
                        // never infected: the rest of the cold record is known (dates = never, inf_at = 0), so the whole
                        // 32-byte sector is overwritten -- a scattered 4-byte store into a sector that is not in L2 costs a
                        // DRAM read-modify-write, and a million of those at 08:00 were half of this kernel's time
                        int4* p = reinterpret_cast<int4*>(G.cold) + (size_t)slot * 2;
                        p[0] = make_int4(T_NEVER, T_NEVER, T_NEVER, T_NEVER);
                        p[1] = make_int4(t_ret, 0, 0, 0);
                    } else {
                        G.cold[(size_t)slot * ABM_COLD_WORDS + ABM_COLD_T_RETURN] = t_ret;
                    }
                    f = set_loc(f, LOC_DEST) | F_AWAY;
                    a = (a & 0xFFFF0000u) | dest;
                } else {
                    f = set_loc(f, econ_loc(f));
                }
            } else {
                Cold c = load_cold(G, slot);
                bool cold_dirty = false;
                if (e.x >> 31) {                                     // I(k-1): the draw infected this agent
                    f = infect_agent(T, f, a, c, cid, s.k_inf, s.now_inf, (uint32_t)s.k_inf + 1u, s_tally);
                    atomicAdd(s_tally + CUM_INFECTED, 1);
                    cold_dirty = true;
                }
                if (mode & M_STEP) {
                    if ((a >> 16) <= (uint32_t)s.k) {                // T(k)
                        if (mode & M_INCR) {
                            DeltaSink sink;
                            sink.acc_r = ((s.k & 1) ? Wk.acc2[1] : Wk.acc2[0]) + (size_t)(blockIdx.x & Wk.rep_mask) * (2u * T.P * A);
                            sink.acc_n = sink.acc_r + (size_t)T.P * A;
                            sink.hbig = (s.k & 1) ? Wk.hbig2[1] : Wk.hbig2[0];
                            sink.L0 = L0; sink.cid = cid;
                            f = transition_agent(T, G, f, a, slot, c, s.now, s_tally, &sink);
                        } else {
                            f = transition_agent(T, G, f, a, slot, c, s.now, s_tally);
                        }
                    }
                    if ((mode & M_GO_HOME) && is_active(f)) {                               // base_model.py:274-302
                        if (f & F_AWAY) {
                            if (c.u.x < s.now) {                                            // :285-293
                                a = (a & 0xFFFF0000u) | (uint32_t)L0;
                                f &= ~(F_AWAY | LOCF);
                            }
                        } else {
                            f &= ~LOCF;                                                     // LOC_HOME == 0
                        }
                    }
                    if ((mode & M_GO_OUT) && is_active(f)) {                                // :304-340
                        const bool mild = T.mild_active && (f & F_ONSET) && (f & F_EPI_MASK) < EPI_R;     // :212-219, :316
                        const uint32_t cls = G.move_class[slot];
                        const uint32_t thr = __ldg(T.move_thr + (((mode & M_WEEKDAY) ? 0u : 2u) + 4u * cls + (mild ? 1u : 0u)));
                        const uint32_t xm = word_of(group_draws(T, cid, s.k, SITE_MOVER), J);
                        if ((xm >> 1) < thr) {                                              // :317-318 a mover
                            const uint32_t* od_row = T.od_thr + ((size_t)s.weekday * T.D + L0) * T.D;
                            uint32_t u = 0;
                            bool leaves = false;
                            if (f & F_DMOVER) {
                                u = word_of(group_draws(T, cid, s.k, SITE_OD), J) >> 1;
                                const uint32_t od_lo = L0 > 0 ? __ldg(od_row + L0 - 1) : 0u, od_hi = __ldg(od_row + L0);
                                leaves = !(u >= od_lo && u < od_hi);
                            }
                            bool go = false;
                            if (f & F_AWAY) {
                                // travellers: back in the home district when the stay is over (:342-360), then like everybody else
                                if (c.u.x < s.now) {
                                    a = (a & 0xFFFF0000u) | (uint32_t)L0;
                                    f &= ~F_AWAY;
                                    go = leaves;
                                }
                            } else {
                                go = leaves;
                            }
                            bool econ = !go;
                            if (go) {                                                       // :366-441
                                const uint32_t dest = od_destination(od_row, T.D, u);
                                if (dest != (uint32_t)L0) {                                 // :410-439
                                    const uint32_t xs = word_of(group_draws(T, cid, s.k, SITE_STAY), J);
                                    c.u.x = stay_return_tick(T, xs, s.weekday, (uint32_t)L0, dest, s.now);
                                    cold_dirty = true;
                                    f = set_loc(f, LOC_DEST) | F_AWAY;
                                    a = (a & 0xFFFF0000u) | dest;
                                } else {
                                    econ = true;
                                }
                            }
                            if (econ) f = set_loc(f, econ_loc(f));                          // :331-339, :346-352
                        }
                    }
                }
                if (cold_dirty) {
                    int4* p = reinterpret_cast<int4*>(G.cold) + (size_t)slot * 2;
                    p[0] = c.t;
                    p[1] = c.u;
                }
            }
            if (mode & M_STEP) {
                if (!(mode & M_INCR)) {                                                 // A(k) of this one agent
                    const Contribution cb = contribution_of(G, f, a, slot, L0);
                    if (cb.pool >= 0) {
                        unsigned long long* acc_r = ((s.k & 1) ? Wk.acc2[1] : Wk.acc2[0]) + (size_t)(blockIdx.x & Wk.rep_mask) * (2u * T.P * A);
                        const uint32_t en = (uint32_t)cb.pool * A + status_of(f);
                        atomicAdd(acc_r + (size_t)T.P * A + en, 1ull);
                        if (cb.cont) atomicAdd(acc_r + en, (unsigned long long)__ldg(T.rfx + rate_index(f)));
                    } else if (cb.big_home) {
                        const uint32_t qv = __ldg(T.qfx + ((T.hw_rows > 1 ? (uint32_t)L0 * 256u : 0u) + rate_index(f)));
                        atomicAdd(((s.k & 1) ? Wk.hbig2[1] : Wk.hbig2[0]) + big_slot(T, cid), (unsigned long long)qv);
                    }
                }
            }
            G.flags[slot] = f;
            G.aux[slot] = a;
        }
        __syncthreads();                  // every thread has read the count
        if (threadIdx.x < 2) Wk.q_count[2u * q + threadIdx.x] = 0u;
    }
    __syncthreads();
    if (threadIdx.x < N_TALLY) {
        const int v = s_tally[threadIdx.x];
        if (v) atomicAdd(Wk.tally + (size_t)(blockIdx.x & Wk.rep_mask) * N_TALLY + threadIdx.x, (unsigned long long)(long long)v);
    }
    trace_stamp(Wk, s.k, 1, t_start);
}

// ------------------------------------------------------------------------------------------------ table kernel
// District x status tables -> infection thresholds:
//   lambda[L][b] = hw[L] * (sum_a W[a][b] * R[L][a]) / Ncnt[L][b],  thr = threshold(1 - exp(-lambda))
// One block per outside pool L:
//   reduce     sum the replicas of row L of the tables of the sub-step just executed into acc_sum and
//              clear the row in every replica of the OTHER buffer (used by the next sub-step)
//   threshold  thresholds of row L from acc_sum (after the all-reduce over ranks when there is one)
// With several ranks on one node the sum over ranks is part of this kernel: block L publishes its row in the rank's
// exchange buffer (release, system scope), waits for the same row of every peer (acquire) and adds them up through
// peer loads over NVLink -- 2 A words per peer and row, no separate collective, no extra launch.  (Across nodes, or
// when peer mapping is unavailable, the rows go through ncclAllReduce between a reduce and a threshold launch.)
// Block P (the last one) folds the per-launch tallies into the running counters and, at log sub-steps,
// writes the day-log row: cumulative date counters include this sub-step's transitions, the "current"
// counters are taken before them (base_model.py:36-38: log_model_output runs before update_agent_states).
struct TableArgs {
    int32_t do_reduce, do_threshold;
    int32_t replicas;
    int32_t write_log;
    int32_t incremental;      // the sub-step just executed emitted deltas (M_INCR): add them to the kept sums
};

constexpr int TABLE_THREADS = 256;

__global__ void __launch_bounds__(TABLE_THREADS)
abm_hazard_table_kernel(const DevTables T, const DevWork Wk, DevClock* __restrict__ clk, const TableArgs ta) {
    __shared__ unsigned long long s_part[TABLE_THREADS / 32][32];
    __shared__ unsigned long long s_row[2 * ABM_MAX_STATUS];
    __shared__ int s_timeout;
    if (threadIdx.x == 0) s_timeout = 0;
    const int A = T.A, n = T.P * A, t = threadIdx.x;
    const int L = blockIdx.x;
    const unsigned long long t_start = Wk.trace ? global_timer_ns() : 0ull;
    // ---- before the dependency wait: everything that only this kernel's predecessor (the previous table kernel)
    // or the host wrote -- the clock, the kept sums, constants.  These loads overlap the tail of the sub-step kernel.
    const int k = clk->k;                               // the sub-step that is being executed by the primary grid
    const unsigned int gen = clk->xchg_gen + 1u;        // block P advances both only after every row block is done
    const int e = t & 31, g = t >> 5;                   // entry e < A: pressure sum of status e; A <= e < 2A: head count
    const size_t off = e < A ? (size_t)L * A + e : (size_t)n + (size_t)L * A + (e - A);
    unsigned long long kept = 0ull;
    double w_col[ABM_MAX_STATUS];
    double hw_L = 0.0;
    if (L < T.P) {
        if (ta.do_reduce && ta.incremental && t < 2 * A) kept = Wk.acc_local[off];
        if (ta.do_threshold && t < A) {
#pragma unroll
            for (int a = 0; a < ABM_MAX_STATUS; ++a) w_col[a] = a < A ? __ldg(T.W + a * A + t) : 0.0;
            hw_L = __ldg(T.pool_hw + L);
        }
    }
    grid_dependency_wait();       // the sub-step kernel has completed: its atomics are visible
    launch_dependents();          // the next sub-step kernel may start its table-independent prologue
    if (L < T.P) {
        const unsigned long long* __restrict__ acc = (k & 1) ? Wk.acc2[1] : Wk.acc2[0];
        unsigned long long* __restrict__ acc_next = (k & 1) ? Wk.acc2[0] : Wk.acc2[1];
        if (ta.do_reduce) {
            unsigned long long v = 0ull;
            if (e < 2 * A) {
#pragma unroll 8
                for (int r = g; r < ta.replicas; r += TABLE_THREADS / 32) {     // warp g sums every 8th replica
                    v += acc[(size_t)r * 2 * n + off];
                    acc_next[(size_t)r * 2 * n + off] = 0ull;
                }
            }
            s_part[g][e] = v;
            __syncthreads();
            if (t < 2 * A) {
                unsigned long long sum = kept;
#pragma unroll
                for (int w = 0; w < TABLE_THREADS / 32; ++w) sum += s_part[w][t];
                Wk.acc_local[off] = sum;
                Wk.acc_sum[off] = sum;
                s_row[t] = sum;
            }
        } else if (t < 2 * A) {
            s_row[t] = Wk.acc_sum[off];
        }
        if (Wk.n_ranks > 1 && ta.do_reduce) {
            // ---- sum over ranks through peer memory, PUSH + flag-in-data ("LL") protocol: this rank stores its row
            // straight into every peer's buffer; each 64-bit sum travels as (low word, generation, high word,
            // generation) -- two self-validating 8-byte halves -- so there is no fence and no separate flag, the
            // transfer is one one-way NVLink trip, and every rank polls its OWN memory for the peers' rows.
            // Buffer of rank p: [2 parities][n_ranks sources][2][P][A] elements of 16 bytes.  A slot written for
            // generation g was last read for g - 2, which its reader finished before it published g - 1 (which this
            // rank needed to get here), so the two parities never collide.
            const size_t slot = ((size_t)(gen & 1u) * Wk.n_ranks) * 2 * n + off;
            if (t < 2 * A) {
                const unsigned long long v = s_row[t];
                const uint4 pk = make_uint4((uint32_t)v, gen, (uint32_t)(v >> 32), gen);
                for (int r = 0; r < Wk.n_ranks; ++r) {
                    if (r == Wk.rank) continue;
                    st_volatile_v4(reinterpret_cast<uint4*>(Wk.xchg[r]) + slot + (size_t)Wk.rank * 2 * n, pk);
                }
            }
            trace_point(Wk, k, 7, true);          // rows pushed (latest row CTA)
            // warp g takes the peers g, g + 8, ...: lane e polls element e of that peer's row in the local buffer
            unsigned long long from_peers = 0ull;
            for (int r = g; r < Wk.n_ranks; r += TABLE_THREADS / 32) {
                if (r == Wk.rank) continue;
                if (e < 2 * A) {
                    const uint4* src = reinterpret_cast<const uint4*>(Wk.xchg[Wk.rank]) + slot + (size_t)r * 2 * n;
                    uint4 pk = ld_volatile_v4(src);
                    int spins = 0;
                    while (pk.y != gen || pk.w != gen) {
                        // a peer is gone: do not hang the GPU, but never use a stale row either -- raise the (sticky)
                        // flag the host checks at every synchronising call; this row's thresholds become zero below
                        if (++spins > (1 << 22)) { atomicExch(&clk->xchg_error, 1); s_timeout = 1; pk = make_uint4(0u, 0u, 0u, 0u); break; }
                        __nanosleep(32);
                        pk = ld_volatile_v4(src);
                    }
                    from_peers += (unsigned long long)pk.x | ((unsigned long long)pk.z << 32);
                }
            }
            s_part[g][e] = from_peers;
            __syncthreads();
            trace_point(Wk, k, 9, true);          // every peer's row arrived (latest row CTA)
            if (t < 2 * A) {
                unsigned long long total = s_row[t];
#pragma unroll
                for (int w = 0; w < TABLE_THREADS / 32; ++w) total += s_part[w][t];
                Wk.acc_sum[off] = total;
                s_row[t] = total;
            }
        }
        __syncthreads();
        if (ta.do_threshold && t < A) {
            const unsigned long long cnt = s_row[A + t];
            uint32_t thr = 0;
            if (cnt) {
                double sum = 0.0;
#pragma unroll
                for (int a = 0; a < ABM_MAX_STATUS; ++a)
                    if (a < A) sum = __dadd_rn(sum, __dmul_rn(w_col[a], (double)s_row[a]));
                const double lam = __dmul_rn(hw_L, __ddiv_rn(sum, (double)cnt));
                const unsigned long long h = lam >= (double)(ABM_OMEXP_N >> 8) ? ((unsigned long long)ABM_OMEXP_N << 24)
                                                                               : __double2ull_rz(__dmul_rn(lam, 4294967296.0));
                thr = hazard_threshold_fx(T.omexp, h);
            }
            if (s_timeout) thr = 0;              // partial sums over ranks must not infect anybody
            Wk.thr_out[L * A + t] = thr;
        }
        __syncthreads();
        trace_stamp(Wk, k, 2, t_start);
        if (t == 0 && ta.do_threshold) { __threadfence(); atomicAdd(Wk.rows_done, 1u); }
        return;
    }
    if (!ta.do_threshold) return;         // the bookkeeping below runs once per sub-step, with the thresholds
    unsigned long long* hbig_clear = (k & 1) ? Wk.hbig2[0] : Wk.hbig2[1];
    unsigned long long* hbig_cur = (k & 1) ? Wk.hbig2[1] : Wk.hbig2[0];
    for (int i = t; i < T.n_big; i += blockDim.x) {
        if (ta.incremental) hbig_cur[i] += hbig_clear[i];       // deltas of this sub-step + the sums of the previous one
        hbig_clear[i] = 0ull;
    }
    {                                     // fold the replicas of the tallies: 16 threads per replica group
        static_assert(N_TALLY == 16 && TABLE_THREADS == 256, "tally fold layout");
        const int e = t & 15, g = t >> 4;
        unsigned long long v = 0ull;
        for (int r = g; r < ta.replicas; r += 16) v += Wk.tally[(size_t)r * N_TALLY + e];
        for (int r = g; r < ta.replicas; r += 16) Wk.tally[(size_t)r * N_TALLY + e] = 0ull;
        s_part[g >> 1][(g & 1) * 16 + e] = v;
    }
    __syncthreads();
    if (t < N_TALLY) {
        unsigned long long v = 0ull;
#pragma unroll
        for (int g = 0; g < 16; ++g) v += s_part[g >> 1][(g & 1) * 16 + t];
        s_row[t] = v;
    }
    __syncthreads();
    const int n_rows = clk->n_log_rows;
    long long* log_row = Wk.daylog + (size_t)(n_rows < Wk.max_log_rows ? n_rows : Wk.max_log_rows) * ABM_LOG_STRIDE;
    const long long seeds_before = clk->seeds_before;
    const int now = clk->now;
    if (t < 12) {
        // counters[0..7] cumulative, counters[8..11] current exposed / contagious / hospitalised / critical
        long long v = Wk.counters[t];
        if (t < 8) {
            v += (long long)s_row[t];
            Wk.counters[t] = v;
            if (ta.write_log) {
                const int col = t < 6 ? t : (t == 6 ? ABM_C_ASYMPTOMATIC_COUNT : ABM_C_SYMPTOMATIC_COUNT);
                log_row[col] = v + (t == CUM_INFECTED ? seeds_before : 0);
            }
        } else {
            v += (long long)s_row[PRE_EXPOSED + (t - 8)];
            if (ta.write_log) log_row[ABM_C_CURRENT_EXPOSED_CASES + (t - 8)] = v;
            v += (long long)s_row[POST_EXPOSED + (t - 8)];
            Wk.counters[t] = v;
        }
        if (t == 0 && ta.write_log) log_row[ABM_C_TICK] = now;
    }
    __syncthreads();
    if (t == 0) {                         // advance the clock (base_model.py:177-179)
        // ... once every row block has read it (they are dispatched before this block and never wait for it)
        while (atomicAdd(Wk.rows_done, 0u) < (unsigned int)T.P) __nanosleep(32);
        *Wk.rows_done = 0u;
        clk->xchg_gen = gen;
        const int minute_abs = clk->minute_abs + T.step_ticks;
        clk->k = k + 1;
        clk->now = now + T.step_ticks;
        clk->minute_abs = minute_abs;
        clk->weekday = (clk->weekday0 + minute_abs / DAY) % 7;
        if (ta.write_log && n_rows < Wk.max_log_rows) clk->n_log_rows = n_rows + 1;
        clk->seeds_before = seeds_before + clk->seeds_at_now;
        clk->seeds_at_now = 0;
    }
    trace_stamp(Wk, k, 2, t_start);
}

// ------------------------------------------------------------------------------------------------ seeding
__global__ void abm_seed_kernel(const DevTables T, const DevAgents G, unsigned long long* tally, DevClock* clk,
                                const int64_t* slots, const int64_t* ids, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t k = clk->k, now = clk->now;
    if (i == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&clk->seeds_at_now), (unsigned long long)n);
    if (i < n) {
        const int64_t slot = slots[i];
        uint32_t a = G.aux[slot];
        Cold c = load_cold(G, (uint32_t)slot);
        const uint32_t f = infect_agent(T, G.flags[slot], a, c, ids[i], k, now, (uint32_t)k, tally);
        G.flags[slot] = f;
        G.aux[slot] = a;
        int4* p = reinterpret_cast<int4*>(G.cold) + (size_t)slot * 2;
        p[0] = c.t;
        p[1] = c.u;
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
    const uint32_t d = (uint32_t)(__ldg(T.sub_info + slot / SUB) >> 48);
    atomicAdd(pop + d, 1ull);
    const uint32_t epi = f & F_EPI_MASK;
    if ((epi == EPI_I || epi == EPI_C) && (f & F_SYMPT)) atomicAdd(sym + d, 1ull);
}

// current-state counters over all slots (abm_counters): an independent recount of what the running
// counters of the table kernel hold
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
