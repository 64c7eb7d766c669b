/*
 * abm_b200.h -- C ABI of the B200 agent-step library (libabm_b200.so).
 *
 * The reference (worldbank/covid19-agent-based-model) is pure Python and has no FFI of its own; the seam
 * this library sits behind is the object pair Country <-> Country.scheduler
 * (/root/reference/src/covid19_abm/base_model.py:518) and the calls the Python layers make through it.
 * Each entry point below names the reference interface it replaces.  All functions return 0 on
 * success or a negative abm_status; abm_last_error() gives the message.  Plain pointers and sizes
 * only: no torch types cross this boundary.
 *
 * Ownership: every per-agent buffer is allocated by the caller (torch tensors on the handle's device)
 * and must outlive the handle; the library owns its constant tables, accumulators, streams/events and
 * (multi-GPU) its NCCL communicator.  Threading: one host thread per handle; all device work is
 * enqueued on the stream given at creation.
 */
#ifndef ABM_B200_H
#define ABM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABM_TILE 1024            /* agent slots per tile (one thread block of 8 warps)            */
#define ABM_SUBTILE 128          /* agent slots per sub-tile (one warp): whole households of one   */
                                 /* home district; slot offset o holds canonical id base + o       */
#define ABM_BIG_HOUSEHOLD 32     /* households above this size use the global-atomic path         */
#define ABM_MAX_STATUS 12        /* economic statuses                                             */
#define ABM_QN 4096              /* bins of the quantile tables                                    */
#define ABM_OMEXP_N 8192         /* entries of the 1 - exp(-i/256) table (hazards up to 32)        */
#define ABM_N_COUNTERS 12        /* fields of Country.log_model_output (base_model.py:1086-1104)   */
#define ABM_LOG_STRIDE 16        /* int64 words per day-log row                                    */
#define ABM_MAX_RANKS 16         /* ranks of one node that can share tables through peer memory    */
#define ABM_IPC_HANDLE_BYTES 64  /* size of the opaque handle abm_peer_export produces             */
#define ABM_TRACE_WORDS 12       /* uint64 words per sub-step of the device-side timeline            */
#define ABM_COLD_WORDS 8         /* int32 words of an agent's cold (event-time) record              */
#define ABM_COLD_T_INFECTED 0
#define ABM_COLD_T_CONTAGIOUS 1
#define ABM_COLD_T_ONSET 2
#define ABM_COLD_T_RECOVERED 3
#define ABM_COLD_T_RETURN 4
#define ABM_COLD_INF_AT 5

typedef enum {
    ABM_OK = 0,
    ABM_ERR_ARG = -1,
    ABM_ERR_CUDA = -2,
    ABM_ERR_STATE = -3,
    ABM_ERR_NCCL = -4
} abm_status;

/* order of the 12 counters in abm_counters()/day-log rows (names as in base_model.py:1086-1104) */
enum {
    ABM_C_CURRENT_CONTAGIOUS_COUNT = 0, /* sum(date_start_contagious < now)  */
    ABM_C_INFECTED_COUNT = 1,           /* sum(date_infected < now)          */
    ABM_C_HOSPITALIZED_COUNT = 2,
    ABM_C_CRITICAL_COUNT = 3,
    ABM_C_DIED_COUNT = 4,
    ABM_C_RECOVERED_COUNT = 5,
    ABM_C_CURRENT_EXPOSED_CASES = 6,    /* sum(epidemic_state == INFECTED)   */
    ABM_C_CURRENT_CONTAGIOUS_CASES = 7,
    ABM_C_CURRENT_HOSPITALIZED_CASES = 8,
    ABM_C_CURRENT_CRITICAL_CASES = 9,
    ABM_C_ASYMPTOMATIC_COUNT = 10,
    ABM_C_SYMPTOMATIC_COUNT = 11,
    ABM_C_TICK = 12                     /* minute tick of the log row        */
};

typedef struct abm_handle abm_handle;

/* ParamsConfig scalars the step needs (params.py:127-136, 179-192, 199-215). */
typedef struct {
    int64_t n_slots;            /* padded agent slots on this rank, multiple of ABM_TILE            */
    int64_t n_tiles;
    int64_t agent_id_base;      /* canonical id of this rank's first agent (Philox key)             */
    int32_t n_districts;        /* D                                                                */
    int32_t n_status;           /* A <= ABM_MAX_STATUS                                              */
    int32_t n_pools;            /* P = D + extra outside locations (schools)                        */
    int32_t hw_rows;            /* 1, or D under HANDWASHING_RISK* (base_model.py:189-191)          */
    int32_t n_move_classes;
    int32_t n_big_households;
    int32_t step_ticks;         /* minutes per sub-step (params.step_timedelta)                     */
    int32_t start_minute_of_day;/* scheduler.real_time at step 0                                    */
    int32_t start_weekday;      /* datetime.weekday() at step 0                                     */
    int32_t mild_active;        /* MILD_SYMPTOM_MOVEMENT_PROBABILITY < 1                             */
    int32_t max_log_rows;       /* capacity of the device day log                                   */
    int32_t device;             /* CUDA device ordinal                                              */
    uint64_t seed;              /* Philox key                                                       */
    double symptomatic_rate;    /* params.SYMPTOMATIC_RATE                                          */
    double critical_fatality_rate;
    void* stream;               /* cudaStream_t to enqueue on (NULL = the library creates its own)  */
    int32_t use_graph;          /* replay whole simulated days as one CUDA graph (needs a capturable stream) */
    int32_t use_pdl;            /* programmatic dependent launch: the next kernel's prologue overlaps the table kernel */
    int32_t world_size;         /* ranks sharing the district tables (1 = no collective)            */
    int32_t rank;
    const void* nccl_unique_id; /* 128-byte ncclUniqueId from rank 0 when world_size > 1 and exchange == 0 */
    int32_t exchange;           /* how the district x status tables are summed over ranks: 0 = ncclAllReduce between two */
                                /* table launches; 1 = inside the table kernel through peer memory (one node; needs      */
                                /* abm_peer_export / abm_peer_attach before the first step)                              */
    int32_t reserved;
} abm_config;

/* Host pointers to the derived constant tables (abm_b200/tables.py); copied to the device. */
typedef struct {
    const uint32_t* rfx;          /* [2][128]  infection rate * 2^24 by (symptomatic, age)   params.py:236-242 */
    const uint32_t* qfx;          /* [hw_rows][2][128]  -log(1 - 3 r hw) * 2^24              base_model.py:184-191 */
    const double*   W;            /* [A][A]  k_a * M[a][b] * 2^-24                            params.py:278-297 */
    const double*   pool_hw;      /* [P]     hand-washing multiplier of each outside pool               */
    const uint32_t* od_thr;       /* [7][D][D] ceil(cumulative OD prob * 2^31)                params.py:252-259 */
    const double*   stay_par;     /* [7][D][D][3] (mean h, a, b) of the stay transform        params.py:576-580 */
    const uint32_t* dwell;        /* [4][2][QN+1] dwell quantiles, minutes                    params.py:145-189 */
    const int32_t*  znorm;        /* [2][QN+1] half-normal quantile * 2^16                               */
    const double*   p_hosp;       /* [128] AGE_HOSPITALIZATION_PROBABILITY                    params.py:223 */
    const double*   p_crit;       /* [128] AGE_CRITICAL_CARE_PROBABILITY                      params.py:224 */
    const double*   p_hfat;       /* [128] AGE_HOSPITALIZATION_FATALITY_PROBABILITY           params.py:225 */
    const double*   severe_risk;  /* [D]   severe_disease_risk (ones under quirk Q2)          base_model.py:559 */
    const uint32_t* move_thr;     /* [n_move_classes][4] mover thresholds                     base_model.py:316-318 */
    const uint64_t* sub_info;     /* [n_tiles * 8] per sub-tile: canonical id of slot 0 (multiple of 4) */
                                  /*   in bits 0..47, home district of its agents in bits 48..63        */
    const int64_t*  big_start;    /* [n_big_households+1] canonical id of first member (+ sentinel)     */
    const uint32_t* omexp;        /* [ABM_OMEXP_N] round((1 - exp(-i/256)) * 2^32), the hazard -> probability table */
} abm_tables;

/* Device pointers of the caller-owned SoA agent store (Country's vectors, base_model.py:457-517). */
typedef struct {
    uint32_t* flags;        /* [n_slots] packed state word                       */
    uint32_t* aux;          /* [n_slots] next-event sub-step << 16 | current district */
    int32_t*  cold;         /* [n_slots][ABM_COLD_WORDS] event-time record, one 32-byte sector per agent,  */
                            /* touched only when an event fires: word ABM_COLD_T_INFECTED date_infected,   */
                            /* _T_CONTAGIOUS date_start_contagious, _T_ONSET date_start_symptomatic,       */
                            /* _T_RECOVERED date_recovered, _T_RETURN return_district_at, _INF_AT          */
                            /* infected_at_{district,location}_ids (kind << 16 | district); minute ticks,  */
                            /* INT32_MAX = datetime.max; 32-byte aligned                                    */
    const uint16_t* move_class; /* movement-probability class                    */
    const uint32_t* econ_pool;  /* outside pool of EKIND_POOL agents, may be NULL */
    const uint8_t*  hh_index;   /* [n_slots] household index within the sub-tile  */
} abm_agent_ptrs;

/* Country.__init__ + EpidemicScheduler.__init__ (base_model.py:24-31, 445-525). */
int abm_create(const abm_config* cfg, abm_handle** out);
/* ParamsConfig derived tables -> device (params.py:208-297). */
int abm_set_tables(abm_handle* h, const abm_tables* t);
/* Country.initialize_*_vectors (base_model.py:527-634): bind the caller's device buffers. */
int abm_bind_agents(abm_handle* h, const abm_agent_ptrs* dev);
/* Host hooks changed district_mover / movement probabilities / economic_activity_location_ids
 * (base_model.py:642-731): new mover-threshold table (may be NULL = unchanged). */
int abm_refresh(abm_handle* h, const uint32_t* move_thr, int32_t n_move_classes);
/* Country.set_epidemic_status(ids) at scheduler.real_time (base_model.py:796, 838-1055).
 * slots/ids are host arrays: slot index and canonical id of each agent to infect. */
int abm_seed(abm_handle* h, const int64_t* slots, const int64_t* ids, int64_t n);
/* n_substeps x EpidemicScheduler.step() (base_model.py:33-179); returns after enqueue. */
int abm_step(abm_handle* h, int32_t n_substeps);
/* Apply the infection draws of the last completed sub-step (they are otherwise fused into the next
 * sub-step's kernel) so the agent vectors read exactly as after the reference's step(). */
int abm_flush(abm_handle* h);
/* Country.log_model_output rows produced so far (base_model.py:1082-1106): copies up to max_rows rows
 * of ABM_LOG_STRIDE int64 (indices ABM_C_*) and returns the number of rows; synchronises. */
int abm_read_log(abm_handle* h, int64_t* out, int32_t max_rows, int32_t* n_rows);
/* The 12 counters at the current time, over this rank's agents (synchronises, flushes). */
int abm_counters(abm_handle* h, int64_t out[ABM_N_COUNTERS]);
/* run_scenario's stop test `epidemic_state in (INFECTED, CONTAGIOUS)).any()` (scenario_models.py:494-495): exposed +
 * contagious agents after the transitions of the last executed sub-step, from the running device counters.  Synchronises
 * but launches nothing and does NOT flush, so a driver loop that calls it after every step() keeps the fused launch
 * sequence; 0 if and only if the reference's test is false. */
int abm_active_infections(abm_handle* h, int64_t* out);
/* Country.get_safe_and_unsafe_districts inputs (base_model.py:733-758): per home district, number of
 * symptomatic infected/contagious agents and population, over this rank's agents. */
int abm_district_symptomatic(abm_handle* h, int64_t* sym, int64_t* pop);
/* Device pointers of the district x status tables of the last sub-step: R (fixed-point infectious
 * pressure, uint64 [P][A]), Ncnt (uint64 [P][A]) snapshot and thresholds (uint32 [P][A]); for tests. */
int abm_debug_tables(abm_handle* h, uint64_t* rsum, uint64_t* ncnt, uint32_t* thr);
/* Bracket every main-kernel launch with CUDA events (bench roofline); totals since last reset. */
int abm_set_profiling(abm_handle* h, int32_t on);
int abm_kernel_time(abm_handle* h, double* total_ms, int64_t* launches);
/* Same bracket around the table kernel(s) (and the all-reduce) of every sub-step. */
int abm_table_time(abm_handle* h, double* total_ms);
/* Same bracket around the event kernel (the agents the sweep deferred: infections, due transitions, departures). */
int abm_event_time(abm_handle* h, double* total_ms);
/* Device-side timeline (measurement only): after abm_set_trace(h, n) every kernel of sub-steps 0..n-1 stamps the
 * window [first CTA start, last CTA end] of its launch with the GPU's global nanosecond timer; abm_read_trace copies
 * n rows of ABM_TRACE_WORDS uint64 (sweep start, end, event kernel start, end, table kernel start, end, then
 * kernel-internal debug stamps).  n = 0 switches it off. */
int abm_set_trace(abm_handle* h, int32_t n_substeps);
int abm_read_trace(abm_handle* h, uint64_t* out, int32_t n_substeps);
/* Checkpoint / resume (the reference pickles the whole model, scenario_models.py:497-500): the agent
 * vectors are the caller's buffers; these two calls carry the rest -- the clock, the 12 running counters
 * and the number of seeded infections.  abm_get_state applies pending draws first (abm_flush). */
int abm_get_state(abm_handle* h, int64_t* k, int64_t counters[ABM_N_COUNTERS], int64_t* seeded);
int abm_set_state(abm_handle* h, int64_t k, const int64_t counters[ABM_N_COUNTERS], int64_t seeded);   /* also valid on a handle that has run */
/* scheduler.steps (base_model.py:177): sub-steps enqueued so far. */
int64_t abm_current_step(abm_handle* h);
/* Kernels launched by this handle so far (graph replays count their nodes); for bench.py's `gpu_launches`. */
int64_t abm_launch_count(abm_handle* h);
/* Wait for everything enqueued on the handle's stream (the reference's step() is synchronous; abm_step is not).
 * Like every synchronising call it returns ABM_ERR_STATE (sticky) when a peer rank failed to publish its table rows
 * in time during a multi-GPU sub-step. */
int abm_sync(abm_handle* h);
/* Message of the last failed call on this handle (every entry point returns 0 or a negative abm_status). */
const char* abm_last_error(abm_handle* h);
/* Frees what the library owns (tables, accumulators, graph, communicator); never the caller's agent buffers. */
void abm_destroy(abm_handle* h);
/* ABI version, major * 100 + minor. */
int abm_version(void);
/* ncclGetUniqueId into a 128-byte buffer (rank 0 calls it and broadcasts the bytes out of band). */
int abm_nccl_unique_id(void* out128);
/* Peer-memory exchange (abm_config.exchange == 1): every rank exports the handle of its exchange buffer
 * (ABM_IPC_HANDLE_BYTES bytes), the caller gathers the handles of all ranks in rank order (any out-of-band
 * transport) and hands them to abm_peer_attach, which maps the peers' buffers.  abm_peer_detach unmaps them; call
 * it on every rank, then synchronise the ranks, before abm_destroy frees the exported buffer. */
int abm_peer_export(abm_handle* h, void* out_handle);
int abm_peer_attach(abm_handle* h, const void* handles, int32_t n_ranks);
int abm_peer_detach(abm_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* ABM_B200_H */
