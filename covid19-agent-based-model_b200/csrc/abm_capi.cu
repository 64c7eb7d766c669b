// abm_capi.cu -- host side of libabm_b200.so: the C ABI declared in include/abm_b200.h.
//
// Replaces, for the hot path only, what Country / EpidemicScheduler do on the host in the reference
// (/root/reference/src/covid19_abm/base_model.py:20-441, 838-1106): owns the constant tables, the
// district x status accumulators, the event queues, the day log, the sub-step clock, and launches the kernels of
// abm_kernels.cuh (sweep, event kernel, table kernel per sub-step; whole days as CUDA graphs) on one CUDA stream.
// Multi-GPU: one handle per rank; the accumulators are summed over ranks inside the table kernel through peer memory
// (one node), or with one ncclAllReduce per sub-step (NCCL is loaded with dlopen so a single-GPU build has no
// link-time dependency on it).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "abm_kernels.cuh"

using namespace abm;

namespace {

typedef ncclResult_t (*nccl_init_rank_fn)(ncclComm_t*, int, ncclUniqueId, int);
typedef ncclResult_t (*nccl_allreduce_fn)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*nccl_destroy_fn)(ncclComm_t);
typedef const char* (*nccl_errstr_fn)(ncclResult_t);

struct SeedEvent { int32_t tick; int64_t n; };

}  // namespace

struct abm_handle {
    abm_config cfg{};
    DevTables T{};
    DevAgents G{};
    bool have_tables = false, have_agents = false;
    std::vector<void*> owned;           // device allocations to free
    unsigned long long* acc[2] = {nullptr, nullptr};   // [replicas][2][P][A] each; sub-step k accumulates into acc[k & 1]
    unsigned long long* acc_sum = nullptr;             // [2][P][A] replicas (and ranks) summed: tables of the last sub-step
    unsigned long long* acc_local = nullptr;           // [2][P][A] this rank's replicas summed, kept across incremental sub-steps
    bool tables_valid = false;                         // acc_local / hbig hold the sums of sub-step k - 1
    int n_sms = 1;
    // peer-memory exchange
    unsigned long long* xchg_self = nullptr;           // exported buffer the peers push their table rows into (16-byte elements)
    unsigned long long* xchg[ABM_MAX_RANKS] = {};      // every rank's buffer as mapped in this process
    bool peers_attached = false;
    unsigned int* rows_done = nullptr;
    uint4* q_entries = nullptr;         // event queues (see DevWork)
    uint32_t* q_count = nullptr;
    uint32_t q_mask = 0, q_cap = 0;
    unsigned long long* scratch = nullptr;   // [2 * D + 8] words for the read-back kernels (abm_counters, abm_district_symptomatic)
    unsigned long long* trace = nullptr;     // optional timeline (abm_set_trace)
    int32_t trace_rows = 0;
    bool exchange_failed = false;       // sticky: a peer did not publish its table rows in time
    int32_t replicas = 1;
    uint32_t* thr_out = nullptr;        // [P][A]
    unsigned long long* hbig[2] = {nullptr, nullptr};
    unsigned long long* tally = nullptr;   // [N_TALLY] per-launch event tallies, folded by the table kernel
    long long* counters = nullptr;      // [12] running cumulative (0..7) and current (8..11) counters
    bool daytime = false;               // between go_out and go_home: most agents are in an outside pool
    long long* daylog = nullptr;        // [max_log_rows + 1][ABM_LOG_STRIDE]; last row is scratch
    uint32_t* move_thr = nullptr;
    int32_t n_move_classes = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    DevClock* clock = nullptr;          // device-resident simulation clock (advanced by the table kernel)
    int64_t k = 0;                      // host mirror: next sub-step to execute
    bool pending_infect = false;        // I(k-1) not applied yet
    int32_t n_log_rows = 0;
    // one simulated day as a replayable CUDA graph
    cudaGraphExec_t day_graph = nullptr;
    int32_t graph_anchor_minute = -1;   // minute of day at which a replay may start
    int32_t graph_logs = 0;             // log rows one replay writes
    bool graph_failed = false;
    int64_t graph_launches = 0;
    std::vector<SeedEvent> seeds;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    double prof_ms = 0.0;
    int64_t prof_launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_table_events;   // around the table kernel(s) of a sub-step
    size_t prof_table_used = 0;
    double prof_table_ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_event_events;   // around the event kernel of a sub-step
    size_t prof_event_used = 0;
    double prof_event_ms = 0.0;
    int64_t launches = 0;
    // NCCL
    void* nccl_lib = nullptr;
    ncclComm_t comm = nullptr;
    nccl_allreduce_fn nccl_allreduce = nullptr;
    nccl_destroy_fn nccl_destroy = nullptr;
    nccl_errstr_fn nccl_errstr = nullptr;
    std::string err;
};

namespace {

int fail(abm_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CUDA_TRY(h, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail(h, ABM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

template <typename Tp>
int upload(abm_handle* h, const Tp* host, size_t count, const Tp** dev_out) {
    Tp* d = nullptr;
    if (count == 0) count = 1;
    CUDA_TRY(h, cudaMalloc(&d, count * sizeof(Tp)));
    h->owned.push_back(d);
    if (host) CUDA_TRY(h, cudaMemcpyAsync(d, host, count * sizeof(Tp), cudaMemcpyHostToDevice, h->stream));
    else CUDA_TRY(h, cudaMemsetAsync(d, 0, count * sizeof(Tp), h->stream));
    *dev_out = d;
    return ABM_OK;
}

template <typename Tp>
int alloc_zero(abm_handle* h, size_t count, Tp** dev_out) {
    const Tp* p = nullptr;
    int rc = upload<Tp>(h, nullptr, count, &p);
    *dev_out = const_cast<Tp*>(p);
    return rc;
}

int32_t hour_of(const abm_config& c, int64_t tick) { return (int32_t)(((c.start_minute_of_day + tick) / 60) % 24); }

// Launch with the programmatic-serialization attribute: the kernel may start while its predecessor in the
// stream is still draining (see launch_dependents / grid_dependency_wait in abm_kernels.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

DevWork make_work(const abm_handle* h) {
    DevWork w{};
    w.acc2[0] = h->acc[0]; w.acc2[1] = h->acc[1];
    w.thr_out = h->thr_out;
    w.hbig2[0] = h->hbig[0]; w.hbig2[1] = h->hbig[1];
    w.tally = h->tally;
    w.acc_sum = h->acc_sum;
    w.acc_local = h->acc_local;
    w.counters = h->counters;
    w.daylog = h->daylog;
    w.rep_mask = (uint32_t)h->replicas - 1u;
    w.max_log_rows = h->cfg.max_log_rows;
    w.rows_done = h->rows_done;
    w.trace = h->trace; w.trace_rows = h->trace_rows;
    w.n_sub = (uint32_t)(h->cfg.n_tiles * (ABM_TILE / ABM_SUBTILE));
    w.q_entries = h->q_entries; w.q_count = h->q_count; w.q_mask = h->q_mask; w.q_cap = h->q_cap;
    w.n_ranks = 1; w.rank = 0;
    if (h->peers_attached) {
        w.n_ranks = h->cfg.world_size; w.rank = h->cfg.rank;
        for (int r = 0; r < h->cfg.world_size; ++r) w.xchg[r] = h->xchg[r];
    }
    return w;
}

int timed_pair(abm_handle* h, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& pool, size_t& used, cudaEvent_t* e0, cudaEvent_t* e1) {
    if (used == pool.size()) {
        cudaEvent_t a = nullptr, b = nullptr;
        CUDA_TRY(h, cudaEventCreate(&a));
        CUDA_TRY(h, cudaEventCreate(&b));
        pool.emplace_back(a, b);
    }
    *e0 = pool[used].first;
    *e1 = pool[used].second;
    used++;
    return ABM_OK;
}

// The agent part of a sub-step: the sweep over every slot, then the event kernel over the agents it deferred.
int launch_main(abm_handle* h, uint32_t mode) {
    const DevWork w = make_work(h);
    cudaEvent_t e0 = nullptr, e1 = nullptr, v0 = nullptr, v1 = nullptr;
    if (h->profiling) {
        int rc = timed_pair(h, h->prof_events, h->prof_used, &e0, &e1);
        if (rc) return rc;
        if ((rc = timed_pair(h, h->prof_event_events, h->prof_event_used, &v0, &v1))) return rc;
        CUDA_TRY(h, cudaEventRecord(e0, h->stream));
    }
    const int64_t n_ctas = h->cfg.n_tiles * (ABM_TILE / ABM_SUBTILE) / WARPS;      // one warp per sub-tile
    const dim3 block(THREADS);
    uint32_t m = mode & ~(M_WEEKDAY | M_LOG);
    const bool pdl = h->cfg.use_pdl != 0;
    const dim3 grid((unsigned)n_ctas);
    // the sub-step kinds of a simulated day get their own instantiation; anything else is generic
    // (the instantiation with the timeline stamps only while a trace buffer is attached: abm_set_trace)
#define ABM_LAUNCH(MODE)                                                                                                          \
    do {                                                                                                                          \
        if (h->trace) CUDA_TRY(h, launch_pdl(abm_sweep_kernel<MODE, true>, grid, block, h->stream, pdl, h->T, h->G, w, (const DevClock*)h->clock, m));  \
        else CUDA_TRY(h, launch_pdl(abm_sweep_kernel<MODE, false>, grid, block, h->stream, pdl, h->T, h->G, w, (const DevClock*)h->clock, m)); \
    } while (0)
    if (m == (M_INFECT | M_STEP | M_INCR)) ABM_LAUNCH(M_INFECT | M_STEP | M_INCR);                       // night
    else if (m == (M_INFECT | M_STEP | M_DENSE_I | M_DENSE_A | M_INCR)) ABM_LAUNCH(M_INFECT | M_STEP | M_DENSE_I | M_DENSE_A | M_INCR);   // day
    else if (m == (M_INFECT | M_STEP | M_GO_OUT | M_DENSE_A)) ABM_LAUNCH(M_INFECT | M_STEP | M_GO_OUT | M_DENSE_A);
    else if (m == (M_INFECT | M_STEP | M_GO_HOME | M_DENSE_I)) ABM_LAUNCH(M_INFECT | M_STEP | M_GO_HOME | M_DENSE_I);
    else ABM_LAUNCH(RUNTIME_MODE);
#undef ABM_LAUNCH
    if (h->profiling) {
        CUDA_TRY(h, cudaEventRecord(e1, h->stream));
        CUDA_TRY(h, cudaEventRecord(v0, h->stream));
    }
    // a flush is followed by a sweep (which reads the agent store before its dependency wait), not by a table kernel
    const uint32_t rt_bits = (mode & M_STEP) ? 0u : M_NO_EARLY_DEPENDENTS;
    CUDA_TRY(h, launch_pdl(abm_event_kernel, dim3(h->q_mask + 1u), dim3(EVENT_THREADS), h->stream, pdl, h->T, h->G, w,
                           (const DevClock*)h->clock, m | rt_bits));
    if (h->profiling) CUDA_TRY(h, cudaEventRecord(v1, h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return ABM_OK;
}

// What the next sub-step does, from the host mirror of the clock (EpidemicScheduler.step, base_model.py:33-58).
struct SubstepPlan { uint32_t mode; bool log; };

SubstepPlan plan_substep(const abm_config& c, int64_t k, bool pending_infect, bool& daytime, bool tables_valid) {
    const int64_t tick = k * c.step_ticks;
    const int32_t hour = hour_of(c, tick);
    SubstepPlan p{};
    p.mode = M_STEP | (pending_infect ? M_INFECT : 0u);
    if (daytime && pending_infect) p.mode |= M_DENSE_I;           // the draw applied now belongs to a daytime sub-step
    if (hour == 8) { p.mode |= M_GO_OUT; daytime = true; }        // params.py:131,135
    if (hour == 16) { p.mode |= M_GO_HOME; daytime = false; }     // params.py:132,136
    if (daytime) p.mode |= M_DENSE_A;
    // nobody moves outside the two movement hours: the tables of the previous sub-step are updated by deltas
    if (tables_valid && hour != 8 && hour != 16) p.mode |= M_INCR;
    p.log = hour == 8;                                            // base_model.py:1083
    if (p.log) p.mode |= M_LOG;
    return p;
}

// Enqueue one sub-step: the agent kernel, then the table kernel(s) (with the all-reduce over ranks between
// the replica sum and the thresholds when there are several ranks).  Everything time-dependent is read
// from the device clock, so the same sequence can be captured into a graph.
int enqueue_substep(abm_handle* h, const SubstepPlan& p, int64_t* launches) {
    const abm_config& c = h->cfg;
    const int pa = c.n_pools * c.n_status;
    int rc = launch_main(h, p.mode);
    if (rc) return rc;
    const DevWork w = make_work(h);
    TableArgs ta{};
    ta.replicas = h->replicas;
    ta.write_log = p.log ? 1 : 0;
    ta.incremental = (p.mode & M_INCR) ? 1 : 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (h->profiling) {
        if ((rc = timed_pair(h, h->prof_table_events, h->prof_table_used, &t0, &t1))) return rc;
        CUDA_TRY(h, cudaEventRecord(t0, h->stream));
    }
    if (c.world_size > 1 && c.exchange == 1 && !h->peers_attached)
        return fail(h, ABM_ERR_STATE, "exchange == 1 needs abm_peer_attach before the first step");
    if (h->comm) {
        ta.do_reduce = 1; ta.do_threshold = 0;
        CUDA_TRY(h, launch_pdl(abm_hazard_table_kernel, dim3(c.n_pools), dim3(TABLE_THREADS), h->stream, c.use_pdl != 0, h->T, w, h->clock, ta));
        ncclResult_t r = h->nccl_allreduce(h->acc_sum, h->acc_sum, (size_t)2 * pa, ncclUint64, ncclSum, h->comm, h->stream);
        if (r != ncclSuccess) return fail(h, ABM_ERR_NCCL, "ncclAllReduce: %s", h->nccl_errstr ? h->nccl_errstr(r) : "?");
        ta.do_reduce = 0; ta.do_threshold = 1;
        CUDA_TRY(h, launch_pdl(abm_hazard_table_kernel, dim3(c.n_pools + 1), dim3(TABLE_THREADS), h->stream, false, h->T, w, h->clock, ta));
        *launches += 4;
    } else {
        ta.do_reduce = 1; ta.do_threshold = 1;
        CUDA_TRY(h, launch_pdl(abm_hazard_table_kernel, dim3(c.n_pools + 1), dim3(TABLE_THREADS), h->stream, c.use_pdl != 0, h->T, w, h->clock, ta));
        *launches += 3;
    }
    if (h->profiling) CUDA_TRY(h, cudaEventRecord(t1, h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return ABM_OK;
}

// A peer that did not publish its table rows in time (abm_hazard_table_kernel gives up after a bounded wait so a dead
// peer cannot hang the GPU) leaves the thresholds of that sub-step at zero and raises a device flag; the host reads it
// at every call that synchronises anyway.  Sticky: the trajectory is no longer the single-rank one.
int check_exchange(abm_handle* h) {
    if (h->exchange_failed) return fail(h, ABM_ERR_STATE, "table exchange failed earlier: a peer rank did not publish its rows in time");
    if (!h->peers_attached) return ABM_OK;
    int32_t flag = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&flag, &h->clock->xchg_error, sizeof flag, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (flag) {
        h->exchange_failed = true;
        return fail(h, ABM_ERR_STATE, "table exchange failed: a peer rank did not publish its rows in time (rank %d of %d)", h->cfg.rank, h->cfg.world_size);
    }
    return ABM_OK;
}

int check_limits(abm_handle* h, int64_t k_after, int32_t logs_after) {
    const abm_config& c = h->cfg;
    if (k_after >= 0xFFF0) return fail(h, ABM_ERR_STATE, "sub-step index exceeds the 16-bit event field");
    if ((k_after + 2) * c.step_ticks >= 0x7FFFFFFF) return fail(h, ABM_ERR_STATE, "tick overflow");
    if (logs_after > c.max_log_rows) return fail(h, ABM_ERR_STATE, "day log full (%d rows)", c.max_log_rows);
    return ABM_OK;
}

// Capture the next `steps_per_day` sub-steps into a graph.  Valid whenever the simulation is in its periodic
// regime (an infection draw pending, at least one full day executed) and starts at the same minute of day.
int build_day_graph(abm_handle* h) {
    const abm_config& c = h->cfg;
    const int spd = 1440 / c.step_ticks;
    bool daytime = h->daytime;
    cudaGraph_t graph = nullptr;
    int64_t launches = 0;
    int logs = 0;
    cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaGetLastError(); h->graph_failed = true; return ABM_OK; }
    int rc = ABM_OK;
    const bool was_profiling = h->profiling;
    h->profiling = false;
    for (int i = 0; i < spd && rc == ABM_OK; ++i) {
        const SubstepPlan p = plan_substep(c, h->k + i, true, daytime, true);
        logs += p.log ? 1 : 0;
        rc = enqueue_substep(h, p, &launches);
    }
    h->profiling = was_profiling;
    e = cudaStreamEndCapture(h->stream, &graph);
    if (rc != ABM_OK || e != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        h->graph_failed = true;          // fall back to plain launches for the rest of the run
        return ABM_OK;
    }
    e = cudaGraphInstantiate(&h->day_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaGetLastError(); h->day_graph = nullptr; h->graph_failed = true; return ABM_OK; }
    h->graph_anchor_minute = (int32_t)((c.start_minute_of_day + h->k * c.step_ticks) % 1440);
    h->graph_logs = logs;
    h->graph_launches = launches;
    return ABM_OK;
}

int collect_profile(abm_handle* h) {
    if (h->prof_used == 0) return ABM_OK;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < h->prof_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second));
        h->prof_ms += ms;
    }
    h->prof_launches += (int64_t)h->prof_used;
    h->prof_used = 0;
    for (size_t i = 0; i < h->prof_table_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->prof_table_events[i].first, h->prof_table_events[i].second));
        h->prof_table_ms += ms;
    }
    h->prof_table_used = 0;
    for (size_t i = 0; i < h->prof_event_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->prof_event_events[i].first, h->prof_event_events[i].second));
        h->prof_event_ms += ms;
    }
    h->prof_event_used = 0;
    return ABM_OK;
}

}  // namespace

extern "C" {

int abm_version(void) { return 200; }

int abm_nccl_unique_id(void* out128) {
    if (!out128) return ABM_ERR_ARG;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return ABM_ERR_NCCL;
    typedef ncclResult_t (*get_id_fn)(ncclUniqueId*);
    get_id_fn get_id = (get_id_fn)dlsym(lib, "ncclGetUniqueId");
    if (!get_id) return ABM_ERR_NCCL;
    ncclUniqueId id;
    if (get_id(&id) != ncclSuccess) return ABM_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, sizeof id);
    return ABM_OK;
}

const char* abm_last_error(abm_handle* h) { return h ? h->err.c_str() : "null handle"; }

int abm_create(const abm_config* cfg, abm_handle** out) {
    if (!cfg || !out) return ABM_ERR_ARG;
    *out = nullptr;
    abm_handle* h = new abm_handle();
    *out = h;   // returned even on failure so the caller can read the message, then abm_destroy
    h->cfg = *cfg;
    if (cfg->n_slots <= 0 || cfg->n_slots != cfg->n_tiles * ABM_TILE) return fail(h, ABM_ERR_ARG, "n_slots must be n_tiles * %d", ABM_TILE);
    if (cfg->n_slots >= (1ll << 31)) return fail(h, ABM_ERR_ARG, "at most 2^31 - 1 agent slots per rank (shard the population)");
    if (cfg->n_status < 1 || cfg->n_status > ABM_MAX_STATUS) return fail(h, ABM_ERR_ARG, "n_status must be 1..%d", ABM_MAX_STATUS);
    if (cfg->n_districts < 1 || cfg->n_districts > 65535 || cfg->n_pools < cfg->n_districts) return fail(h, ABM_ERR_ARG, "bad district / pool count");
    if (cfg->hw_rows != 1 && cfg->hw_rows != cfg->n_districts) return fail(h, ABM_ERR_ARG, "hw_rows must be 1 or n_districts");
    if (cfg->step_ticks < 2 || cfg->max_log_rows < 1) return fail(h, ABM_ERR_ARG, "bad step_ticks / max_log_rows");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(h, ABM_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    CUDA_TRY(h, cudaSetDevice(cfg->device));
    CUDA_TRY(h, cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, cfg->device));
    h->stream = (cudaStream_t)cfg->stream;
    if (!h->stream) {                        // the legacy default stream cannot be captured into a graph
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    const size_t pa = (size_t)cfg->n_pools * cfg->n_status;
    int rc;
    // replicas of the accumulators: a power of two <= ABM_MAX_REPLICAS keeping each table buffer under 8 MiB
    h->replicas = ABM_MAX_REPLICAS;
    while (h->replicas > 1 && (size_t)h->replicas * 2 * pa * sizeof(unsigned long long) > (8u << 20)) h->replicas >>= 1;
    if ((rc = alloc_zero(h, (size_t)h->replicas * 2 * pa, &h->acc[0]))) return rc;
    if ((rc = alloc_zero(h, (size_t)h->replicas * 2 * pa, &h->acc[1]))) return rc;
    if ((rc = alloc_zero(h, 2 * pa, &h->acc_sum))) return rc;
    if ((rc = alloc_zero(h, 2 * pa, &h->acc_local))) return rc;
    if ((rc = alloc_zero(h, pa, &h->thr_out))) return rc;
    if ((rc = alloc_zero(h, (size_t)cfg->n_big_households + 1, &h->hbig[0]))) return rc;
    if ((rc = alloc_zero(h, (size_t)cfg->n_big_households + 1, &h->hbig[1]))) return rc;
    if ((rc = alloc_zero(h, (size_t)h->replicas * N_TALLY, &h->tally))) return rc;
    if ((rc = alloc_zero(h, (size_t)12, &h->counters))) return rc;
    if ((rc = alloc_zero(h, (size_t)1, &h->clock))) return rc;
    if ((rc = alloc_zero(h, (size_t)1, &h->rows_done))) return rc;
    if ((rc = alloc_zero(h, (size_t)2 * cfg->n_districts + 8, &h->scratch))) return rc;
    {
        // event queues: a power of two <= 1024 of them; queue q takes the events of the sweep CTAs b with
        // b & q_mask == q, so its worst case (every agent of those CTAs deferred) is ceil(n_ctas / n_queues) tiles
        const int64_t n_ctas = cfg->n_tiles * (ABM_TILE / ABM_SUBTILE) / WARPS;
        uint32_t nq = 1;
        while (nq < 1024u && (int64_t)nq * 2 <= n_ctas) nq <<= 1;
        h->q_mask = nq - 1u;
        h->q_cap = (uint32_t)(((n_ctas + nq - 1) / nq) * (WARPS * ABM_SUBTILE));
        if ((rc = alloc_zero(h, (size_t)2 * nq, &h->q_count))) return rc;
        uint4* qe = nullptr;
        CUDA_TRY(h, cudaMalloc(&qe, (size_t)nq * h->q_cap * sizeof(uint4)));
        h->owned.push_back(qe);
        h->q_entries = qe;
    }
    {
        DevClock clk{};
        clk.k = 0; clk.now = 0; clk.minute_abs = cfg->start_minute_of_day; clk.weekday0 = cfg->start_weekday;
        clk.weekday = (cfg->start_weekday + cfg->start_minute_of_day / 1440) % 7;
        CUDA_TRY(h, cudaMemcpyAsync(h->clock, &clk, sizeof clk, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    if ((rc = alloc_zero(h, (size_t)(cfg->max_log_rows + 1) * ABM_LOG_STRIDE, &h->daylog))) return rc;
    if (cfg->world_size > 1 && cfg->exchange == 1) {
        if (cfg->world_size > ABM_MAX_RANKS) return fail(h, ABM_ERR_ARG, "peer exchange supports at most %d ranks", ABM_MAX_RANKS);
        // exported to the peers with cudaIpcGetMemHandle: its own allocation (not part of `owned`, freed after detach)
        const size_t words = (size_t)2 * cfg->world_size * 2 * pa * 2;      // [2 parities][ranks][2][P][A] elements of two 64-bit words
        CUDA_TRY(h, cudaMalloc(&h->xchg_self, words * sizeof(unsigned long long)));
        CUDA_TRY(h, cudaMemsetAsync(h->xchg_self, 0, words * sizeof(unsigned long long), h->stream));
    } else if (cfg->world_size > 1) {
        if (!cfg->nccl_unique_id) return fail(h, ABM_ERR_ARG, "world_size > 1 needs an ncclUniqueId");
        h->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h->nccl_lib) return fail(h, ABM_ERR_NCCL, "dlopen libnccl.so.2: %s", dlerror());
        nccl_init_rank_fn init = (nccl_init_rank_fn)dlsym(h->nccl_lib, "ncclCommInitRank");
        h->nccl_allreduce = (nccl_allreduce_fn)dlsym(h->nccl_lib, "ncclAllReduce");
        h->nccl_destroy = (nccl_destroy_fn)dlsym(h->nccl_lib, "ncclCommDestroy");
        h->nccl_errstr = (nccl_errstr_fn)dlsym(h->nccl_lib, "ncclGetErrorString");
        if (!init || !h->nccl_allreduce || !h->nccl_destroy) return fail(h, ABM_ERR_NCCL, "NCCL symbols missing");
        ncclUniqueId id;
        memcpy(&id, cfg->nccl_unique_id, sizeof id);
        ncclResult_t r = init(&h->comm, cfg->world_size, id, cfg->rank);
        if (r != ncclSuccess) return fail(h, ABM_ERR_NCCL, "ncclCommInitRank: %s", h->nccl_errstr ? h->nccl_errstr(r) : "?");
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ABM_OK;
}

int abm_set_tables(abm_handle* h, const abm_tables* t) {
    if (!h || !t) return ABM_ERR_ARG;
    const abm_config& c = h->cfg;
    const size_t D = c.n_districts, A = c.n_status, P = c.n_pools;
    DevTables& T = h->T;
    int rc;
    if ((rc = upload(h, t->rfx, 2 * 128, &T.rfx))) return rc;
    if ((rc = upload(h, t->qfx, (size_t)c.hw_rows * 2 * 128, &T.qfx))) return rc;
    if ((rc = upload(h, t->W, A * A, &T.W))) return rc;
    if ((rc = upload(h, t->pool_hw, P, &T.pool_hw))) return rc;
    if ((rc = upload(h, t->od_thr, 7 * D * D, &T.od_thr))) return rc;
    if ((rc = upload(h, t->stay_par, 7 * D * D * 3, &T.stay_par))) return rc;
    if ((rc = upload(h, t->dwell, (size_t)4 * 2 * (ABM_QN + 1), &T.dwell))) return rc;
    if ((rc = upload(h, t->znorm, (size_t)2 * (ABM_QN + 1), &T.znorm))) return rc;
    if ((rc = upload(h, t->p_hosp, 128, &T.p_hosp))) return rc;
    if ((rc = upload(h, t->p_crit, 128, &T.p_crit))) return rc;
    if ((rc = upload(h, t->p_hfat, 128, &T.p_hfat))) return rc;
    if ((rc = upload(h, t->severe_risk, D, &T.severe_risk))) return rc;
    const uint32_t* mt = nullptr;
    if ((rc = upload(h, t->move_thr, (size_t)c.n_move_classes * 4, &mt))) return rc;
    T.move_thr = mt;
    h->n_move_classes = c.n_move_classes;
    {
        const unsigned long long* si = nullptr;
        if ((rc = upload(h, reinterpret_cast<const unsigned long long*>(t->sub_info), (size_t)c.n_tiles * (ABM_TILE / ABM_SUBTILE), &si))) return rc;
        T.sub_info = si;
    }
    if ((rc = upload(h, t->big_start, (size_t)c.n_big_households + 1, &T.big_start))) return rc;
    if ((rc = upload(h, t->omexp, (size_t)ABM_OMEXP_N, &T.omexp))) return rc;
    T.D = c.n_districts; T.A = c.n_status; T.P = c.n_pools; T.hw_rows = c.hw_rows; T.n_big = c.n_big_households;
    T.step_ticks = c.step_ticks; T.mild_active = c.mild_active;
    T.step_inv = ~0ull / (unsigned long long)c.step_ticks + 1ull;   // floor(2^64 / step) + 1 for step >= 2 not a power of two; exact use in fire_index
    for (int r = 0; r < 10; ++r) {      // Philox4x32 key schedule
        T.rk[2 * r] = (uint32_t)c.seed + (uint32_t)r * 0x9E3779B9u;
        T.rk[2 * r + 1] = (uint32_t)(c.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    T.symptomatic_rate = c.symptomatic_rate; T.critical_fatality_rate = c.critical_fatality_rate;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // host staging buffers may be released by the caller
    h->have_tables = true;
    return ABM_OK;
}

int abm_bind_agents(abm_handle* h, const abm_agent_ptrs* dev) {
    if (!h || !dev) return ABM_ERR_ARG;
    if (!dev->flags || !dev->aux || !dev->cold || !dev->move_class || !dev->hh_index)
        return fail(h, ABM_ERR_ARG, "null agent buffer");
    if ((uintptr_t)dev->cold & 31) return fail(h, ABM_ERR_ARG, "cold records must be 32-byte aligned");
    if (((uintptr_t)dev->flags | (uintptr_t)dev->aux) & 15) return fail(h, ABM_ERR_ARG, "flags/aux must be 16-byte aligned");
    if (((uintptr_t)dev->move_class & 7) || ((uintptr_t)dev->hh_index & 3)) return fail(h, ABM_ERR_ARG, "move_class / hh_index misaligned");
    DevAgents& G = h->G;
    G.flags = dev->flags; G.aux = dev->aux; G.cold = dev->cold;
    G.move_class = dev->move_class; G.econ_pool = dev->econ_pool;
    G.hh_index = reinterpret_cast<const uint32_t*>(dev->hh_index);
    h->have_agents = true;
    return ABM_OK;
}

int abm_refresh(abm_handle* h, const uint32_t* move_thr, int32_t n_move_classes) {
    if (!h) return ABM_ERR_ARG;
    h->tables_valid = false;             // the caller may have rewritten the agent store: recount at the next sub-step
    if (move_thr) {
        if (n_move_classes < 1) return fail(h, ABM_ERR_ARG, "n_move_classes < 1");
        const uint32_t* mt = nullptr;
        int rc = upload(h, move_thr, (size_t)n_move_classes * 4, &mt);
        if (rc) return rc;
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        h->T.move_thr = mt;
        h->n_move_classes = n_move_classes;
        if (h->day_graph) {              // kernel arguments are baked into the captured day: capture again
            cudaGraphExecDestroy(h->day_graph);
            h->day_graph = nullptr;
        }
    }
    return ABM_OK;
}

int abm_seed(abm_handle* h, const int64_t* slots, const int64_t* ids, int64_t n) {
    if (!h || (n > 0 && (!slots || !ids))) return ABM_ERR_ARG;
    if (!h->have_tables || !h->have_agents) return fail(h, ABM_ERR_STATE, "tables / agents not set");
    if (n <= 0) return ABM_OK;
    int rc = abm_flush(h);
    if (rc) return rc;
    int64_t *d_slots = nullptr, *d_ids = nullptr;
    CUDA_TRY(h, cudaMalloc(&d_slots, n * sizeof(int64_t)));
    CUDA_TRY(h, cudaMalloc(&d_ids, n * sizeof(int64_t)));
    CUDA_TRY(h, cudaMemcpyAsync(d_slots, slots, n * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_ids, ids, n * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    const int32_t now = (int32_t)(h->k * h->cfg.step_ticks);
    abm_seed_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(h->T, h->G, h->tally, h->clock, d_slots, d_ids, n);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_slots);
    cudaFree(d_ids);
    h->seeds.push_back(SeedEvent{now, n});
    h->launches++;
    return ABM_OK;
}

int abm_step(abm_handle* h, int32_t n_substeps) {
    if (!h || n_substeps < 0) return ABM_ERR_ARG;
    if (!h->have_tables || !h->have_agents) return fail(h, ABM_ERR_STATE, "tables / agents not set");
    if (h->exchange_failed) return fail(h, ABM_ERR_STATE, "table exchange failed earlier: a peer rank did not publish its rows in time");
    const abm_config& c = h->cfg;
    const bool periodic = 1440 % c.step_ticks == 0;
    const int spd = periodic ? 1440 / c.step_ticks : 0;
    int32_t left = n_substeps;
    while (left > 0) {
        // whole simulated days in the periodic regime replay the captured graph
        if (periodic && c.use_graph && !h->graph_failed && !h->profiling && h->pending_infect && h->tables_valid && left >= spd && h->k >= spd) {
            if (!h->day_graph) {
                int rc = build_day_graph(h);
                if (rc) return rc;
            }
            const int32_t minute = (int32_t)((c.start_minute_of_day + h->k * c.step_ticks) % 1440);
            if (h->day_graph && minute == h->graph_anchor_minute) {
                int rc = check_limits(h, h->k + spd, h->n_log_rows + h->graph_logs);
                if (rc) return rc;
                CUDA_TRY(h, cudaGraphLaunch(h->day_graph, h->stream));
                bool daytime = h->daytime;
                for (int i = 0; i < spd; ++i) plan_substep(c, h->k + i, true, daytime, true);
                h->daytime = daytime;
                h->k += spd;
                h->n_log_rows += h->graph_logs;
                h->launches += h->graph_launches;
                left -= spd;
                continue;
            }
        }
        bool daytime = h->daytime;
        const SubstepPlan p = plan_substep(c, h->k, h->pending_infect, daytime, h->tables_valid);
        int rc = check_limits(h, h->k + 1, h->n_log_rows + (p.log ? 1 : 0));
        if (rc) return rc;
        rc = enqueue_substep(h, p, &h->launches);
        if (rc) return rc;
        h->tables_valid = true;
        h->daytime = daytime;
        if (p.log) h->n_log_rows++;
        h->k++;
        h->pending_infect = true;
        left--;
    }
    return ABM_OK;
}

int abm_flush(abm_handle* h) {
    if (!h) return ABM_ERR_ARG;
    if (!h->pending_infect) return ABM_OK;
    int rc = launch_main(h, M_INFECT | (h->daytime ? M_DENSE_I : 0u));
    if (rc) return rc;
    h->launches += 2;
    h->pending_infect = false;
    return ABM_OK;
}

int abm_sync(abm_handle* h) {
    if (!h) return ABM_ERR_ARG;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return check_exchange(h);
}

int abm_read_log(abm_handle* h, int64_t* out, int32_t max_rows, int32_t* n_rows) {
    if (!h || !out || !n_rows) return ABM_ERR_ARG;
    const int32_t n = h->n_log_rows < max_rows ? h->n_log_rows : max_rows;
    CUDA_TRY(h, cudaMemcpyAsync(out, h->daylog, (size_t)n * ABM_LOG_STRIDE * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *n_rows = n;
    return check_exchange(h);
}

int abm_counters(abm_handle* h, int64_t out[ABM_N_COUNTERS]) {
    if (!h || !out) return ABM_ERR_ARG;
    int rc = abm_flush(h);
    if (rc) return rc;
    unsigned long long* d4 = h->scratch;
    CUDA_TRY(h, cudaMemsetAsync(d4, 0, 4 * sizeof(unsigned long long), h->stream));
    const int64_t n = h->cfg.n_slots;
    abm_count_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->G, n, d4);
    unsigned long long cur[4];
    std::vector<unsigned long long> pend((size_t)h->replicas * N_TALLY);
    long long run[12], cum[8];
    CUDA_TRY(h, cudaMemcpyAsync(cur, d4, sizeof cur, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(pend.data(), h->tally, pend.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(run, h->counters, sizeof run, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if ((rc = check_exchange(h))) return rc;
    for (int i = 0; i < 8; ++i) {       // tallies of a flush are folded by the next table kernel
        cum[i] = run[i];
        for (int r = 0; r < h->replicas; ++r) cum[i] += (long long)pend[(size_t)r * N_TALLY + i];
    }
    h->launches++;
    // NOTE: cumulative date counters cover events whose date precedes the LAST executed sub-step's clock;
    // the day log (abm_read_log) holds the exact log_model_output values.
    const int64_t tick = h->k * h->cfg.step_ticks;
    int64_t seeded = 0;
    for (const SeedEvent& e : h->seeds)
        if (e.tick < tick) seeded += e.n;
    out[ABM_C_CURRENT_CONTAGIOUS_COUNT] = (int64_t)cum[CUM_CONTAGIOUS];
    out[ABM_C_INFECTED_COUNT] = (int64_t)cum[CUM_INFECTED] + seeded;
    out[ABM_C_HOSPITALIZED_COUNT] = (int64_t)cum[CUM_HOSPITALIZED];
    out[ABM_C_CRITICAL_COUNT] = (int64_t)cum[CUM_CRITICAL];
    out[ABM_C_DIED_COUNT] = (int64_t)cum[CUM_DIED];
    out[ABM_C_RECOVERED_COUNT] = (int64_t)cum[CUM_RECOVERED];
    out[ABM_C_CURRENT_EXPOSED_CASES] = (int64_t)cur[0];
    out[ABM_C_CURRENT_CONTAGIOUS_CASES] = (int64_t)cur[1];
    out[ABM_C_CURRENT_HOSPITALIZED_CASES] = (int64_t)cur[2];
    out[ABM_C_CURRENT_CRITICAL_CASES] = (int64_t)cur[3];
    out[ABM_C_ASYMPTOMATIC_COUNT] = (int64_t)cum[CUM_ASYM];
    out[ABM_C_SYMPTOMATIC_COUNT] = (int64_t)cum[CUM_SYM];
    return ABM_OK;
}

int abm_active_infections(abm_handle* h, int64_t* out) {
    if (!h || !out) return ABM_ERR_ARG;
    // running counters 8, 9 = exposed, contagious after the transitions of the last executed sub-step; the infection
    // draw of that sub-step is still pending, but it cannot infect anybody when nobody is contagious, so `== 0` means
    // what the reference's test on the state vector means -- and the fused launch sequence stays intact (no flush)
    long long cur[2];
    CUDA_TRY(h, cudaMemcpyAsync(cur, h->counters + 8, sizeof cur, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *out = (int64_t)(cur[0] + cur[1]);
    return check_exchange(h);
}

int abm_district_symptomatic(abm_handle* h, int64_t* sym, int64_t* pop) {
    if (!h || !sym || !pop) return ABM_ERR_ARG;
    int rc = abm_flush(h);
    if (rc) return rc;
    const int D = h->cfg.n_districts;
    unsigned long long* d = h->scratch + 8;
    CUDA_TRY(h, cudaMemsetAsync(d, 0, 2 * D * sizeof(unsigned long long), h->stream));
    const int64_t n = h->cfg.n_slots;
    abm_district_sym_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->T, h->G, n, d, d + D);
    CUDA_TRY(h, cudaMemcpyAsync(sym, d, D * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(pop, d + D, D * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->launches++;
    return ABM_OK;
}

int abm_debug_tables(abm_handle* h, uint64_t* rsum, uint64_t* ncnt, uint32_t* thr) {
    if (!h) return ABM_ERR_ARG;
    const size_t pa = (size_t)h->cfg.n_pools * h->cfg.n_status;
    if (h->k < 1) return fail(h, ABM_ERR_STATE, "no sub-step executed yet");
    const unsigned long long* last = h->acc_sum;                // tables of the last executed sub-step
    if (rsum) CUDA_TRY(h, cudaMemcpyAsync(rsum, last, pa * 8, cudaMemcpyDeviceToHost, h->stream));
    if (ncnt) CUDA_TRY(h, cudaMemcpyAsync(ncnt, last + pa, pa * 8, cudaMemcpyDeviceToHost, h->stream));
    if (thr) CUDA_TRY(h, cudaMemcpyAsync(thr, h->thr_out, pa * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ABM_OK;
}

int abm_set_profiling(abm_handle* h, int32_t on) {
    if (!h) return ABM_ERR_ARG;
    int rc = collect_profile(h);
    if (rc) return rc;
    h->profiling = on != 0;
    h->prof_ms = 0.0;
    h->prof_launches = 0;
    h->prof_table_ms = 0.0;
    h->prof_event_ms = 0.0;
    return ABM_OK;
}

int abm_kernel_time(abm_handle* h, double* total_ms, int64_t* launches) {
    if (!h || !total_ms || !launches) return ABM_ERR_ARG;
    int rc = collect_profile(h);
    if (rc) return rc;
    *total_ms = h->prof_ms;
    *launches = h->prof_launches;
    return ABM_OK;
}

int abm_get_state(abm_handle* h, int64_t* k, int64_t counters[ABM_N_COUNTERS], int64_t* seeded) {
    if (!h || !k || !counters || !seeded) return ABM_ERR_ARG;
    int rc = abm_counters(h, counters);          // flushes the pending infection draws
    if (rc) return rc;
    *k = h->k;
    *seeded = 0;
    for (const SeedEvent& e : h->seeds) *seeded += e.n;
    return ABM_OK;
}

int abm_set_state(abm_handle* h, int64_t k, const int64_t counters[ABM_N_COUNTERS], int64_t seeded) {
    if (!h || !counters || k < 0 || k >= 0xFFF0) return ABM_ERR_ARG;
    const abm_config& c = h->cfg;
    // running counters in the library's order: 8 cumulative, then the 4 current ones
    long long run[12];
    run[CUM_CONTAGIOUS] = counters[ABM_C_CURRENT_CONTAGIOUS_COUNT];
    run[CUM_INFECTED] = counters[ABM_C_INFECTED_COUNT] - seeded;
    run[CUM_HOSPITALIZED] = counters[ABM_C_HOSPITALIZED_COUNT];
    run[CUM_CRITICAL] = counters[ABM_C_CRITICAL_COUNT];
    run[CUM_DIED] = counters[ABM_C_DIED_COUNT];
    run[CUM_RECOVERED] = counters[ABM_C_RECOVERED_COUNT];
    run[CUM_ASYM] = counters[ABM_C_ASYMPTOMATIC_COUNT];
    run[CUM_SYM] = counters[ABM_C_SYMPTOMATIC_COUNT];
    for (int i = 0; i < 4; ++i) run[8 + i] = counters[ABM_C_CURRENT_EXPOSED_CASES + i];
    // Valid on a handle that has already run: the generation of the peer exchange is kept (peers compare their flags
    // against it), every accumulator that carries state from one sub-step to the next is cleared, and the day graph is
    // dropped (it was captured for another phase of the day).
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    DevClock clk{};
    CUDA_TRY(h, cudaMemcpy(&clk, h->clock, sizeof clk, cudaMemcpyDeviceToHost));
    const uint32_t keep_gen = clk.xchg_gen;
    clk = DevClock{};
    clk.xchg_gen = keep_gen;
    clk.k = (int32_t)k; clk.now = (int32_t)(k * c.step_ticks);
    clk.minute_abs = c.start_minute_of_day + clk.now; clk.weekday0 = c.start_weekday;
    clk.weekday = (c.start_weekday + clk.minute_abs / 1440) % 7;
    clk.n_log_rows = 0; clk.seeds_before = seeded; clk.seeds_at_now = 0;
    const size_t pa = (size_t)c.n_pools * c.n_status;
    CUDA_TRY(h, cudaMemcpyAsync(h->clock, &clk, sizeof clk, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->counters, run, sizeof run, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->tally, 0, (size_t)h->replicas * N_TALLY * sizeof(unsigned long long), h->stream));
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(h, cudaMemsetAsync(h->acc[b], 0, (size_t)h->replicas * 2 * pa * sizeof(unsigned long long), h->stream));
        CUDA_TRY(h, cudaMemsetAsync(h->hbig[b], 0, ((size_t)c.n_big_households + 1) * sizeof(unsigned long long), h->stream));
    }
    CUDA_TRY(h, cudaMemsetAsync(h->acc_local, 0, 2 * pa * sizeof(unsigned long long), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->acc_sum, 0, 2 * pa * sizeof(unsigned long long), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->rows_done, 0, sizeof(unsigned int), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->q_count, 0, (size_t)2 * (h->q_mask + 1u) * sizeof(uint32_t), h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->day_graph) { cudaGraphExecDestroy(h->day_graph); h->day_graph = nullptr; }
    h->k = k;
    h->n_log_rows = 0;
    h->pending_infect = false;
    h->tables_valid = false;
    h->seeds.clear();
    if (seeded) h->seeds.push_back(SeedEvent{-1, seeded});
    // who is out and about follows from the hour: go_out at 8, go_home at 16 (params.py:131-136)
    h->daytime = false;
    for (int64_t j = k - 1; j >= 0 && j > k - 1 - 1440 / c.step_ticks - 1; --j) {
        const int32_t hour = hour_of(c, j * c.step_ticks);
        if (hour == 8) { h->daytime = true; break; }
        if (hour == 16) { h->daytime = false; break; }
    }
    return ABM_OK;
}

int abm_peer_export(abm_handle* h, void* out_handle) {
    if (!h || !out_handle) return ABM_ERR_ARG;
    if (!h->xchg_self) return fail(h, ABM_ERR_STATE, "no exchange buffer (world_size > 1 and exchange == 1 required)");
    static_assert(sizeof(cudaIpcMemHandle_t) == ABM_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t mh;
    CUDA_TRY(h, cudaIpcGetMemHandle(&mh, h->xchg_self));
    memcpy(out_handle, &mh, sizeof mh);
    return ABM_OK;
}

int abm_peer_attach(abm_handle* h, const void* handles, int32_t n_ranks) {
    if (!h || !handles) return ABM_ERR_ARG;
    if (!h->xchg_self || n_ranks != h->cfg.world_size) return fail(h, ABM_ERR_ARG, "abm_peer_attach: %d handles for world size %d", n_ranks, h->cfg.world_size);
    if (h->peers_attached) return fail(h, ABM_ERR_STATE, "peers already attached");
    const char* bytes = static_cast<const char*>(handles);
    for (int r = 0; r < n_ranks; ++r) {
        if (r == h->cfg.rank) { h->xchg[r] = h->xchg_self; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, bytes + (size_t)r * ABM_IPC_HANDLE_BYTES, sizeof mh);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q)
                if (q != h->cfg.rank && h->xchg[q]) { cudaIpcCloseMemHandle(h->xchg[q]); h->xchg[q] = nullptr; }
            return fail(h, ABM_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        h->xchg[r] = static_cast<unsigned long long*>(p);
    }
    h->peers_attached = true;
    if (h->day_graph) { cudaGraphExecDestroy(h->day_graph); h->day_graph = nullptr; }
    return ABM_OK;
}

int abm_peer_detach(abm_handle* h) {
    if (!h) return ABM_ERR_ARG;
    if (!h->peers_attached) return ABM_OK;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->day_graph) { cudaGraphExecDestroy(h->day_graph); h->day_graph = nullptr; }    // kernel arguments hold the mapped pointers
    for (int r = 0; r < h->cfg.world_size; ++r) {
        if (r != h->cfg.rank && h->xchg[r]) cudaIpcCloseMemHandle(h->xchg[r]);
        h->xchg[r] = nullptr;
    }
    h->peers_attached = false;
    return ABM_OK;
}

int abm_table_time(abm_handle* h, double* total_ms) {
    if (!h || !total_ms) return ABM_ERR_ARG;
    int rc = collect_profile(h);
    if (rc) return rc;
    *total_ms = h->prof_table_ms;
    return ABM_OK;
}

int abm_set_trace(abm_handle* h, int32_t n_substeps) {
    if (!h || n_substeps < 0) return ABM_ERR_ARG;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->day_graph) { cudaGraphExecDestroy(h->day_graph); h->day_graph = nullptr; }     // kernel arguments change
    if (h->trace) { cudaFree(h->trace); h->trace = nullptr; h->trace_rows = 0; }
    if (n_substeps == 0) return ABM_OK;
    const size_t words = (size_t)n_substeps * ABM_TRACE_WORDS;
    CUDA_TRY(h, cudaMalloc(&h->trace, words * sizeof(unsigned long long)));
    std::vector<unsigned long long> init(words);
    for (size_t i = 0; i < words; ++i) init[i] = ((i % ABM_TRACE_WORDS) & 1) ? 0ull : ~0ull;        // [min start, max end] pairs; debug words 6.. (even: min, odd: max)
    CUDA_TRY(h, cudaMemcpy(h->trace, init.data(), words * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    h->trace_rows = n_substeps;
    return ABM_OK;
}

int abm_read_trace(abm_handle* h, uint64_t* out, int32_t n_substeps) {
    if (!h || !out || n_substeps < 0 || n_substeps > h->trace_rows) return ABM_ERR_ARG;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemcpy(out, h->trace, (size_t)n_substeps * ABM_TRACE_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return ABM_OK;
}

int abm_event_time(abm_handle* h, double* total_ms) {
    if (!h || !total_ms) return ABM_ERR_ARG;
    int rc = collect_profile(h);
    if (rc) return rc;
    *total_ms = h->prof_event_ms;
    return ABM_OK;
}

int64_t abm_current_step(abm_handle* h) { return h ? h->k : -1; }
int64_t abm_launch_count(abm_handle* h) { return h ? h->launches : -1; }

void abm_destroy(abm_handle* h) {
    if (!h) return;
    if (h->stream) cudaStreamSynchronize(h->stream);
    // the captured day holds a reference on the communicator (its all-reduce nodes): NCCL's destroy waits
    // until every graph that captured it is gone, so the graph must die first
    if (h->day_graph) { cudaGraphExecDestroy(h->day_graph); h->day_graph = nullptr; }
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && h->nccl_destroy) h->nccl_destroy(h->comm);
    for (auto& p : h->prof_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto& p : h->prof_table_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto& p : h->prof_event_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    abm_peer_detach(h);
    if (h->trace) cudaFree(h->trace);
    if (h->xchg_self) cudaFree(h->xchg_self);
    for (void* p : h->owned) cudaFree(p);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

}  // extern "C"
