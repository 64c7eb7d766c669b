#!/usr/bin/env python
"""bench.py -- agent-days/s of the covid19_abm agent step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the real reference (baseline/_ref) on the
                                                                     # host cores; the numpy port only if it is absent

A "step" is one simulated day = 6 four-hour sub-steps (params.py:127) of the synthetic Zimbabwe-shaped
population: 15 M agents per GPU, 60 districts, unmitigated, R0 = 1.9 (BASELINE.json configs[1]); N GPUs hold
N x 15 M agents (weak scaling), sharded on household boundaries, the district x status tables summed over
ranks inside the table kernel (peer memory) once per sub-step.  Before the warm-up the epidemic is advanced
(untimed) for --spinup days so the timed days run at high prevalence (both arms: same seeding fraction, same
spin-up).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'covid19-agent-based-model_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = 'agent_days_per_sec'
UNIT = 'agent-days/s'
ALGO_BYTES_PER_AGENT_DAY = 104.0      # 6 x 16 B hot record + 2 x 4 B movement write-back
SUBSTEPS_PER_DAY = 6
POP_SEED = 2020


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--agents-per-gpu', type=int, default=15_000_000)
    ap.add_argument('--districts', type=int, default=60)
    ap.add_argument('--spinup', type=int, default=40, help='untimed simulated days before the warm-up')
    ap.add_argument('--seed-fraction', type=float, default=0.002)
    ap.add_argument('--cpu-agents', type=int, default=1_500_000, help='agents of the bounded CPU sample (numpy port)')
    ap.add_argument('--ref-agents', type=int, default=100_000, help='agents per replica of the real reference (baseline/_ref)')
    ap.add_argument('--ref-procs', type=int, default=0, help='replicas of the reference arm (0 = one per host core, at most 32)')
    ap.add_argument('--no-ingest', action='store_true', help='skip the timed device-side ingestion leg')
    ap.add_argument('--no-parity-check', action='store_true', help='skip the sharding-invariance check before the timed region')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--shards', default='dealt', choices=['contiguous', 'dealt'],
                    help='multi-GPU sharding: one piece of every district per rank (default: the ranks are statistically alike, '
                         'no rank is the slowest of every sub-step -- 125.6 vs 131.4 us per sub-step at N = 8, '
                         'profiles/r02_timeline_8gpu_{dealt,contiguous}.txt), or contiguous ranges of the canonical order')
    ap.add_argument('--exchange', default='p2p', choices=['p2p', 'nccl'],
                    help='N > 1: sum the tables over ranks inside the table kernel through peer memory, or with ncclAllReduce')
    ap.add_argument('--full-run-days', type=int, default=365, help='also time a whole run of this many days from the seeding (0 = skip)')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread
    (the timed region of a 15 M-agent run lasts tens of milliseconds, too short for `nvidia-smi -lms`)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _run(self):
        nv, h = self._nvml, self._handle
        bits = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get_reasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nvml = nv
            self._handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._handle, nv.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def mark(self):
        """Number of samples taken so far (to delimit a window afterwards)."""
        return len(self.samples)

    def wait_first(self, timeout=1.0):
        """Block until the polling thread has delivered its first sample (an NVML query can take several milliseconds)."""
        t0 = time.perf_counter()
        while self._nvml is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.001)

    def summary(self, ranges=None):
        """Median SM clock over the sample index ranges [(start, end), ...] (default: everything); the first range is the
        timed region."""
        if ranges is None:
            ranges = [(0, len(self.samples))]
        window = [v for lo, hi in ranges for v in self.samples[lo:hi]]
        out = {'sm_mhz': float(np.median(window)) if window else None,
               'sm_max_mhz': float(self.max_mhz) if self.max_mhz is not None else None, 'reasons': sorted(self.reasons),
               'samples': len(window)}
        if len(ranges) > 1:
            out['samples_in_timed_region'] = max(0, ranges[0][1] - ranges[0][0])
            out['window'] = ('NVML polled every ~2 ms during the timed region and during the per-launch-profiled days that follow '
                             '(same kernels; the timed region alone lasts ~20 ms, a few NVML queries)')
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_replica(args):
    """One oracle replica: `days` simulated days after `warm` warm-up days; returns seconds."""
    n_agents, districts, seed_fraction, warm, days, member = args
    from abm_b200 import synth, tables
    from oracle.abm_oracle import OracleModel
    spec = synth.make_spec(n_districts=districts, R0=1.9, seed=POP_SEED)
    pop = synth.make_population(n_agents, spec, seed=POP_SEED)
    seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(seed_fraction * n_agents)), seed=POP_SEED + member)
    ora = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=1.0), tables.build_tables(spec), 1000 + member)
    ora.set_epidemic_status(seeds)
    for _ in range(warm * SUBSTEPS_PER_DAY):
        ora.step()
    t0 = time.perf_counter()
    for _ in range(days * SUBSTEPS_PER_DAY):
        ora.step()
    return time.perf_counter() - t0


def _reference_available():
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    try:
        import ref_harness
        return ref_harness.available()
    except Exception:
        return False


def _ref_replica(args):
    """One replica of the REAL reference (unmodified Country.step(), base_model.py:807-836 -> :33-179, own numpy RNG,
    explicit contact sampling): `spinup` + `warm` untimed simulated days, then `days` timed ones; returns
    (seconds, load seconds, prevalence at the end)."""
    n_agents, districts, seed_fraction, spinup, warm, days, member = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')            # scripts/run_simulation_scenarios_full_data.py:30
    import warnings
    warnings.filterwarnings('ignore')
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_harness as rh
    from datetime import datetime
    from abm_b200 import synth
    spec = synth.make_spec(n_districts=districts, R0=1.9, seed=POP_SEED, start_datetime=datetime(2020, 9, 1, 0))
    pop = synth.make_population(n_agents, spec, seed=POP_SEED)
    world = rh.RefWorld(spec, pop, seed_cases=synth.seed_case_counts(pop, spec, seed=POP_SEED))
    try:
        np.random.seed(1000 + member)
        t0 = time.perf_counter()
        model = world.make_model('UnmitigatedScenario', seed_infected=max(1, int(seed_fraction * n_agents)), log_file=None)
        t_load = time.perf_counter() - t0
        for _ in range((spinup + warm) * SUBSTEPS_PER_DAY):
            model.step()
        t0 = time.perf_counter()
        for _ in range(days * SUBSTEPS_PER_DAY):
            model.step()
        secs = time.perf_counter() - t0
        epi = np.asarray(model.epidemic_state)
        return secs, t_load, dict(contagious=int((epi == 2).sum()), ever_infected=int((epi > 0).sum()), agents=n_agents)
    finally:
        world.close()


def cpu_baseline_sample(a):
    """CPU figures beside the CUDA line (rank 0, N = 1), bounded samples on ONE host core each: the real reference when
    baseline/_ref travelled with the tree (kind "reference"), and the numpy oracle port (kind "port")."""
    port_days, port_warm = 2, 1
    secs = _cpu_replica((a.cpu_agents, a.districts, a.seed_fraction, port_warm, port_days, 0))
    port = dict(value=a.cpu_agents * port_days / secs, unit=UNIT, cores=1, kind='port',
                sample=f'{a.cpu_agents} agents x {port_days} days after {port_warm} warm-up day(s), {a.districts} districts, numpy '
                       f'oracle (oracle/abm_oracle.py), {secs:.1f} s')
    if not _reference_available():
        return port
    # a short, heavily seeded run so that the sample has real prevalence within ~20 s of CPU time
    n, frac, spin, days = a.ref_agents, 0.02, 6, 2
    secs, t_load, prev = _ref_replica((n, a.districts, frac, spin, 0, days, 0))
    return dict(value=n * days / secs, unit=UNIT, cores=1, kind='reference',
                sample=f'unmodified reference Country.step() (baseline/_ref): {n} agents x {days} days after {spin} untimed days, '
                       f'{frac:.0%} seeded, {a.districts} districts; {prev["contagious"]} contagious / {prev["ever_infected"]} ever '
                       f'infected at the end; {secs:.1f} s (+ {t_load:.1f} s load_agents)',
                port=port)


def run_reference_arm(a):
    """--impl reference: the reference's own CPU implementation on all host cores, the way the reference uses a
    multi-core box -- P independent single-threaded runs (scripts/run_simulation_scenarios_full_data.py:30,41-46).
    Same generator, districts, seeding fraction and spin-up as the CUDA arm; a bounded sample in size (agents per
    replica) and in timed days."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = a.ref_procs if a.ref_procs > 0 else max(1, min(cores, 32))
    warm = min(a.warmup, 1)
    days = max(1, min(a.steps, 3))
    real = _reference_available()
    t0 = time.perf_counter()
    with mp.get_context('spawn').Pool(procs) as pool:
        if real:
            n = a.ref_agents
            res = pool.map(_ref_replica, [(n, a.districts, a.seed_fraction, a.spinup, warm, days, m) for m in range(procs)])
            secs = [r[0] for r in res]
            prev = res[0][2]
            what = (f'unmodified reference Country.step() (baseline/_ref, own numpy RNG, explicit contact sampling), '
                    f'{prev["contagious"]} contagious / {prev["ever_infected"]} ever infected of {n} in replica 0 at the end')
            kind = 'reference'
        else:
            n = a.cpu_agents
            secs = pool.map(_cpu_replica, [(n, a.districts, a.seed_fraction, warm, days, m) for m in range(procs)])
            what = 'numpy oracle port of the reference step (baseline/_ref not present on this box)'
            kind = 'port'
    wall = max(secs)
    value = procs * n * days / wall
    sample = (f'{procs} independent single-threaded replicas x {n} agents x {days} timed day(s) after '
              f'{(a.spinup if real else 0) + warm} untimed day(s), seed fraction {a.seed_fraction}; {what}; slowest replica '
              f'{wall:.1f} s; whole arm {time.perf_counter() - t0:.1f} s')
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=days, warmup=warm,
                ms_per_step=wall / days * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64/int64 (numpy)' if real else 'int64', data='synthetic',
                config=dict(workload=f'synthetic Zimbabwe-shaped population, {a.agents_per_gpu} agents per GPU x {a.gpus} GPU(s), '
                                     f'{a.districts} districts, unmitigated R0=1.9, 4 h sub-steps (6 per simulated day)',
                            sample=f'CPU arm: bounded sample of that workload -- {procs} independent replicas of {n} agents '
                                   f'each (same generator, same districts, same seeding fraction and spin-up as the CUDA arm), '
                                   f'{days} day(s) timed',
                            agents_per_replica=n, replicas=procs, spinup_days=a.spinup if real else 0,
                            seed_fraction=a.seed_fraction),
                cpu_baseline=dict(value=value, unit=UNIT, cores=procs, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda_arm(a):
    import torch
    from abm_b200 import distributed, sharding, synth, tables
    from abm_b200.engine import Simulation, load_library
    import __graft_entry__ as entry

    rank, world, local = distributed.init_from_env()
    if rank == 0:
        entry.build()
    if world > 1:
        torch.distributed.barrier()
    lib = load_library()
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    n_total = a.agents_per_gpu * world
    spec = synth.make_spec(n_districts=a.districts, R0=1.9, seed=POP_SEED)
    tab = tables.build_tables(spec)
    from abm_b200.engine import HostStore
    agent_ids = None
    if a.shards == 'dealt' and world > 1:
        pop, agent_ids, dstart = sharding.make_synthetic_shard_dealt(n_total, spec, POP_SEED, world, rank)
        base = int(agent_ids[0])
        local_seeds = agent_ids[synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=POP_SEED + rank)]
        store = HostStore(spec, pop, agent_ids=agent_ids)
    else:
        pop, base, dstart = sharding.make_synthetic_shard(n_total, spec, POP_SEED, world, rank)
        local_seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(a.seed_fraction * pop.n)), seed=POP_SEED + rank) + base
        store = HostStore(spec, pop, agent_id_base=base)     # ingestion (slot map, bit packing, page-locking): untimed
    max_days = max(a.spinup + a.warmup + 2 * a.steps + 16, a.full_run_days + 8)

    def make_sim(upload=True):
        uid = None
        if world > 1 and a.exchange == 'nccl':
            uid = distributed.exchange_unique_id(lib, rank, world, device)       # one id per communicator
        return Simulation(spec, pop, seed=POP_SEED, device=local, tables=tab, max_days=max_days, agent_id_base=base,
                          district_start=dstart, world_size=world, rank=rank, nccl_unique_id=uid, host_store=store,
                          upload=upload, exchange=a.exchange if world > 1 else None, agent_ids=agent_ids)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    # ---- sharding-invariance digest (every N, both layouts): a small fixed population run on the N ranks must give the
    #      single-rank oracle's day log and tables exactly; a rank summing garbage cannot print a throughput line
    parity = None
    if not a.no_parity_check:
        parity = parity_check(a, lib, rank, world, local, device)
        if parity['status'] != 'ok':
            if rank == 0:
                print('bench: sharding parity check FAILED: ' + json.dumps(parity), file=sys.stderr, flush=True)
            sys.exit(1)

    # ---- ingestion: this rank's agents as integer columns in shuffled PERSON order (what load_agents has
    #      after resolving names) -> canonical sort, sub-tile packing, bit packing on the device -> first sub-step
    ingest = None
    if agent_ids is None and not a.no_ingest:
        from abm_b200.ingest import DeviceStore
        # households in random order (members of a household keep their census order, which the stable canonical sort
        # preserves, so the device-built store can be compared word for word with the host-built one)
        perm = np.argsort(np.random.default_rng(POP_SEED + rank).permutation(pop.n_households)[pop.household], kind='stable')
        cols = {name: np.ascontiguousarray(getattr(pop, name)[perm]) for name in
                ('district', 'household', 'age', 'status', 'econ_kind', 'district_mover', 'move_prob_weekday', 'move_prob_other')}
        barrier()
        t0 = time.perf_counter()
        dstore = DeviceStore(spec, device=local, agent_id_base=base, **cols)
        t_store = time.perf_counter() - t0
        isim = Simulation(spec, dstore.pop, seed=POP_SEED, device=local, tables=tab, max_days=4, agent_id_base=base,
                          district_start=dstart, world_size=1, host_store=dstore)
        isim.step(1)
        isim.sync()
        t_first = time.perf_counter() - t0
        same = all(torch.equal(dstore.host[k].cpu(), store.host[k]) for k in store.host)
        isim.close()
        del isim, dstore, cols, perm
        torch.cuda.empty_cache()
        ingest = dict(ingest_s=t_first, store_built_s=t_store, agents=pop.n, equals_host_store=bool(same),
                      what='integer census columns with the households in random order (host numpy) -> upload, device sort by (district, '
                           'household), sub-tile packing, bit packing -> Simulation -> first sub-step executed; per rank')
        if not same:
            print('bench: device-built store differs from the host-built one', file=sys.stderr, flush=True)
            sys.exit(1)

    sim = make_sim()
    sim.seed_infections(local_seeds)
    sim.step(a.spinup * SUBSTEPS_PER_DAY)
    for _ in range(a.warmup):
        sim.step(SUBSTEPS_PER_DAY)
    barrier()
    launches0 = sim.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    clocks = ClockSampler(local).__enter__()          # left running over the profiled days below: the timed region alone is ~20 ms
    clocks.wait_first()
    barrier()
    clk_start = clocks.mark()
    ev[0].record(sim.stream)
    for d in range(a.steps):
        sim.step(SUBSTEPS_PER_DAY)
        ev[d + 1].record(sim.stream)
    barrier()
    clk_timed_end = clocks.mark()
    ms = ev[0].elapsed_time(ev[-1])
    day_ms = [ev[d].elapsed_time(ev[d + 1]) for d in range(a.steps)]
    launches = sim.launch_count - launches0
    t = torch.tensor([ms, max(day_ms)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, worst_timed_day_ms = float(t[0].item()), float(t[1].item())
    value = n_total * a.steps / (ms / 1e3)

    # ---- the state the timed days ended in
    ckpt = sim.checkpoint()
    prevalence = {name: int(ckpt['counters'][i]) for i, name in enumerate(
        ['current_contagious_count', 'infected_count', 'hospitalized_count', 'critical_count', 'died_count', 'recovered_count',
         'current_exposed_cases', 'current_contagious_cases', 'current_hospitalized_cases', 'current_critical_cases',
         'asymptomatic_count', 'symptomatic_count'])}
    del ckpt

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream (live, same state)
    clk_prof_start = clocks.mark()
    sim.set_profiling(True)
    prof_days = max(1, min(a.steps, 5))
    sim.step(prof_days * SUBSTEPS_PER_DAY)
    kms, klaunches = sim.kernel_time()
    ems, tms = sim.event_time(), sim.table_time()
    sim.set_profiling(False)
    clk_prof_end = clocks.mark()
    clocks.__exit__(None, None, None)
    clocks_summary = clocks.summary([(clk_start, clk_timed_end), (clk_prof_start, clk_prof_end)])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    per_launch_bytes = ALGO_BYTES_PER_AGENT_DAY / SUBSTEPS_PER_DAY * pop.n
    achieved = per_launch_bytes / (kms / klaunches * 1e-3) / 1e9
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, 'profiles', 'dram_traffic.json')
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get('bytes_per_launch')
            traffic_source = ('stored profile figure, not measured in this run: ' + tj.get('source', 'profiles/dram_traffic.json')
                              + ' (ncu dram__bytes_read.sum + dram__bytes_write.sum of abm_sweep_kernel, mean over the six '
                                'launches of a simulated day, 15 M agents)')
        except Exception:
            traffic = None
    roofline = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak, traffic=traffic,
                    traffic_source=traffic_source,
                    kernel='abm_sweep_kernel', avg_launch_ms=kms / klaunches, launches_timed=klaunches,
                    algorithmic_bytes_per_launch=per_launch_bytes,
                    event_kernel_avg_ms=ems / klaunches, table_kernel_avg_ms=tms / klaunches,
                    peak_source='MEASURED_PEAKS.json (measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                    own_layout_bytes_per_launch=9.0 * sim.sm.n_slots,
                    whole_day_frac=ALGO_BYTES_PER_AGENT_DAY * pop.n / (ms / a.steps * 1e-3) / 1e9 / peak,
                    worst_timed_day_frac=ALGO_BYTES_PER_AGENT_DAY * pop.n / (worst_timed_day_ms * 1e-3) / 1e9 / peak)

    scenario_days = a.full_run_days if a.full_run_days > 0 else a.steps
    scenario_seeds = synth.pick_seeds(pop, spec, infect_num=max(1, int(0.00015 * pop.n)), seed=POP_SEED + rank)
    scenario_seeds = agent_ids[scenario_seeds] if agent_ids is not None else scenario_seeds + base

    # ---- the whole scenario of BASELINE.json configs[1], device resident: 365 simulated days from the seeding, one CUDA
    #      event per day -> the cost curve of the run and its worst day
    full_run = None
    if a.full_run_days > 0:
        sim.close()
        del sim
        sim = make_sim()
        sim.seed_infections(scenario_seeds)
        sim.step(SUBSTEPS_PER_DAY)            # first day: plain launches + graph capture, untimed
        barrier()
        fe = [torch.cuda.Event(enable_timing=True) for _ in range(a.full_run_days)]
        fe[0].record(sim.stream)
        for d in range(1, a.full_run_days):
            sim.step(SUBSTEPS_PER_DAY)
            fe[d].record(sim.stream)
        barrier()
        per_day = np.array([fe[d].elapsed_time(fe[d + 1]) for d in range(a.full_run_days - 1)])
        tf = torch.tensor([fe[0].elapsed_time(fe[-1])], dtype=torch.float64, device=device)
        td = torch.tensor(per_day, dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(tf, op=torch.distributed.ReduceOp.MAX)
            torch.distributed.all_reduce(td, op=torch.distributed.ReduceOp.MAX)
        per_day = td.cpu().numpy()
        fc = sim.counters()
        worst = int(per_day.argmax())
        from datetime import timedelta
        full_run = dict(days=a.full_run_days - 1, seconds=float(tf.item()) / 1e3,
                        agent_days_per_sec=n_total * (a.full_run_days - 1) / (float(tf.item()) / 1e3),
                        seed_fraction=0.00015, infected_at_end_this_rank=fc['infected_count'], died_at_end_this_rank=fc['died_count'],
                        worst_day=dict(day=worst + 1, date=(spec.start_datetime + timedelta(days=worst + 1)).date().isoformat(),
                                       ms=float(per_day[worst]),
                                       whole_day_frac=ALGO_BYTES_PER_AGENT_DAY * pop.n / (per_day[worst] * 1e-3) / 1e9 / peak),
                        quietest_day_ms=float(per_day.min()), median_day_ms=float(np.median(per_day)),
                        day_ms_every_14=[round(float(v), 4) for v in per_day[::14]])

    # ---- end to end through the public API with HOST buffers: the WHOLE scenario.  The packed agent store sits in
    #      page-locked host memory; timed = host->device copy of every per-agent vector, the seeding call, every simulated
    #      day with a synchronising read-back of the day log (what run_scenario's log line needs), device->host copy of the
    #      final state word.  Handle / table / communicator creation is set-up (the reference's ParamsConfig + __init__).
    e2e = None
    if not a.no_e2e:
        sim.close()
        del sim
        sim = make_sim(upload=False)
        final_state = torch.empty(store.host['flags'].shape, dtype=store.host['flags'].dtype).pin_memory()
        torch.cuda.synchronize(device)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(sim.stream):
            for name, tsr in store.host.items():
                sim.buf[name].copy_(tsr, non_blocking=True)
        h2d = store.nbytes + scenario_seeds.nbytes * 2
        sim.seed_infections(scenario_seeds)
        d2h = 0
        for _ in range(scenario_days):
            sim.step(SUBSTEPS_PER_DAY)
            rows = sim.read_log()             # device -> host read of the log rows so far (synchronises)
            d2h += rows.nbytes
        with torch.cuda.stream(sim.stream):   # final packed state -> page-locked host memory
            final_state.copy_(sim.buf['flags'], non_blocking=True)
        sim.sync()
        d2h += final_state.numel() * final_state.element_size()
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(tw, op=torch.distributed.ReduceOp.MAX)
        wall = float(tw.item())
        e2e = dict(value=n_total * scenario_days / wall, unit=UNIT, h2d_bytes_per_step=h2d / scenario_days,
                   d2h_bytes_per_step=d2h / scenario_days, seconds=wall, days=scenario_days,
                   note=f'the whole {scenario_days}-day scenario through Simulation (C ABI) with the agent store in page-locked '
                        'host memory: host->device copy of every per-agent vector, the seeding call, every day followed by a '
                        'synchronising device->host read of the day log, device->host copy of the final state word; a step '
                        'is one simulated day, bytes are per day of the run')

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline_sample(a)

    if rank == 0:
        hot_mb = 9.0 * sim.sm.n_slots / 1e6
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup, ms_per_step=ms / a.steps,
            higher_is_better=True, scaling='weak', vs_baseline=None, dtype='u32/int64', data='synthetic',
            config=dict(workload=f'synthetic Zimbabwe-shaped population, {a.agents_per_gpu} agents per GPU x {world} GPU(s), '
                                 f'{a.districts} districts, unmitigated R0=1.9, 4 h sub-steps (6 per simulated day), '
                                 f'household-aligned shards' + (' (one piece of every district per rank)' if agent_ids is not None else '')
                                 + ('' if world == 1 else (
                                     ', district x status tables summed over ranks inside the table kernel through peer '
                                     'memory (NVLink)' if a.exchange == 'p2p' else ', ncclAllReduce of the district x status '
                                     'tables per sub-step')),
                        agents_total=n_total, spinup_days=a.spinup, seed_fraction=a.seed_fraction,
                        l2=f'per-GPU agent store {store.nbytes / 1e6:.0f} MB > 126 MB L2; the 9 B/agent hot record swept '
                           f'every sub-step is {hot_mb:.0f} MB (no explicit flush: re-sweeping it is the workload)',
                        state_at_timing={k: prevalence[k] for k in ('current_exposed_cases', 'current_contagious_cases',
                                                                    'infected_count', 'recovered_count')},
                        timing_regime=f'the K timed days are days {a.spinup + a.warmup}..{a.spinup + a.warmup + a.steps - 1} of an epidemic '
                                      f'seeded with {a.seed_fraction:.2%} of the agents (state_at_timing = where they end); '
                                      'worst_timed_day_ms is the most expensive of them, full_run.worst_day the most expensive day '
                                      'of the whole scenario',
                        ingest=ingest, ingest_s=None if ingest is None else ingest['ingest_s'],
                        worst_timed_day_ms=worst_timed_day_ms, timed_day_ms=[round(v, 4) for v in day_ms],
                        full_run=full_run),
            roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=launches, clocks=clocks_summary,
            parity_check=None if parity is None else parity['status'], parity=parity)
        print(json.dumps(line), flush=True)
    _teardown(sim, world)


def parity_check(a, lib, rank, world, local, device, n_agents=60_000, days=3):
    """Sharding invariance before anything is timed: `n_agents` synthetic agents in the bench's district layout run for
    `days` days on the N ranks -- contiguous shards and (N > 1) dealt shards -- and the rank-summed day log and the summed
    district x status tables must equal the single-rank numpy oracle's (rank 0 runs it; test infrastructure used as the
    checker, never as the thing measured)."""
    import torch
    from abm_b200 import distributed, sharding, synth, tables
    from abm_b200.engine import Simulation, COUNTER_NAMES
    spec = synth.make_spec(n_districts=a.districts, R0=2.6, seed=77)
    pop = synth.make_population(n_agents, spec, seed=77)
    seeds = synth.pick_seeds(pop, spec, infect_num=max(50, n_agents // 200), seed=77)
    tab = tables.build_tables(spec)
    dstart = np.searchsorted(pop.district, np.arange(spec.n_districts + 1)).astype(np.int64)
    want_rows = want_tables = None
    if rank == 0:
        from oracle.abm_oracle import OracleModel
        ora = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=spec.mild_symptom_move_prob), tab, 4242)
        ora.set_epidemic_status(seeds)
        for _ in range(SUBSTEPS_PER_DAY * days):
            ora.step()
        want_rows = np.array([[o[name] for name in COUNTER_NAMES] for o in ora.log_rows], dtype=np.int64)
        want_tables = ora.last_tables
    out = dict(status='ok', agents=n_agents, days=days, ranks=world, layouts=[])
    for layout in (['contiguous', 'dealt'] if world > 1 else ['contiguous']):
        uid = distributed.exchange_unique_id(lib, rank, world, device) if (world > 1 and a.exchange == 'nccl') else None
        if layout == 'dealt':
            idx = sharding.deal_population(pop, world)[rank]
            shard = sharding.take_population(pop, idx)
            sim = Simulation(spec, shard, seed=4242, device=local, tables=tab, agent_ids=idx, district_start=dstart,
                             world_size=world, rank=rank, nccl_unique_id=uid, max_days=days + 4,
                             exchange=a.exchange if world > 1 else None)
        else:
            lo, hi = sharding.split_population(pop, world)[rank]
            shard = sharding.slice_population(pop, lo, hi)
            idx = np.arange(lo, hi, dtype=np.int64)
            sim = Simulation(spec, shard, seed=4242, device=local, tables=tab, agent_id_base=lo, district_start=dstart,
                             world_size=world, rank=rank, nccl_unique_id=uid, max_days=days + 4,
                             exchange=a.exchange if world > 1 else None)
        sim.seed_infections(seeds[np.isin(seeds, idx)])
        sim.step(SUBSTEPS_PER_DAY * days)
        rows = sim.read_log()[:, :12].copy()
        rows = distributed.sum_over_ranks(rows, device) if world > 1 else rows
        rs, nc, th = sim.debug_tables()
        ok = True
        if rank == 0:
            ok = (rows.shape == want_rows.shape and np.array_equal(rows, want_rows) and np.array_equal(rs, want_tables['rsum'])
                  and np.array_equal(nc, want_tables['ncnt']) and np.array_equal(th, want_tables['thr_out']))
        flag = torch.tensor([0 if ok else 1], device=device)
        if world > 1:
            torch.distributed.all_reduce(flag)
        sim.close()
        out['layouts'].append(dict(layout=layout, ok=int(flag.item()) == 0,
                                   infected_count=int(rows[-1, 1]) if rows.shape[0] else None))
        if int(flag.item()) != 0:
            out['status'] = 'FAILED'
    out['what'] = (f'{n_agents} agents / {a.districts} districts / {days} days on {world} rank(s): rank-summed day log and summed '
                   f'district x status tables (pressure sums, head counts, thresholds) equal to the single-rank numpy oracle')
    return out


def _teardown(sim, world):
    """Release the handle and the process group, but never let a stuck communicator tear-down hold the job:
    the JSON line is already out, so after a grace period the process exits with status 0 regardless."""
    import torch

    def work():
        sim.close()
        if world > 1:
            torch.distributed.destroy_process_group()
    t = threading.Thread(target=work, daemon=True)
    t.start()
    t.join(timeout=30)
    sys.stdout.flush()
    sys.stderr.flush()
    if t.is_alive():
        print('bench: tear-down still running after 30 s, exiting anyway', file=sys.stderr, flush=True)
        os._exit(0)


if __name__ == '__main__':
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_cuda_arm(args)
