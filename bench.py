#!/usr/bin/env python
"""bench.py -- agent-days/s of the covid19_abm agent step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port on host cores)

A "step" is one simulated day = 6 four-hour sub-steps (params.py:127) of the synthetic Zimbabwe-shaped
population: 15 M agents per GPU, 60 districts, unmitigated, R0 = 1.9 (BASELINE.json configs[1]); N GPUs hold
N x 15 M agents (weak scaling), sharded on household boundaries, one NCCL all-reduce of the
district x status tables per sub-step.  Before the warm-up the epidemic is advanced (untimed) for
--spinup days so the timed days run at substantial prevalence, the expensive regime for every
implementation.  One JSON line is printed by rank 0.
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
    ap.add_argument('--cpu-agents', type=int, default=1_500_000, help='agents of the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--shards', default='contiguous', choices=['contiguous', 'dealt'],
                    help='multi-GPU sharding: contiguous ranges of the canonical order, or one piece of every district per rank')
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

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': float(self.max_mhz), 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


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


def cpu_baseline_sample(n_agents, districts, seed_fraction, days=2, warm=1):
    """Oracle port (numpy restatement of the reference step) on ONE host core, bounded sample."""
    secs = _cpu_replica((n_agents, districts, seed_fraction, warm, days, 0))
    return dict(value=n_agents * days / secs, unit=UNIT, cores=1, kind='port',
                sample=f'{n_agents} agents x {days} days after {warm} warm-up day(s), {districts} districts, numpy oracle '
                       f'(oracle/abm_oracle.py), {secs:.1f} s')


def run_reference_arm(a):
    """--impl reference: the CPU implementation on all host cores, the way the reference uses a multi-core box:
    P independent single-threaded replicas (scripts/run_simulation_scenarios_full_data.py:30,41-46)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    n = a.cpu_agents
    warm = min(a.warmup, 1)
    days = max(1, min(a.steps, 3))
    t0 = time.perf_counter()
    with mp.get_context('spawn').Pool(procs) as pool:
        secs = pool.map(_cpu_replica, [(n, a.districts, a.seed_fraction, warm, days, m) for m in range(procs)])
    wall = max(secs)
    value = procs * n * days / wall
    sample = (f'{procs} independent replicas x {n} agents x {days} days ({warm} warm-up), numpy oracle port of the '
              f'reference step; slowest replica {wall:.1f} s; whole arm {time.perf_counter() - t0:.1f} s')
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=days, warmup=warm,
                ms_per_step=wall / days * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='int64',
                data='synthetic',
                config=dict(workload=f'synthetic Zimbabwe-shaped population, {a.agents_per_gpu} agents per GPU x {a.gpus} GPU(s), '
                                     f'{a.districts} districts, unmitigated R0=1.9, 4 h sub-steps (6 per simulated day)',
                            sample=f'CPU arm: bounded sample of that workload -- {procs} independent replicas of {n} agents '
                                   f'each (same generator, same districts), {days} day(s) after {warm} warm-up day(s)',
                            agents_per_replica=n, replicas=procs),
                cpu_baseline=dict(value=value, unit=UNIT, cores=procs, kind='port', sample=sample),
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

    sim = make_sim()
    sim.seed_infections(local_seeds)
    sim.step(a.spinup * SUBSTEPS_PER_DAY)
    for _ in range(a.warmup):
        sim.step(SUBSTEPS_PER_DAY)
    barrier()
    launches0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(sim.stream)
        for _ in range(a.steps):
            sim.step(SUBSTEPS_PER_DAY)
        e1.record(sim.stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    value = n_total * a.steps / (ms / 1e3)

    # ---- checkpoint of the state reached after the timed days (same epidemic regime for the e2e leg)
    ckpt = sim.checkpoint()
    prevalence = {name: int(ckpt['counters'][i]) for i, name in enumerate(
        ['current_contagious_count', 'infected_count', 'hospitalized_count', 'critical_count', 'died_count', 'recovered_count',
         'current_exposed_cases', 'current_contagious_cases', 'current_hospitalized_cases', 'current_critical_cases',
         'asymptomatic_count', 'symptomatic_count'])}

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream (live, same state)
    sim.set_profiling(True)
    prof_days = max(1, min(a.steps, 5))
    sim.step(prof_days * SUBSTEPS_PER_DAY)
    kms, klaunches = sim.kernel_time()
    ems, tms = sim.event_time(), sim.table_time()
    sim.set_profiling(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    per_launch_bytes = ALGO_BYTES_PER_AGENT_DAY / SUBSTEPS_PER_DAY * pop.n
    achieved = per_launch_bytes / (kms / klaunches * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'dram_traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get('bytes_per_launch')
        except Exception:
            traffic = None
    roofline = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak, traffic=traffic,
                    kernel='abm_sweep_kernel', avg_launch_ms=kms / klaunches, launches_timed=klaunches,
                    algorithmic_bytes_per_launch=per_launch_bytes,
                    event_kernel_avg_ms=ems / klaunches, table_kernel_avg_ms=tms / klaunches,
                    peak_source='MEASURED_PEAKS.json (measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                    own_layout_bytes_per_launch=8.0 * sim.sm.n_slots,
                    whole_day_frac=ALGO_BYTES_PER_AGENT_DAY * pop.n / (ms / a.steps * 1e-3) / 1e9 / peak)

    # ---- end to end through the public API with HOST buffers: the checkpointed agent store sits in page-locked
    #      host memory; timed = upload of every per-agent vector + tables, K days with a log read-back per day,
    #      download of the final packed state word
    e2e = None
    if not a.no_e2e:
        sim.close()
        del sim
        pinned = {name: torch.from_numpy(arr).pin_memory() for name, arr in ckpt['arrays'].items()}
        host_ckpt = dict(arrays=pinned, k=ckpt['k'], counters=ckpt['counters'], seeded=ckpt['seeded'])
        sim = make_sim(upload=False)          # handle, empty device buffers, constant tables, communicator: set-up, untimed
        final_state = torch.empty(pinned['flags'].shape, dtype=pinned['flags'].dtype).pin_memory()   # page-locked landing buffer
        torch.cuda.synchronize(device)
        barrier()
        t0 = time.perf_counter()
        sim.restore(host_ckpt)                # host -> device copy of every per-agent vector
        h2d = sum(v.numel() * v.element_size() for v in pinned.values())
        d2h = 0
        for _ in range(a.steps):
            sim.step(SUBSTEPS_PER_DAY)
            rows = sim.read_log()             # device -> host read of the day's counters (synchronises)
            d2h += 16 * 8
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
        e2e = dict(value=n_total * a.steps / wall, unit=UNIT, h2d_bytes_per_step=h2d / a.steps,
                   d2h_bytes_per_step=d2h / a.steps, seconds=wall,
                   note='K more days from the state the timed region ended in, through Simulation (C ABI) with the agent store in '
                        'page-locked host memory: host->device copy of all per-agent vectors, K days with a log read-back '
                        '(device->host, synchronising) per day, device->host copy of the final state word; handle / '
                        'communicator creation is set-up and not timed')

    # ---- the whole scenario of BASELINE.json configs[1]: 365 simulated days from the seeding, device resident
    full_run = None
    if a.full_run_days > 0:
        sim.close()
        del sim
        sim = make_sim()
        sim.seed_infections(synth.pick_seeds(pop, spec, infect_num=max(1, int(0.00015 * pop.n)), seed=POP_SEED + rank) + base)
        sim.step(SUBSTEPS_PER_DAY)            # first day: plain launches + graph capture, untimed
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(sim.stream)
        sim.step((a.full_run_days - 1) * SUBSTEPS_PER_DAY)
        f1.record(sim.stream)
        barrier()
        tf = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(tf, op=torch.distributed.ReduceOp.MAX)
        fc = sim.counters()
        full_run = dict(days=a.full_run_days - 1, seconds=float(tf.item()) / 1e3,
                        agent_days_per_sec=n_total * (a.full_run_days - 1) / (float(tf.item()) / 1e3),
                        seed_fraction=0.00015, infected_at_end_this_rank=fc['infected_count'], died_at_end_this_rank=fc['died_count'])

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline_sample(a.cpu_agents, a.districts, a.seed_fraction)

    if rank == 0:
        hot_mb = 8.0 * sim.sm.n_slots / 1e6
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
                        l2=f'per-GPU agent store {store.nbytes / 1e6:.0f} MB > 126 MB L2; the 8 B/agent hot record swept '
                           f'every sub-step is {hot_mb:.0f} MB (no explicit flush: re-sweeping it is the workload)',
                        state_at_timing={k: prevalence[k] for k in ('current_exposed_cases', 'current_contagious_cases',
                                                                    'infected_count', 'recovered_count')},
                        timing_regime='the K timed days start after the spin-up, around the epidemic peak: the most expensive '
                                      'days of the run; full_run is the same population over the whole scenario',
                        full_run=full_run),
            roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=launches, clocks=clocks.summary())
        print(json.dumps(line), flush=True)
    _teardown(sim, world)


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
