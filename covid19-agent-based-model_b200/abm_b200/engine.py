"""ctypes binding of libabm_b200.so and the `Simulation` object that owns one rank's agent store.

PyTorch is used only to own device buffers (and, multi-GPU, to rendezvous); every computation of the
agent step happens in the CUDA library.  There is no CPU fallback: constructing a `Simulation` without
the built library or without a CUDA device raises.

Role in the reference: this is what `Country` + `EpidemicScheduler` are on the host
(/root/reference/src/covid19_abm/base_model.py:20-441, 444-634): the vectors, the clock, `step()`.
"""
import ctypes as ct
import os
from datetime import timedelta

import numpy as np

from . import constants as C
from . import layout, tables as tables_mod
from .spec import ModelSpec, Population

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ABM_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'libabm_b200.so')

COUNTER_NAMES = ['current_contagious_count', 'infected_count', 'hospitalized_count', 'critical_count', 'died_count',
                 'recovered_count', 'current_exposed_cases', 'current_contagious_cases', 'current_hospitalized_cases',
                 'current_critical_cases', 'asymptomatic_count', 'symptomatic_count']
LOG_STRIDE = 16
LOG_TICK = 12


class AbmConfig(ct.Structure):
    _fields_ = [('n_slots', ct.c_int64), ('n_tiles', ct.c_int64), ('agent_id_base', ct.c_int64),
                ('n_districts', ct.c_int32), ('n_status', ct.c_int32), ('n_pools', ct.c_int32), ('hw_rows', ct.c_int32),
                ('n_move_classes', ct.c_int32), ('n_big_households', ct.c_int32), ('step_ticks', ct.c_int32),
                ('start_minute_of_day', ct.c_int32), ('start_weekday', ct.c_int32), ('mild_active', ct.c_int32),
                ('max_log_rows', ct.c_int32), ('device', ct.c_int32), ('seed', ct.c_uint64),
                ('symptomatic_rate', ct.c_double), ('critical_fatality_rate', ct.c_double), ('stream', ct.c_void_p),
                ('use_graph', ct.c_int32), ('use_pdl', ct.c_int32), ('world_size', ct.c_int32), ('rank', ct.c_int32), ('nccl_unique_id', ct.c_void_p),
                ('exchange', ct.c_int32), ('reserved', ct.c_int32)]


class AbmTables(ct.Structure):
    _fields_ = [(n, ct.c_void_p) for n in
                ('rfx', 'qfx', 'W', 'pool_hw', 'od_thr', 'stay_par', 'dwell', 'znorm', 'p_hosp', 'p_crit', 'p_hfat',
                 'severe_risk', 'move_thr', 'sub_info', 'big_start', 'omexp')]


class AbmAgentPtrs(ct.Structure):
    _fields_ = [(n, ct.c_void_p) for n in ('flags', 'aux', 'cold', 'move_class', 'econ_pool', 'hh_index')]


# words of an agent's cold record (include/abm_b200.h ABM_COLD_*): one 32-byte sector per agent
COLD_WORDS = 8
COLD_FIELDS = ('t_infected', 't_contagious', 't_onset', 't_recovered', 't_return', 'inf_at')


EXPORTS = ['abm_create', 'abm_set_tables', 'abm_bind_agents', 'abm_refresh', 'abm_seed', 'abm_step', 'abm_flush',
           'abm_read_log', 'abm_counters', 'abm_active_infections', 'abm_district_symptomatic', 'abm_debug_tables', 'abm_set_profiling',
           'abm_kernel_time', 'abm_table_time', 'abm_event_time', 'abm_set_trace', 'abm_read_trace', 'abm_get_state', 'abm_set_state', 'abm_current_step', 'abm_launch_count', 'abm_sync', 'abm_last_error', 'abm_destroy',
           'abm_version', 'abm_nccl_unique_id', 'abm_peer_export', 'abm_peer_attach', 'abm_peer_detach']

_lib = None


def load_library(path=LIB_PATH):
    """dlopen the C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f'{path} not found: build it with `python __graft_entry__.py build` '
                           '(the agent step has no CPU fallback)')
    lib = ct.CDLL(path)
    for name in EXPORTS:
        getattr(lib, name)
    vp = ct.c_void_p
    lib.abm_create.argtypes = [ct.POINTER(AbmConfig), ct.POINTER(vp)]
    lib.abm_set_tables.argtypes = [vp, ct.POINTER(AbmTables)]
    lib.abm_bind_agents.argtypes = [vp, ct.POINTER(AbmAgentPtrs)]
    lib.abm_refresh.argtypes = [vp, vp, ct.c_int32]
    lib.abm_seed.argtypes = [vp, vp, vp, ct.c_int64]
    lib.abm_step.argtypes = [vp, ct.c_int32]
    lib.abm_flush.argtypes = [vp]
    lib.abm_read_log.argtypes = [vp, vp, ct.c_int32, ct.POINTER(ct.c_int32)]
    lib.abm_counters.argtypes = [vp, vp]
    lib.abm_active_infections.argtypes = [vp, ct.POINTER(ct.c_int64)]
    lib.abm_district_symptomatic.argtypes = [vp, vp, vp]
    lib.abm_debug_tables.argtypes = [vp, vp, vp, vp]
    lib.abm_set_profiling.argtypes = [vp, ct.c_int32]
    lib.abm_kernel_time.argtypes = [vp, ct.POINTER(ct.c_double), ct.POINTER(ct.c_int64)]
    lib.abm_table_time.argtypes = [vp, ct.POINTER(ct.c_double)]
    lib.abm_event_time.argtypes = [vp, ct.POINTER(ct.c_double)]
    lib.abm_set_trace.argtypes = [vp, ct.c_int32]
    lib.abm_read_trace.argtypes = [vp, vp, ct.c_int32]
    lib.abm_get_state.argtypes = [vp, ct.POINTER(ct.c_int64), vp, ct.POINTER(ct.c_int64)]
    lib.abm_set_state.argtypes = [vp, ct.c_int64, vp, ct.c_int64]
    lib.abm_current_step.argtypes = [vp]
    lib.abm_current_step.restype = ct.c_int64
    lib.abm_launch_count.argtypes = [vp]
    lib.abm_launch_count.restype = ct.c_int64
    lib.abm_sync.argtypes = [vp]
    lib.abm_last_error.argtypes = [vp]
    lib.abm_last_error.restype = ct.c_char_p
    lib.abm_destroy.argtypes = [vp]
    lib.abm_destroy.restype = None
    lib.abm_nccl_unique_id.argtypes = [vp]
    lib.abm_peer_export.argtypes = [vp, vp]
    lib.abm_peer_attach.argtypes = [vp, vp, ct.c_int32]
    lib.abm_peer_detach.argtypes = [vp]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(ct.c_void_p)


class HostStore:
    """The agent store of one rank packed on the HOST (page-locked torch tensors), ready to upload.

    This is what Country.load_agents / initialize_*_vectors (base_model.py:572-634, 760-799) produce: the
    per-agent vectors in device layout.  Building it (slot map, bit packing) is ingestion work; uploading
    it is a plain host -> device copy of `nbytes` bytes."""

    def __init__(self, spec: ModelSpec, pop: Population, agent_id_base=0, agent_ids=None):
        import torch
        self.spec, self.pop, self.agent_id_base = spec, pop, int(agent_id_base if agent_ids is None else agent_ids[0])
        self.agent_ids = None if agent_ids is None else np.ascontiguousarray(agent_ids, dtype=np.int64)
        self.sm = layout.build_slot_map(pop.household, pop.district, agent_id_base, agent_ids=self.agent_ids)
        sm = self.sm
        self.move_class, self.move_thr, _ = tables_mod.move_classes(pop, spec.mild_symptom_move_prob)
        flags, aux = layout.pack_flags(pop, sm)
        n_slots = sm.n_slots

        def slots_from_agents(values, fill, dtype):
            out = np.full(n_slots, fill, dtype=dtype)
            out[sm.slot_of] = values
            return out

        cold = np.zeros((n_slots, COLD_WORDS), np.int32)
        cold[:, :5] = np.int32(C.T_NEVER)      # the five dates are datetime.max; inf_at and the padding are 0
        arrays = dict(
            flags=flags.view(np.int32), aux=aux.view(np.int32), cold=cold,
            move_class=slots_from_agents(self.move_class, 0, np.uint16).view(np.int16),
            hh_index=layout.household_index(flags),
        )
        if pop.econ_pool is not None:
            arrays['econ_pool'] = slots_from_agents(pop.econ_pool.astype(np.int32), 0, np.int32)
        pin = torch.cuda.is_available()
        self.host = {}
        for k, a in arrays.items():
            t = torch.from_numpy(np.ascontiguousarray(a))
            self.host[k] = t.pin_memory() if pin else t

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.host.values())


class Simulation:
    """One rank's share of the population on one GPU.

    pop            canonical-order agents of this rank (household-aligned shard)
    agent_id_base  canonical id of pop's first agent (Philox key; makes results independent of sharding)
    agent_ids      canonical id of every agent of pop, for shards that are not one contiguous range of the canonical
                   order (sharding.deal_population); overrides agent_id_base
    district_start global canonical id of the first agent of every home district [D + 1]
    """

    def __init__(self, spec: ModelSpec, pop: Population, seed=0, device=0, tables=None, max_days=400,
                 agent_id_base=0, district_start=None, world_size=1, rank=0, nccl_unique_id=None, use_graph=True,
                 host_store=None, upload=True, use_pdl=True, exchange=None, agent_ids=None):
        """exchange (world_size > 1): 'p2p' sums the district x status tables over ranks inside the table kernel through
        peer memory (one node; needs an initialised torch.distributed group to trade the buffer handles), 'nccl' uses
        ncclAllReduce (needs `nccl_unique_id`).  Default: 'p2p' unless a unique id is given."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('abm_b200 needs a CUDA device (no CPU fallback)')
        self.torch = torch
        self.lib = load_library()
        self.spec, self.pop = spec, pop
        self.tab = tables if tables is not None else tables_mod.build_tables(spec)
        self.device = torch.device('cuda', device)
        torch.cuda.set_device(self.device)
        # all device work of this simulation (uploads, kernels, graph replays) is ordered on one stream
        self.stream = torch.cuda.Stream(device=self.device)
        hs = host_store if host_store is not None else HostStore(spec, pop, agent_id_base, agent_ids=agent_ids)
        if agent_ids is None and hs.agent_ids is not None:
            agent_ids = hs.agent_ids
        if agent_ids is not None:
            agent_id_base = int(agent_ids[0])
        if hs.pop is not pop or hs.agent_id_base != int(agent_id_base):
            raise ValueError('host_store was built for another population / id base')
        self.agent_ids = hs.agent_ids
        self.host_store = hs
        self.sm = hs.sm
        sm = self.sm
        D = spec.n_districts
        if district_start is None:
            district_start = np.searchsorted(pop.district, np.arange(D + 1)).astype(np.int64) + agent_id_base
        self.district_start = np.ascontiguousarray(district_start, dtype=np.int64)
        self.move_class, self.move_thr = hs.move_class, hs.move_thr
        n_slots = sm.n_slots
        with torch.cuda.stream(self.stream):      # asynchronous copies from page-locked memory
            if upload:
                self.buf = {k: t.to(self.device, non_blocking=True) for k, t in hs.host.items()}
            else:                                 # the caller restores a checkpoint into the buffers
                self.buf = {k: torch.empty(t.shape, dtype=t.dtype, device=self.device) for k, t in hs.host.items()}

        cfg = AbmConfig()
        cfg.n_slots, cfg.n_tiles, cfg.agent_id_base = n_slots, sm.n_tiles, agent_id_base
        cfg.n_districts, cfg.n_status, cfg.n_pools, cfg.hw_rows = D, spec.n_status, spec.n_pools, self.tab['hw_rows']
        cfg.n_move_classes, cfg.n_big_households = self.move_thr.shape[0], sm.big_size.shape[0]
        cfg.step_ticks = self.tab['step_ticks']
        cfg.start_minute_of_day, cfg.start_weekday = self.tab['start_minute_of_day'], self.tab['start_weekday']
        cfg.mild_active = int(self.tab['mild_active'])
        self.steps_per_day = (24 * 60) // cfg.step_ticks
        cfg.max_log_rows = int(max_days) + 2
        cfg.device = device
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.symptomatic_rate, cfg.critical_fatality_rate = self.tab['symptomatic_rate'], self.tab['critical_fatality_rate']
        cfg.stream = self.stream.cuda_stream
        cfg.use_graph = int(bool(use_graph))
        cfg.use_pdl = int(bool(use_pdl))
        cfg.world_size, cfg.rank = world_size, rank
        self._uid = None
        if exchange is None:
            exchange = 'nccl' if nccl_unique_id is not None else 'p2p'
        if exchange not in ('p2p', 'nccl'):
            raise ValueError(f'unknown exchange {exchange!r}')
        self.exchange = exchange if world_size > 1 else None
        if world_size > 1 and exchange == 'nccl':
            self._uid = (ct.c_char * 128).from_buffer_copy(bytes(nccl_unique_id))
            cfg.nccl_unique_id = ct.cast(self._uid, ct.c_void_p)
        cfg.exchange = 1 if (world_size > 1 and exchange == 'p2p') else 0
        self.h = ct.c_void_p()
        self._check(self.lib.abm_create(ct.byref(cfg), ct.byref(self.h)))
        self.cfg = cfg
        if cfg.exchange == 1:
            from . import distributed
            mine = (ct.c_char * 64)()
            self._check(self.lib.abm_peer_export(self.h, ct.cast(mine, ct.c_void_p)))
            handles = distributed.all_gather_bytes(bytes(mine), self.device)
            self._peer_handles = (ct.c_char * (64 * world_size)).from_buffer_copy(b''.join(handles))
            self._check(self.lib.abm_peer_attach(self.h, ct.cast(self._peer_handles, ct.c_void_p), world_size))

        t = self.tab
        keep = dict(
            rfx=t['rfx'], qfx=t['qfx'], W=t['W'], pool_hw=t['pool_hw'], od_thr=t['od_thr'], stay_par=t['stay_par'],
            dwell=t['dwell'], znorm=t['znorm'], p_hosp=t['p_hosp'], p_crit=t['p_crit'], p_hfat=t['p_hfat'],
            severe_risk=t['severe_risk'], move_thr=self.move_thr, sub_info=np.ascontiguousarray(sm.sub_info),
            big_start=np.ascontiguousarray(sm.big_start), omexp=t['omexp'])
        at = AbmTables()
        for name, arr in keep.items():
            setattr(at, name, _ptr(arr))
        self._check(self.lib.abm_set_tables(self.h, ct.byref(at)))
        ap = AbmAgentPtrs()
        for name in ('flags', 'aux', 'cold', 'move_class', 'hh_index'):
            setattr(ap, name, self.buf[name].data_ptr())
        ap.econ_pool = self.buf['econ_pool'].data_ptr() if 'econ_pool' in self.buf else None
        self._check(self.lib.abm_bind_agents(self.h, ct.byref(ap)))

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.abm_last_error(self.h) if self.h else b'create failed'
            raise RuntimeError(f'abm_b200 error {rc}: {msg.decode()}')

    @property
    def n_agents(self):
        return self.pop.n

    @property
    def step_index(self):
        return int(self.lib.abm_current_step(self.h))

    @property
    def real_time(self):
        return self.spec.start_datetime + timedelta(minutes=self.step_index * self.cfg.step_ticks)

    # ------------------------------------------------------------------ the path
    def seed_infections(self, agent_ids):
        """Country.set_epidemic_status(ids) at the current time (base_model.py:796)."""
        ids = np.ascontiguousarray(agent_ids, dtype=np.int64)
        if self.agent_ids is None:
            local = ids - self.sm.agent_id_base
        else:                                  # shard made of several id runs: look the ids up
            local = np.searchsorted(self.agent_ids, ids)
            if ids.size and (local.max() >= self.agent_ids.shape[0] or not np.array_equal(self.agent_ids[local], ids)):
                raise ValueError('seed ids outside this rank\'s shard')
        slots = np.ascontiguousarray(self.sm.slot_of[local])
        self._check(self.lib.abm_seed(self.h, _ptr(slots), _ptr(ids), ids.shape[0]))

    def step(self, n_substeps=1):
        """n x EpidemicScheduler.step() (base_model.py:33-179); asynchronous."""
        self._check(self.lib.abm_step(self.h, int(n_substeps)))

    def flush(self):
        self._check(self.lib.abm_flush(self.h))

    def sync(self):
        self._check(self.lib.abm_sync(self.h))

    def refresh(self):
        """Re-encode the vectors scenario hooks mutate on the host (district_mover, movement probabilities,
        economic_activity_location_ids; base_model.py:642-731) and push them to the device."""
        torch = self.torch
        self.flush()
        self.sync()
        flags = self.buf['flags'].cpu().numpy().view(np.uint32)
        layout.update_static_fields(flags, self.pop, self.sm)
        self.buf['flags'].copy_(torch.from_numpy(flags.view(np.int32)))
        self.move_class, self.move_thr, _ = tables_mod.move_classes(self.pop, self.spec.mild_symptom_move_prob)
        mc = np.zeros(self.sm.n_slots, dtype=np.uint16)
        mc[self.sm.slot_of] = self.move_class
        self.buf['move_class'].copy_(torch.from_numpy(mc.view(np.int16)))
        if self.pop.econ_pool is not None and 'econ_pool' in self.buf:
            ep = np.zeros(self.sm.n_slots, dtype=np.int32)
            ep[self.sm.slot_of] = self.pop.econ_pool
            self.buf['econ_pool'].copy_(torch.from_numpy(ep))
        torch.cuda.synchronize(self.device)
        self._check(self.lib.abm_refresh(self.h, _ptr(self.move_thr), self.move_thr.shape[0]))

    # ------------------------------------------------------------------ read-back
    def read_log(self):
        """Rows of Country.log_model_output produced so far: list of dicts (this rank's agents)."""
        rows = np.zeros((self.cfg.max_log_rows, LOG_STRIDE), dtype=np.int64)
        n = ct.c_int32()
        self._check(self.lib.abm_read_log(self.h, _ptr(rows), self.cfg.max_log_rows, ct.byref(n)))
        return rows[:n.value]

    def log_dicts(self, rows=None):
        rows = self.read_log() if rows is None else rows
        out = []
        for r in rows:
            d = {name: int(r[i]) for i, name in enumerate(COUNTER_NAMES)}
            d['tick'] = int(r[LOG_TICK])
            d['date'] = (self.spec.start_datetime + timedelta(minutes=int(r[LOG_TICK]))).isoformat()
            out.append(d)
        return out

    def counters(self):
        out = np.zeros(12, dtype=np.int64)
        self._check(self.lib.abm_counters(self.h, _ptr(out)))
        return {name: int(out[i]) for i, name in enumerate(COUNTER_NAMES)}

    def active_infections(self):
        """Exposed + contagious agents of this rank from the running device counters: run_scenario's stop test
        (scenario_models.py:494-495) without a flush and without a kernel launch."""
        n = ct.c_int64()
        self._check(self.lib.abm_active_infections(self.h, ct.byref(n)))
        return int(n.value)

    def district_symptomatic(self):
        D = self.spec.n_districts
        sym, pop = np.zeros(D, dtype=np.int64), np.zeros(D, dtype=np.int64)
        self._check(self.lib.abm_district_symptomatic(self.h, _ptr(sym), _ptr(pop)))
        return sym, pop

    def debug_tables(self):
        pa = self.spec.n_pools * self.spec.n_status
        r, n, t = np.zeros(pa, np.uint64), np.zeros(pa, np.uint64), np.zeros(pa, np.uint32)
        self._check(self.lib.abm_debug_tables(self.h, _ptr(r), _ptr(n), _ptr(t)))
        shp = (self.spec.n_pools, self.spec.n_status)
        return r.reshape(shp), n.reshape(shp), t.reshape(shp)

    def download(self):
        """Device SoA -> host numpy (slot order); applies pending infection draws first."""
        self.flush()
        self.sync()
        out = {}
        for k, v in self.buf.items():
            a = v.cpu().numpy()
            if k == 'cold':                    # split the record into the per-field vectors layout.unpack_state reads
                for j, name in enumerate(COLD_FIELDS):
                    out[name] = np.ascontiguousarray(a[:, j]).view(np.uint32 if name == 'inf_at' else np.int32)
                continue
            if k in ('flags', 'aux'):
                a = a.view(np.uint32)
            elif k == 'move_class':
                a = a.view(np.uint16)
            out[k] = a
        return out

    def state(self, household_ids=None, econ_loc=None):
        """Reference-style vectors over this rank's agents (see layout.unpack_state)."""
        return layout.unpack_state(self.download(), self.pop, self.sm, self.spec.n_districts,
                                   household_ids=household_ids, econ_loc=econ_loc)

    def checkpoint(self):
        """Everything needed to resume: the agent vectors (numpy, slot order), the clock, the running
        counters and the number of seeded infections.  Pending infection draws are applied first."""
        k, seeded = ct.c_int64(), ct.c_int64()
        counters = np.zeros(12, dtype=np.int64)
        self._check(self.lib.abm_get_state(self.h, ct.byref(k), _ptr(counters), ct.byref(seeded)))
        self.sync()
        arrays = {name: v.cpu().numpy() for name, v in self.buf.items()}
        return dict(arrays=arrays, k=int(k.value), counters=counters, seeded=int(seeded.value))

    def restore(self, ckpt, arrays_on_device=False):
        """Resume from `checkpoint()` output; `ckpt['arrays']` may hold numpy arrays or (pinned) torch tensors."""
        torch = self.torch
        if not arrays_on_device:
            with torch.cuda.stream(self.stream):
                for name, a in ckpt['arrays'].items():
                    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
                    self.buf[name].copy_(t, non_blocking=True)
        counters = np.ascontiguousarray(ckpt['counters'], dtype=np.int64)
        self._check(self.lib.abm_set_state(self.h, int(ckpt['k']), _ptr(counters), int(ckpt['seeded'])))

    def set_profiling(self, on):
        self._check(self.lib.abm_set_profiling(self.h, int(bool(on))))

    def kernel_time(self):
        ms, n = ct.c_double(), ct.c_int64()
        self._check(self.lib.abm_kernel_time(self.h, ct.byref(ms), ct.byref(n)))
        return ms.value, n.value

    def table_time(self):
        ms = ct.c_double()
        self._check(self.lib.abm_table_time(self.h, ct.byref(ms)))
        return ms.value

    def set_trace(self, n_substeps):
        """Device-side timeline of sub-steps 0..n-1 (abm_set_trace); 0 switches it off."""
        self._check(self.lib.abm_set_trace(self.h, int(n_substeps)))

    def read_trace(self, n_substeps):
        """uint64 [n, 3 kernels (sweep, events, table), 2 (first CTA start, last CTA end)] in ns of the GPU timer."""
        out = np.zeros((int(n_substeps), 12), dtype=np.uint64)
        self._check(self.lib.abm_read_trace(self.h, _ptr(out), int(n_substeps)))
        self.trace_debug = out[:, 6:]
        return np.ascontiguousarray(out[:, :6]).reshape(-1, 3, 2)

    def event_time(self):
        ms = ct.c_double()
        self._check(self.lib.abm_event_time(self.h, ct.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        return int(self.lib.abm_launch_count(self.h))

    def close(self):
        if getattr(self, 'h', None):
            if getattr(self, 'exchange', None) == 'p2p':
                # peers read this rank's exchange buffer: unmap everywhere before anybody frees
                import torch.distributed as dist
                self.lib.abm_sync(self.h)
                if dist.is_initialized():
                    dist.barrier()
                self.lib.abm_peer_detach(self.h)
                if dist.is_initialized():
                    dist.barrier()
            self.lib.abm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
