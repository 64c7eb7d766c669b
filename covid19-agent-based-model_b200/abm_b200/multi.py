"""Whole-population facade over one rank's `Simulation`: what lets the drop-in `Country` run unchanged on N GPUs.

One process per GPU (torchrun); every process executes the reference's host sequence -- ParamsConfig, scenario class,
load_agents, the scenario hooks -- on the FULL census with the same numpy seed, so the host vectors are identical on all
ranks.  `ShardedSimulation` then keeps only this rank's household-aligned range of the canonical order on its GPU
(draws are keyed by the global agent id, so the trajectory does not depend on the number of ranks) and presents the
interface `Country` uses for one GPU:

  seed_infections(global ids)   each rank infects the ids of its own range
  step / flush / sync           local; the district x status tables are summed over ranks inside the table kernel
  state()                       the per-agent vectors of ALL ranks, gathered on access (rank order = canonical order)
  read_log / counters / active_infections / district_symptomatic     per-rank partial sums added over ranks

Role in the reference: none -- the reference is one Python process per run (scenario_models.py:443-500); this is SURVEY
section 8(e) behind the reference's own API.
"""
import numpy as np

from . import sharding
from .engine import COUNTER_NAMES, Simulation


def world_size():
    try:
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    except Exception:
        return 1


def _device_for_collectives(local_device):
    import torch.distributed as dist
    return local_device if dist.get_backend() == 'nccl' else 'cpu'


def broadcast_int(value, local_device):
    """Rank 0's 63-bit integer on every rank (the Philox key must be the same everywhere)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(value)], dtype=torch.int64, device=_device_for_collectives(local_device))
    dist.broadcast(t, src=0)
    return int(t.item())


class ShardedSimulation:
    def __init__(self, spec, pop, seed=0, device=0, tables=None, max_days=400, upload=True, sim_class=None, exchange='p2p'):
        import torch
        import torch.distributed as dist
        if not upload:
            raise NotImplementedError('resuming a pickled model on several ranks is not supported: resume it on one GPU')
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.spec, self.full_pop = spec, pop
        self.bounds = sharding.split_population(pop, self.world)
        self.lo, self.hi = self.bounds[self.rank]
        self.pop = sharding.slice_population(pop, self.lo, self.hi)
        dstart = np.searchsorted(pop.district, np.arange(spec.n_districts + 1)).astype(np.int64)
        cls = sim_class or Simulation
        self.sim = cls(spec, self.pop, seed=seed, device=device, tables=tables, max_days=max_days, agent_id_base=self.lo,
                       district_start=dstart, world_size=self.world, rank=self.rank, exchange=exchange)
        self.local_device = torch.device('cuda', device) if torch.cuda.is_available() and sim_class is None else 'cpu'
        self.cdev = _device_for_collectives(self.local_device)

    # ------------------------------------------------------------------ collectives on small / large host arrays
    def _sum(self, arr):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.cdev)
        dist.all_reduce(t)
        return t.cpu().numpy()

    def _gather_concat(self, arr):
        """This rank's per-agent vector -> the vector over all agents (ranks hold consecutive canonical ranges)."""
        import torch
        import torch.distributed as dist
        sizes = [hi - lo for lo, hi in self.bounds]
        width = max(sizes)
        a = np.ascontiguousarray(arr)
        pad = np.zeros(width, dtype=a.dtype)
        pad[:a.shape[0]] = a
        mine = torch.from_numpy(pad).to(self.cdev)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine)
        return np.concatenate([o.cpu().numpy()[:n] for o, n in zip(out, sizes)])

    # ------------------------------------------------------------------ the Simulation interface Country uses
    @property
    def step_index(self):
        return self.sim.step_index

    @property
    def launch_count(self):
        return self.sim.launch_count

    def seed_infections(self, agent_ids):
        ids = np.asarray(agent_ids, dtype=np.int64)
        self.sim.seed_infections(ids[(ids >= self.lo) & (ids < self.hi)])

    def step(self, n_substeps=1):
        self.sim.step(n_substeps)

    def flush(self):
        self.sim.flush()

    def sync(self):
        self.sim.sync()

    def refresh(self):
        """Country mutated the scenario-controlled fields of the FULL population: re-slice them, then re-encode."""
        full, part, lo, hi = self.full_pop, self.pop, self.lo, self.hi
        part.district_mover = full.district_mover[lo:hi]
        part.move_prob_weekday = full.move_prob_weekday[lo:hi]
        part.move_prob_other = full.move_prob_other[lo:hi]
        part.econ_kind = full.econ_kind[lo:hi]
        if full.econ_pool is not None:
            part.econ_pool = full.econ_pool[lo:hi]
        self.sim.refresh()

    def state(self, household_ids=None, econ_loc=None):
        lo, hi = self.lo, self.hi
        local = self.sim.state(household_ids=None if household_ids is None else household_ids[lo:hi],
                               econ_loc=None if econ_loc is None else econ_loc[lo:hi])
        return {k: self._gather_concat(np.asarray(v)) for k, v in local.items()}

    def read_log(self):
        rows = self.sim.read_log().copy()
        n = len(COUNTER_NAMES)
        rows[:, :n] = self._sum(rows[:, :n])          # the tick column is the same on every rank
        return rows

    def log_dicts(self, rows=None):
        return self.sim.log_dicts(self.read_log() if rows is None else rows)

    def counters(self):
        c = self.sim.counters()
        tot = self._sum(np.array([c[k] for k in COUNTER_NAMES], dtype=np.int64))
        return {k: int(v) for k, v in zip(COUNTER_NAMES, tot)}

    def active_infections(self):
        return int(self._sum(np.array([self.sim.active_infections()], dtype=np.int64))[0])

    def district_symptomatic(self):
        sym, pop = self.sim.district_symptomatic()
        both = self._sum(np.stack([sym, pop]).astype(np.int64))
        return both[0], both[1]

    def debug_tables(self):
        return self.sim.debug_tables()

    def checkpoint(self):
        raise NotImplementedError('a multi-rank model is pickled without its device state (host vectors only)')

    def close(self):
        self.sim.close()
