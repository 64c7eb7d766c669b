"""Ingestion of a census on the device: integer columns in PERSON order -> the packed agent store in HBM.

The reference builds its vectors with string-keyed dictionary maps over pandas columns (Country.load_agents /
initialize_agent_vectors, /root/reference/src/covid19_abm/base_model.py:572-634, 760-799); `engine.HostStore` is the
host (numpy) form of the layout work that follows it -- canonical (district, household) order, sub-tile packing, bit
packing, household index -- and takes longer at 15 M agents than the 365-day run it prepares.  `DeviceStore` does the same
work where the store is going to live: the raw integer columns are uploaded once (a few hundred MB over PCIe), sorted,
gathered and packed with device-wide primitives (torch: sort, scan, gather/scatter -- plumbing, not the hot path); only
the greedy sub-tile chain (one hop per 128-slot sub-tile) runs on the host, over a 4-byte-per-household array.  The result
is bit-identical to HostStore (tests/test_ingest.py compares every array on the CPU device; tests/test_gpu_ingest.py on the
GPU) and plugs into `Simulation(host_store=...)` without an upload.

The canonical-order columns come back to the host as one `Population` (the drop-in `Country` keeps it for refresh / state
read-back); at PCIe speed that is tens of milliseconds, against seconds for the same gathers in numpy.
"""
import numpy as np

from . import constants as C
from . import layout, tables as tables_mod
from .spec import ModelSpec, Population

SUBS_PER_TILE = C.TILE // C.SUBTILE


def _dev(device):
    import torch
    if device is None:
        return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
    return torch.device(device) if not isinstance(device, int) else torch.device('cuda', device)


class DeviceStore:
    """Same attributes as engine.HostStore (spec, pop, sm, host, move_class, move_thr, agent_id_base, agent_ids, nbytes);
    `host` holds DEVICE tensors, so `Simulation` binds them as they are.  `order[i]` = person index of canonical agent i."""

    def __init__(self, spec: ModelSpec, district, household, age, status, econ_kind, district_mover, move_prob_weekday,
                 move_prob_other, econ_pool=None, device=None, agent_id_base=0, presorted=False):
        import torch
        dev = _dev(device)
        self.spec, self.agent_id_base, self.agent_ids, self.device = spec, int(agent_id_base), None, dev
        up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dev, non_blocking=True).to(dt)
        d_p, h_p = up(district, torch.int64), up(household, torch.int64)
        n = int(d_p.shape[0])
        if n == 0:
            raise ValueError('empty population')
        # ---- canonical order: (home district, household), ties in person order -- one stable device sort
        if presorted:
            order = torch.arange(n, device=dev)
            key = d_p * (int(h_p.max().item()) + 1) + h_p
        else:
            span = int(h_p.max().item()) + 1
            if int(d_p.max().item()) >= (2 ** 62) // span:
                raise ValueError('district x household key overflows 64 bits')
            key, order = torch.sort(d_p * span + h_p, stable=True)
        g = lambda a, dt: up(a, dt)[order]
        district_c = d_p[order]
        age_c, status_c = g(age, torch.int32), g(status, torch.int32)
        kind_c, mover_c = g(econ_kind, torch.int32), g(district_mover, torch.int32)
        pw_c, po_c = g(move_prob_weekday, torch.float64), g(move_prob_other, torch.float64)
        pool_c = None if econ_pool is None else g(econ_pool, torch.int32)
        # ---- households: heads, dense index, sizes
        head = torch.ones(n, dtype=torch.bool, device=dev)
        head[1:] = key[1:] != key[:-1]
        dense = torch.cumsum(head.to(torch.int64), 0) - 1
        starts = torch.nonzero(head).reshape(-1)
        sizes = torch.diff(torch.cat([starts, torch.tensor([n], device=dev)]))
        n_households = int(starts.shape[0])
        # a household must not span two home districts (the layout is household-aligned)
        hh_sorted_unique = int(torch.unique(h_p).shape[0]) if not presorted else n_households
        if hh_sorted_unique != n_households:
            raise NotImplementedError('a household has members with different home districts: the household-aligned device '
                                      'layout cannot hold it')
        # ---- greedy sub-tile packing (layout.build_slot_map, contiguous ids): units = whole small households, single
        #      members of big ones
        big = sizes > C.BIG_HOUSEHOLD
        unit_size = torch.where(big, torch.ones_like(sizes), sizes)
        unit_rep = torch.where(big, sizes, torch.ones_like(sizes))
        unit_len = torch.repeat_interleave(unit_size, unit_rep)
        unit_end = torch.cumsum(unit_len, 0)
        unit_start = unit_end - unit_len
        m = int(unit_start.shape[0])
        d_unit = district_c[unit_start]
        first_of_run = torch.ones(m, dtype=torch.bool, device=dev)
        first_of_run[1:] = d_unit[1:] != d_unit[:-1]
        brk = torch.nonzero(first_of_run).reshape(-1)
        brk_next = torch.cat([brk[1:], torch.tensor([m], device=dev)])
        next_brk = brk_next[torch.searchsorted(brk, torch.arange(m, device=dev), right=True) - 1]
        lead_unit = (unit_start + self.agent_id_base) & 3
        nxt = torch.minimum(torch.searchsorted(unit_end, unit_start + (C.SUBTILE - lead_unit), right=True), next_brk)
        if bool((nxt <= torch.arange(m, device=dev)).any()):
            raise RuntimeError('placement unit larger than a sub-tile')
        nxt_l = nxt.to(torch.int32).cpu().numpy().tolist()       # the one sequential piece: one hop per sub-tile
        chain, i = [0], 0
        while i < m:
            i = nxt_l[i]
            chain.append(i)
        first_unit = torch.tensor(chain[:-1], dtype=torch.int64, device=dev)
        n_sub_used = int(first_unit.shape[0])
        sub_first_used = unit_start[first_unit]
        sub_lead_used = lead_unit[first_unit]
        sub_home_used = d_unit[first_unit]
        n_tiles = (n_sub_used + SUBS_PER_TILE - 1) // SUBS_PER_TILE
        pad = n_tiles * SUBS_PER_TILE - n_sub_used
        sub_first_all = torch.cat([sub_first_used, torch.tensor([n], device=dev)])
        counts = torch.diff(sub_first_all)
        if bool((counts + sub_lead_used > C.SUBTILE).any()):
            raise RuntimeError('sub-tile overflow')
        shift = torch.arange(n_sub_used, device=dev) * C.SUBTILE + sub_lead_used - sub_first_used
        slot_of = torch.arange(n, device=dev) + torch.repeat_interleave(shift, counts)
        n_slots = n_tiles * C.TILE
        end_id = n + self.agent_id_base
        # ---- the hot words
        big_agent = torch.repeat_interleave(big, sizes)
        f = torch.full((n,), C.LOC_HOME << C.F_LOC_SHIFT, dtype=torch.int64, device=dev)
        f |= (mover_c.to(torch.int64) & 1) * C.F_DMOVER
        f |= status_c.to(torch.int64) << C.F_STATUS_SHIFT
        f |= age_c.to(torch.int64) << C.F_AGE_SHIFT
        f |= head.to(torch.int64) * C.F_HEAD
        f |= kind_c.to(torch.int64) << C.F_EKIND_SHIFT
        f |= big_agent.to(torch.int64) * C.F_BIGHH
        filler = (C.F_FILLER | C.F_HEAD | (C.LOC_DEAD << C.F_LOC_SHIFT) | (C.STATE_DEAD << C.F_EPI_SHIFT)
                  | (C.CLINICAL_RELEASED_OR_DEAD << C.F_CLIN_SHIFT))
        flags64 = torch.full((n_slots,), int(filler), dtype=torch.int64, device=dev)
        flags64[slot_of] = f
        never = int(C.FIRE_NEVER) << C.AUX_FIRE_SHIFT
        aux64 = torch.full((n_slots,), never, dtype=torch.int64, device=dev)
        aux64[slot_of] = district_c | never
        to_i32 = lambda t: ((t + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)      # uint32 bit pattern in an int32 tensor
        flags, aux = to_i32(flags64), to_i32(aux64)
        # household index of every slot within its sub-tile (fillers count as heads)
        headbit = ((flags64 & C.F_HEAD) != 0).reshape(-1, C.SUBTILE)
        hh_index = ((torch.cumsum(headbit.to(torch.int32), 1) - 1) & 0xFF).to(torch.uint8).reshape(-1)
        # ---- movement classes: few distinct values per column -> codes of the pairs
        uw, iw = torch.unique(pw_c, return_inverse=True)
        uo, io = torch.unique(po_c, return_inverse=True)
        codes, inv = torch.unique(iw * int(uo.shape[0]) + io, return_inverse=True)
        uniq = np.stack([uw[codes // int(uo.shape[0])].cpu().numpy(), uo[codes % int(uo.shape[0])].cpu().numpy()], axis=1)
        if uniq.shape[0] > 65535:
            raise ValueError('more than 65535 distinct movement-probability pairs')
        mild = spec.mild_symptom_move_prob
        bt = tables_mod.bern_threshold
        self.move_thr = np.ascontiguousarray(np.stack([bt(uniq[:, 0] * 1.0), bt(uniq[:, 0] * mild), bt(uniq[:, 1] * 1.0),
                                                       bt(uniq[:, 1] * mild)], axis=1))
        mc_slots = torch.zeros(n_slots, dtype=torch.int16, device=dev)
        mc_slots[slot_of] = inv.to(torch.int16)
        cold = torch.zeros((n_slots, 8), dtype=torch.int32, device=dev)
        cold[:, :5] = int(C.T_NEVER)
        self.host = dict(flags=flags, aux=aux, cold=cold, move_class=mc_slots, hh_index=hh_index)
        if pool_c is not None:
            ep = torch.zeros(n_slots, dtype=torch.int32, device=dev)
            ep[slot_of] = pool_c
            self.host['econ_pool'] = ep
        # ---- what the host keeps: the canonical-order columns (one Population), the permutation, the slot map
        cpu = lambda t: t.cpu().numpy()
        self.order = cpu(order)
        self.pop = Population(district=cpu(district_c).astype(np.int32), household=cpu(dense), age=cpu(age_c),
                              status=cpu(status_c), econ_kind=cpu(kind_c).astype(np.int8), district_mover=cpu(mover_c).astype(np.int8),
                              move_prob_weekday=cpu(pw_c), move_prob_other=cpu(po_c),
                              econ_pool=None if pool_c is None else cpu(pool_c))
        self.move_class = cpu(inv).astype(np.uint16)
        sub_first = np.r_[cpu(sub_first_used), np.full(pad + 1, n, dtype=np.int64)].astype(np.int64)
        sub_lead = np.r_[cpu(sub_lead_used), np.full(pad, end_id & 3, dtype=np.int64)].astype(np.int64)
        sub_home = np.r_[cpu(sub_home_used), np.zeros(pad, dtype=np.int64)].astype(np.int64)
        big_start = np.append(cpu(starts[big]), n).astype(np.int64) + self.agent_id_base
        self.sm = layout.SlotMap(n_agents=n, n_tiles=n_tiles, n_subtiles=n_tiles * SUBS_PER_TILE, sub_first=sub_first,
                                 sub_lead=sub_lead, sub_home=sub_home, slot_of=cpu(slot_of), big_start=big_start,
                                 big_size=cpu(sizes[big]).astype(np.int64), agent_id_base=self.agent_id_base, sub_id=None)

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.host.values())
