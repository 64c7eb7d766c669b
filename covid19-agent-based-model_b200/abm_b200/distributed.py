"""Multi-GPU plumbing: one process per GPU, torch.distributed for the rendezvous.

The data path has exactly one exchange per sub-step -- the all-reduce of the district x status tables
issued by the library itself on its stream (ncclAllReduce, abm_capi.cu) -- so all Python has to do is
hand every rank the same ncclUniqueId and sum the per-rank log counters when they are read.
"""
import ctypes as ct

import numpy as np


def init_from_env():
    """(rank, world_size, local_rank); initialises torch.distributed (NCCL) when launched by torchrun."""
    import os
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, world, local


def exchange_unique_id(lib, rank, world_size, device):
    """ncclUniqueId bytes created on rank 0 and broadcast over the existing process group."""
    import torch
    import torch.distributed as dist
    buf = (ct.c_char * 128)()
    if rank == 0:
        rc = lib.abm_nccl_unique_id(ct.cast(buf, ct.c_void_p))
        if rc != 0:
            raise RuntimeError(f'abm_nccl_unique_id failed: {rc}')
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def sum_over_ranks(rows: np.ndarray, device):
    """Sum an int64 array over ranks (log counters are per-rank partial sums)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return rows
    t = torch.from_numpy(np.ascontiguousarray(rows)).to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()


def all_gather_bytes(payload: bytes, device):
    """Every rank's `payload` (same length everywhere), in rank order, over the existing process group."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError('torch.distributed is not initialised: the peer exchange needs it to trade buffer handles')
    world = dist.get_world_size()
    dev = device if dist.get_backend() == 'nccl' else 'cpu'
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]
