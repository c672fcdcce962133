"""Pocket-sharded multi-GPU sampling: one process per GPU, no data-path collective.

The reference scales by launching one Python process per GPU and assigning test-set pockets round-robin
(reference scripts/batch_sample_diffusion.sh:15-21: task i runs on worker i % NODE_ALL).  Same partitioning here, on
`torch.distributed` (backend "nccl" over NVLink on GPUs; "gloo" in the CPU unit tests).  The only collective is one
broadcast of the packed weights at start-up (rank 0 loads the checkpoint, everyone else receives ~11 MB over NVLink)
plus an optional gather of per-rank timings; results are written per pocket exactly like the reference
(`result_{id}.pt`), so nothing flows between ranks while sampling."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun contract)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kw['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def shard_round_robin(items, rank, world):
    """The reference's partitioning: item i -> worker i % world (scripts/batch_sample_diffusion.sh:16-17)."""
    return [x for i, x in enumerate(items) if i % world == rank]


def shard_longest_first(costs, world):
    """Greedy longest-processing-time assignment of pockets to ranks (pockets differ in atom count; cost ~ nodes * k).
    Returns a list of index lists, one per rank; deterministic, identical on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(ix) for ix in out]


def broadcast_state_dict(module, src=0):
    """One flat broadcast of every parameter/buffer of `module` from rank `src` (the weight broadcast of SURVEY.md 8(e))."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return module
    tensors = list(module.state_dict().values())
    flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in tensors])
    dist.broadcast(flat, src)
    off = 0
    for t in tensors:
        n = t.numel()
        t.data.copy_(flat[off:off + n].view_as(t))
        off += n
    if hasattr(module, '_drop_engine'):
        module._drop_engine()          # weights changed: the libtdiff engine is rebuilt on next use
    return module


def max_over_ranks(value, device=None):
    """Max of a python float over ranks (device-timed sections are reported as the slowest rank's time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
