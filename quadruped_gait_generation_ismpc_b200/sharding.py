"""Instance-level data parallelism across the GPUs of one box.

Instances are independent, so the only multi-GPU logic is: contiguous shard per rank, no per-tick
communication, one gather of the result array at the end (torch.distributed: NCCL on GPUs, gloo in the
CPU tests).  One process per GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_records(local, n_total, device=None):
    """All-gather per-rank record arrays (numpy structured or plain, first axis = instances of this rank's
    shard) into the full n_total-long array on every rank.  Records travel as raw bytes."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    rec = local.dtype.itemsize * int(np.prod(local.shape[1:], dtype=np.int64))
    sizes = shard_sizes(n_total, world)
    mx = max(sizes)
    buf = torch.zeros(mx * rec, dtype=torch.uint8, device=device)
    raw = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1))
    buf[: raw.numel()] = raw.to(buf.device)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    parts = [o[: s * rec].cpu().numpy() for o, s in zip(outs, sizes)]
    full = np.concatenate(parts).view(local.dtype)
    return full.reshape((n_total,) + tuple(local.shape[1:]))


def max_over_ranks(value, device=None):
    """Max of a python float over ranks (timing is reported as the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
