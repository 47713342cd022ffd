"""Sharding of the hot path over the GPUs of one box.

The path shards by *window*: every (window) is independent and all receivers
and pairs of a window live on one GPU, so each receiver-window is transformed
once and reused locally.  There is no collective on the data path;
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) is used
only to gather per-window CCFs to the rank that writes them, or to reduce
stacked CCFs.  The reference's unit of distribution is the same independent
(pair, day) task (``core/lazy.py:181-184``).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_windows", "owner_of_window", "gather_ccf", "reduce_stack"]


def shard_windows(n_windows, rank, world):
    """Window indices processed by ``rank``: round-robin ``w % world == rank``
    (365 days over 8 GPUs -> 46/45 windows, <= 2 % imbalance)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_windows, world))


def owner_of_window(w, world):
    return w % world


def gather_ccf(local_ccf, n_windows, dst=0, group=None):
    """Assemble ``[n_pairs, n_windows, n_lags]`` on rank ``dst`` from each
    rank's ``[n_pairs, len(shard_windows(...)), n_lags]`` block.  Returns the
    full tensor on ``dst`` and ``None`` elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_windows(n_windows, rank, world)
    if local_ccf.shape[1] != len(mine):
        raise ValueError(f"rank {rank} holds {local_ccf.shape[1]} windows, expected {len(mine)}")
    n_pairs, _, n_lags = local_ccf.shape
    per = (n_windows + world - 1) // world            # padded so every rank sends the same shape
    send = local_ccf.new_zeros((n_pairs, per, n_lags))
    send[:, :len(mine)] = local_ccf
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = local_ccf.new_empty((n_pairs, n_windows, n_lags))
    for r in range(world):
        ws = shard_windows(n_windows, r, world)
        out[:, ws] = bufs[r][:, :len(ws)]
    return out


def reduce_stack(local_stack, group=None):
    """Sum the per-rank stacks ``[n_pairs, n_lags]`` (each rank stacked its own
    windows) in place on every rank."""
    dist.all_reduce(local_stack, op=dist.ReduceOp.SUM, group=group)
    return local_stack
