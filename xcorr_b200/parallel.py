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


def gather_ccf(local_ccf, n_windows, dst=0, group=None, out=None, chunk_bytes=256 << 20):
    """Assemble ``[n_pairs, n_windows, n_lags]`` on rank ``dst`` from each rank's
    ``[n_pairs, len(shard_windows(...)), n_lags]`` block.  Returns the full tensor on ``dst`` and
    ``None`` elsewhere.

    The blocks are gathered in slabs of pairs of at most `chunk_bytes` per rank, so the staging memory is
    bounded whatever the CCF size (config 3: 3.9 GB per window and rank).  `out` (optional, on ``dst``)
    receives the result -- e.g. a page-locked host-mapped or preallocated device tensor; when it is None the
    full tensor is allocated on ``dst`` and a ValueError is raised if that exceeds the free device memory."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_windows(n_windows, rank, world)
    if local_ccf.shape[1] != len(mine):
        raise ValueError(f"rank {rank} holds {local_ccf.shape[1]} windows, expected {len(mine)}")
    n_pairs, _, n_lags = local_ccf.shape
    per = (n_windows + world - 1) // world            # padded so every rank sends the same shape
    if rank == dst and out is None:
        need = n_pairs * n_windows * n_lags * local_ccf.element_size()
        if local_ccf.is_cuda:
            free, _ = torch.cuda.mem_get_info(local_ccf.device)
            if need > free:
                raise ValueError(f"gather_ccf: the full CCF block needs {need} bytes, {free} are free on "
                                 f"{local_ccf.device}; pass `out` or gather fewer windows at a time")
        out = local_ccf.new_empty((n_pairs, n_windows, n_lags))
    slab = max(1, int(chunk_bytes // max(1, per * n_lags * local_ccf.element_size())))
    send = local_ccf.new_zeros((min(slab, n_pairs), per, n_lags))
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    for p0 in range(0, n_pairs, slab):
        p1 = min(n_pairs, p0 + slab)
        send[:p1 - p0, :len(mine)] = local_ccf[p0:p1]
        dist.gather(send, bufs, dst=dst, group=group)
        if rank == dst:
            for r in range(world):
                ws = shard_windows(n_windows, r, world)
                out[p0:p1, ws] = bufs[r][:p1 - p0, :len(ws)].to(out.device)
    return out if rank == dst else None


def reduce_stack(local_stack, group=None):
    """Sum the per-rank stacks ``[n_pairs, n_lags]`` (each rank stacked its own
    windows) in place on every rank."""
    dist.all_reduce(local_stack, op=dist.ReduceOp.SUM, group=group)
    return local_stack
