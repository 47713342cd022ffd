"""World-size-2 gloo tests of the window sharding / gather / reduce logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xcorr_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_windows, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_pairs, n_lags = 3, 17
        full = torch.arange(n_pairs * n_windows * n_lags, dtype=torch.float32).reshape(n_pairs, n_windows, n_lags)
        mine = parallel.shard_windows(n_windows, rank, world)
        local = full[:, mine].clone()
        got = parallel.gather_ccf(local, n_windows, dst=0)
        ok_gather = (got is None) if rank != 0 else bool(torch.equal(got, full))
        # several slabs of pairs (staging bounded by chunk_bytes) into a caller-provided output
        out = torch.zeros_like(full) if rank == 0 else None
        got2 = parallel.gather_ccf(local, n_windows, dst=0, out=out, chunk_bytes=4 * 17 * 4 + 1)
        ok_gather = ok_gather and ((got2 is None) if rank != 0 else (got2 is out and bool(torch.equal(out, full))))
        stack = local.sum(dim=1)
        parallel.reduce_stack(stack)
        ok_reduce = bool(torch.allclose(stack, full.sum(dim=1)))
        out_q.put((rank, ok_gather, ok_reduce, mine))
    finally:
        dist.destroy_process_group()


def test_shard_windows_partition():
    for n, world in ((365, 8), (7, 2), (1, 4), (0, 2)):
        seen = sorted(w for r in range(world) for w in parallel.shard_windows(n, r, world))
        assert seen == list(range(n))
        sizes = [len(parallel.shard_windows(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
        for w in range(n):
            assert w in parallel.shard_windows(n, parallel.owner_of_window(w, world), world)


def test_gather_and_reduce_world2():
    world, n_windows = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_windows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_g, ok_r, mine in res:
        assert ok_g and ok_r, (rank, ok_g, ok_r)
        assert mine == list(range(rank, n_windows, world))


def test_timelapse_pairs_tile_over_ranks():
    """Config 5 sharding: the masked triangle of scripts/timelapse.py:348-360 split over ranks, the
    day transforms split in contiguous slices."""
    from xcorr_b200.signal.correlate import timelapse_pairs
    D = 365
    pairs = timelapse_pairs(D)
    assert len(pairs) == D * (D + 1) // 2 + D - 1 and (0, 1) in pairs and (0, 2) not in pairs and (364, 0) in pairs
    mask = np.tri(D, D, k=1, dtype=bool)                     # signal/tri.py:19-55
    assert sorted(pairs) == sorted(zip(*np.nonzero(mask)))
    for world in (1, 2, 8):
        shares = [pairs[r::world] for r in range(world)]
        assert sorted(p for s in shares for p in s) == sorted(pairs)
        assert max(map(len, shares)) - min(map(len, shares)) <= 1
        per = -(-D // world)
        days = [list(range(min(D, r * per), min(D, (r + 1) * per))) for r in range(world)]
        assert sorted(d for s in days for d in s) == list(range(D))
