"""Two ranks on two GPUs (or skipped): windows sharded over ranks, per-rank
correlation through the C ABI, NCCL gather of the CCFs / all-reduce of stacks.
Run by the gpu tier only when >= 2 devices are visible."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import xcorr_b200 as xb
    from xcorr_b200 import parallel
    from oracle import cc_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, R, W = 30_000, 3, 5
        pairs = xb.all_pairs(R)
        rec = np.stack([orc.synthetic_receivers(R, n, seed=900 + w, delay=23) for w in range(W)], axis=1)
        mine = parallel.shard_windows(W, rank, world)
        x = torch.from_numpy(rec[:, mine].astype(np.float32)).cuda()
        plan = xb.CorrelationPlan(n, R, pairs, n_windows=len(mine), dtype=torch.float32)
        local = plan(x)
        full = parallel.gather_ccf(local, W, dst=0)
        stack = parallel.reduce_stack(local.sum(dim=1))
        ok = True
        # config 5 sharding: day transforms split over the ranks, ONE all-gather of the half spectra, the day
        # pairs tiled over the ranks; must equal the unsharded plan bit for bit
        from xcorr_b200.signal import Correlate2dPlan, timelapse_pairs
        F, T, D = 64, 16, 7
        g = torch.Generator(device="cuda")
        g.manual_seed(5)
        days = torch.randn((D, F, T), generator=g, device="cuda", dtype=torch.float32) ** 2
        allp = timelapse_pairs(D)
        sharded = Correlate2dPlan(F, T, D, allp[rank::world], dtype=torch.float32, day_shard=(rank, world))
        v, dF, dT = sharded.correlate_argmax(days[sharded.day_slice])
        single = Correlate2dPlan(F, T, D, allp[rank::world], dtype=torch.float32)
        v1, dF1, dT1 = single.correlate_argmax(days)
        ok &= bool(torch.equal(v, v1) and torch.equal(dF, dF1) and torch.equal(dT, dT1))
        if rank == 0:
            got = full.cpu().numpy()
            for qq, (i, j) in enumerate(pairs):
                for w in range(W):
                    ref = orc.cc(rec[i, w], rec[j, w], unbiased=False)
                    ok &= bool(np.max(np.abs(got[qq, w] - ref)) <= 1e-5 * np.max(np.abs(ref)))
            ok &= bool(np.allclose(stack.cpu().numpy(), got.sum(axis=1), atol=1e-5))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_window_sharding_nccl():
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
