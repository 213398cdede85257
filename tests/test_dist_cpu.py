"""CPU tests (gloo, world_size 2) of the series-sharded multi-GPU plumbing:
row sharding, the fit-sample gather and the chunked feature all-gather.  The
CUDA compute is replaced by a stand-in that marks every row, so the test
checks exactly what crosses ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fruits_b200.parallel import (gather_fit_sample, shard_rows, sync_numpy_rng,
                                  transform_sharded)


def test_shard_rows_cover_everything():
    for n in (0, 1, 7, 8, 1000, 4 * 1024 * 1024):
        for world in (1, 2, 4, 8):
            blocks = [shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = torch.arange(n_total * 2 * 5, dtype=torch.float64).reshape(n_total, 2, 5)
        # bit patterns a floating point sum over zero-filled rows would lose:
        # -0.0 (1 / -0.0 = -inf in a letter with a negative exponent) and a NaN payload
        X[:, 1, 0] = -0.0
        X[:, 1, 1] = torch.tensor([0x7ff8000000000abc], dtype=torch.int64).view(torch.float64)
        lo, hi = shard_rows(n_total, world, rank)
        Xl = X[lo:hi].clone()
        # fit sample: all ranks end up with the rows rank 0's RNG state selects
        np.random.seed(100 + rank)         # different states before the sync
        sync_numpy_rng()
        one = gather_fit_sample(Xl, n_total, 1)
        frac = gather_fit_sample(Xl, n_total, 0.5)
        after = np.random.random()
        np.random.seed(100)
        i1 = np.random.randint(0, n_total)
        i2 = np.random.choice(n_total, size=n_total // 2, replace=False)
        expect_after = np.random.random()
        bits = lambda a: a.contiguous().view(torch.int64)        # noqa: E731
        ok = (torch.equal(bits(one), bits(X[i1:i1 + 1])) and torch.equal(bits(frac), bits(X[i2]))
              and after == expect_after)
        ok = ok and bool(torch.signbit(frac[:, 1, 0]).all())

        # chunked all-gather of features: S rows per rank, rank-major result
        S = hi - lo
        assert S * world == n_total

        def compute(x, o):
            o.copy_(x[:, 0, :3] * 10 + 1)

        feats = transform_sharded(compute, Xl, 3, chunks=3)
        ok = ok and torch.equal(feats, X[:, 0, :3] * 10 + 1)
        feats1 = transform_sharded(compute, Xl, 3, chunks=1)
        ok = ok and torch.equal(feats1, feats)
        # rows of a row-sharded tensor, requested by global row number (any order, repeats)
        from fruits_b200.parallel import exchange_rows
        need = [n_total - 1, 0, 3, 3, lo] if rank == 0 else [1, hi - 1]
        got = exchange_rows(Xl, n_total, need)
        ok = ok and torch.equal(bits(got), bits(X[need]))
        # uneven shards are refused on every rank instead of hanging or mis-placing rows
        try:
            transform_sharded(compute, Xl[:S - rank], 3, chunks=1)
            ok = False
        except ValueError as exc:
            ok = ok and "equally sized" in str(exc)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_fit_sample():
    world, n_total = 2, 14
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True), (1, True)]


def test_shard_by_cost_balances_contiguous_blocks():
    from fruits_b200.parallel import shard_by_cost
    costs = [1] * 10 + [27] * 10 + [81] * 10
    for world in (1, 2, 3, 4, 8):
        blocks = [shard_by_cost(30, costs, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 30
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        loads = [sum(costs[lo:hi]) for lo, hi in blocks]
        assert max(loads) <= sum(costs) / world + max(costs)
    assert shard_by_cost(7, None, 3, 1) == shard_rows(7, 3, 1)
    # more ranks than items: empty blocks are allowed, coverage is not lost
    blocks = [shard_by_cost(2, [5, 5], 4, r) for r in range(4)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 2
    assert sum(hi - lo for lo, hi in blocks) == 2
