"""Two-GPU tests of the series-sharded transform (skipped on one-GPU boxes):
the peer-memory push gather and the NCCL all-gather assemble the same feature
matrix as a single-GPU transform of the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import specs

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import fruits_b200 as fruits
    from fruits_b200.parallel import PeerGather, fit_sharded, transform_sharded
    S = 4500
    X = np.random.default_rng(5).standard_normal((world * S, 3, 96))
    Xl = torch.from_numpy(X[rank * S:(rank + 1) * S]).to(dev)
    fruit = specs.build_fruit(fruits, specs.SPECS["C5_sweep"])
    np.random.seed(3)
    fit_sharded(fruit, Xl, world * S)
    nf = fruit.nfeatures()
    out = PeerGather(S, nf) if mode == "peer" else None
    for _ in range(2):                      # the second pass reuses the peer buffers
        res = transform_sharded(lambda x, o: fruit.transform_device(x, out=o), Xl, nf,
                                chunks=3, out=out)
    torch.cuda.synchronize()
    # single-GPU result of the whole batch with the same thresholds
    ref = fruit.transform_device(torch.from_numpy(X).to(dev))
    q.put((rank, bool(torch.equal(res, ref)), res.shape))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_two_gpu_sharded_transform(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, same, shape in results:
        assert shape == (9000, 2225)
        assert same, f"rank {rank}: assembled matrix differs from the single-GPU transform"
