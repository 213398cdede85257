"""Fruit.fit on N GPUs (parallel.fit_sharded: row mode -- the sample stays
sharded, histograms are all-reduced -- and node mode -- the sample is gathered,
the iterated sums are split) against the single-GPU fit of the same batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        scripts/fit_sharded_time.py C3_full [rows|nodes]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fruits_b200 as fruits  # noqa: E402
import specs  # noqa: E402
from fruits_b200.parallel import fit_sharded, shard_rows  # noqa: E402
from helpers import fitted_thresholds  # noqa: E402

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "C3_general"
    mode = sys.argv[2] if len(sys.argv) > 2 else "auto"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    X = specs.make_input(name)
    n = X.shape[0]
    lo, hi = shard_rows(n, world, rank)
    Xl = torch.from_numpy(X[lo:hi]).to(dev)
    times = []
    for rep in range(2):
        fruit = specs.build_fruit(fruits, specs.SPECS[name])
        np.random.seed(0)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch.cuda.reset_peak_memory_stats()
        fit_sharded(fruit, Xl, n, shard=mode)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t0)
    thr = fitted_thresholds(fruit)
    if rank == 0:
        single = specs.build_fruit(fruits, specs.SPECS[name])
        np.random.seed(0)
        Xd = torch.from_numpy(X).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        single.fit(Xd)
        torch.cuda.synchronize()
        t1 = time.perf_counter() - t0
        same = bool(np.array_equal(thr, fitted_thresholds(single), equal_nan=True))
        print(f"{name} [{mode}]: fit on {world} GPUs {times[-1]:.3f} s (first call {times[0]:.3f} s), "
              f"peak device memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB, "
              f"single GPU {t1:.3f} s, {len(thr)} thresholds identical={same}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
