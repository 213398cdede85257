"""NVLink traffic of the fused feature gather (development aid / evidence for profiles/).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        scripts/nvlink_traffic.py [--series-per-gpu S] [--steps K] [--no-multicast]

Every rank transforms its C5 shard with the features stored through the NVSwitch
multicast mapping (as bench.py does at N > 1); rank 0 reads the NVLink data
counters of all GPUs (``nvidia-smi nvlink -gt d``) before and after the timed
steps and prints the bytes per GPU and step next to the algorithmic figures:
sent = S x F x 8 (one multicast store stream), received = (N - 1) x S x F x 8.
"""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def counters(n_gpus):
    """{gpu: (tx_kib, rx_kib)} summed over the links."""
    out = {}
    for i in range(n_gpus):
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(i)],
                             capture_output=True, text=True).stdout
        tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt))
        rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt))
        out[i] = (tx, rx)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--series-per-gpu", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-multicast", action="store_true")
    args = ap.parse_args()
    import fruits_b200 as fruits
    import specs
    from fruits_b200.parallel import PeerGather, transform_sharded
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S, F = args.series_per_gpu, 2225
    fruit = specs.build_fruit(fruits, specs.SPECS["C5_sweep"])
    np.random.seed(0)
    fruit.fit(specs.make_input("C5_sweep", 64))
    gen = torch.Generator(device=dev)
    gen.manual_seed(77 + rank)
    X = torch.randn((S, 3, 1024), dtype=torch.float64, device=dev, generator=gen)
    peer = PeerGather(S, F, multicast=not args.no_multicast)
    for _ in range(2):
        transform_sharded(fruit, X, F, out=peer)
    torch.cuda.synchronize()
    dist.barrier()
    before = counters(world) if rank == 0 else None
    dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        transform_sharded(fruit, X, F, out=peer)
    ev[1].record()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        after = counters(world)
        alg_tx = S * F * 8
        alg_rx = (world - 1) * S * F * 8
        per_gpu = {}
        for i in range(world):
            tx = (after[i][0] - before[i][0]) * 1024 / args.steps
            rx = (after[i][1] - before[i][1]) * 1024 / args.steps
            per_gpu[i] = {"tx_bytes_per_step": tx, "rx_bytes_per_step": rx,
                          "tx_over_algorithmic": tx / alg_tx, "rx_over_algorithmic": rx / max(alg_rx, 1)}
        print(json.dumps({"n_gpus": world, "series_per_gpu": S, "steps": args.steps,
                          "fused_multicast": bool(peer.fused), "ms_per_step": ev[0].elapsed_time(ev[1]) / args.steps,
                          "algorithmic_tx_bytes_per_gpu_and_step": alg_tx,
                          "algorithmic_rx_bytes_per_gpu_and_step": alg_rx, "per_gpu": per_gpu}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
