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


def _run_guarded(fn, rank, *args):
    """Report an exception of one rank through the queue instead of leaving the
    other rank waiting in a collective until the test times out."""
    import traceback
    q = args[-1]
    try:
        fn(rank, *args)
    except BaseException:          # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))
        q.close()
        q.join_thread()                # the feeder thread must flush before the process dies
        os._exit(1)


def _collect(procs, q, n):
    results = []
    for _ in range(n):
        r = q.get(timeout=300)
        if len(r) > 1 and isinstance(r[1], str) and r[1] == "error":
            for p in procs:
                p.kill()
            pytest.fail(f"rank {r[0]} raised:\n{r[2]}")
        results.append(r)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return results


def _worker_(rank, world, port, mode, q):
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
    out = None
    if mode in ("peer", "multicast"):
        out = PeerGather(S, nf, multicast=(mode == "multicast"))
        if mode == "multicast" and not out.fused:
            q.put((rank, None, None))          # no NVLS on this box
            dist.barrier()
            dist.destroy_process_group()
            return
    for _ in range(2):                      # the second pass reuses the peer buffers
        res = transform_sharded(fruit, Xl, nf, chunks=3, out=out)
    torch.cuda.synchronize()
    # single-GPU result of the whole batch with the same thresholds
    ref = fruit.transform_device(torch.from_numpy(X).to(dev))
    same = bool(torch.equal(res, ref))
    if mode == "multicast":
        # a fruit whose slices take different routes: generated kernel (multimem.st in
        # its epilogue), generic kernel and composed route (fb_multimem_copy)
        mixed = specs.build_fruit(fruits, {"slices": specs.SPECS["C1_readme"]["slices"]
                                           + [specs.SPECS["R_mixed"]["slices"][1]]})
        Xm = np.random.default_rng(6).random((world * 4200, 3, 60)) + 0.1
        Xml = torch.from_numpy(Xm[rank * 4200:(rank + 1) * 4200]).to(dev)
        np.random.seed(4)
        fit_sharded(mixed, Xml, world * 4200)
        pg = PeerGather(4200, mixed.nfeatures(), multicast=True)
        resm = transform_sharded(mixed, Xml, mixed.nfeatures(), out=pg)
        torch.cuda.synchronize()
        refm = mixed.transform_device(torch.from_numpy(Xm).to(dev))
        routes = {getattr(slc, "_last_launch", ("?",))[0] for slc in mixed}
        same = same and bool(torch.equal(resm, refm)) and len(routes) >= 2
    q.put((rank, same, res.shape))
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, mode, q):
    _run_guarded(_worker_, rank, world, port, mode, q)


@pytest.mark.parametrize("mode", ["multicast", "peer", "nccl"])
def test_two_gpu_sharded_transform(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = _collect(procs, q, len(procs))
    if any(same is None for _, same, _ in results):
        pytest.skip("no NVSwitch multicast support on this box")
    for rank, same, shape in results:
        assert shape == (9000, 2225)
        assert same, f"rank {rank}: assembled matrix differs from the single-GPU transform"


def _fit_worker_(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import fruits_b200 as fruits
    from fruits_b200.parallel import fit_sharded, shard_rows
    from helpers import fitted_thresholds
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": {"of_weight": [3, 2]}, "mode": "extended"}],
                        "sieves": [["NPI", {"q": [0.3, 1.0]}], ["PPV", {"sample_size": 0.5}],
                                   ["MPI", {"q": [0.2, 0.7], "inc": 2}], ["END", {}]],
                        "fit_sample_size": 0.6},
                       {"iss": [{"words": ["[1][2]", "[2][1][1]", "[1]"], "mode": "extended",
                                 "semiring": "arctic"}],
                        "sieves": [["MAX", {"q": [-1.0, 0.5]}], ["NPI", {"q": [0.5, 1.0]}]],
                        "fit_sample_size": 1.0}]}
    # the reference's cache quirk (fruits/cache.py:97-112): L1 weighting on the raw
    # input + fitted sieves + a subsample; an arctic chain whose increments are mostly
    # 0 (the eight-pass fallback of the select); a one-series sample (one rank empty)
    spec["slices"] += [
        {"preps": [["INC", {}]],
         "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended", "weighting": ["L1", {}]}],
         "sieves": [["NPI", {"q": [0.4, 1.0]}], ["MPI", {"q": [0.1, 0.9], "inc": 0}], ["END", {}]],
         "fit_sample_size": 0.5},
        {"iss": [{"words": {"alternate_sign": [6 * "[1]", 3 * "[1][2]"]}, "mode": "extended",
                  "semiring": "arctic"}],
         "sieves": [["NPI", {"q": [0.5, 1.0], "inc": i}] for i in range(3)] + [["END", {}]],
         "fit_sample_size": 1.0},
        {"iss": [{"words": ["[1]", "[12]"], "mode": "extended"}],
         "sieves": [["NPI", {"q": [0.5, 1.0]}], ["PPV", {}]], "fit_sample_size": 1},
        {"iss": [{"words": ["[1][1]"]}], "sieves": [["NPI", {}], ["END", {}]],
         "fit_sample_size": 1.0}]
    n = 301
    X = np.random.default_rng(8).standard_normal((n, 2, 120)).cumsum(axis=2)
    lo, hi = shard_rows(n, world, rank)
    Xl = torch.from_numpy(X[lo:hi]).to(dev)
    single = specs.build_fruit(fruits, spec)
    np.random.seed(40)
    single.fit(torch.from_numpy(X).to(dev))
    after_single = np.random.random()
    ok_thr = ok_feats = ok_rng = True
    n_thr = 0
    for mode in ("rows", "nodes"):
        np.random.seed(40 + rank)                # different states: rank 0's is broadcast
        if rank == 0:
            np.random.seed(40)
        sp = spec if mode == "rows" else {"slices": spec["slices"][:2] + spec["slices"][3:]}
        fruit = specs.build_fruit(fruits, sp)
        fit_sharded(fruit, Xl, n, shard="auto" if mode == "rows" else "nodes")
        after = np.random.random()
        if mode == "rows":
            ref, ref_after = single, after_single
        else:                                    # (node mode refuses the L1 quirk slice)
            ref = specs.build_fruit(fruits, sp)
            np.random.seed(40)
            ref.fit(torch.from_numpy(X).to(dev))
            ref_after = np.random.random()
        ok_thr &= bool(np.array_equal(fitted_thresholds(fruit), fitted_thresholds(ref)))
        ok_feats &= bool(torch.equal(fruit.transform_device(Xl), ref.transform_device(Xl)))
        ok_rng &= after == ref_after
        n_thr += len(fitted_thresholds(fruit))
    q.put((rank, ok_thr, ok_feats, ok_rng, n_thr))
    dist.barrier()
    dist.destroy_process_group()


def _fit_worker(rank, world, port, q):
    _run_guarded(_fit_worker_, rank, world, port, q)


def test_two_gpu_fit_sharded():
    """fit_sharded, row mode (the sample stays sharded, the selections sum their
    histograms over the ranks) and node mode (the sample is gathered, each rank
    fits its share of the iterated sums): thresholds, features and the
    consumption of the global numpy RNG equal a single-GPU fit of the whole batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = _collect(procs, q, len(procs))
    for rank, same, feats, rng_ok, n_thr in results:
        assert n_thr > 0
        assert same, f"rank {rank}: thresholds differ from the single-GPU fit"
        assert feats, f"rank {rank}: features differ"
        assert rng_ok, f"rank {rank}: numpy RNG consumed differently"
