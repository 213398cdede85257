"""Host time per ``transform_device`` call against the device time of its kernels
(development aid): small configurations (C1, C2) are bound by the Python above the C ABI,
not by the kernels.

    python scripts/host_overhead.py C2_full [calls]
"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import fruits_b200 as fruits  # noqa: E402
import specs  # noqa: E402

if __name__ == "__main__":
    name = sys.argv[1]
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    X = torch.from_numpy(specs.make_input(name)).cuda()
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    np.random.seed(0)
    fruit.fit(X)
    out = fruit.transform_device(X)
    for _ in range(5):
        fruit.transform_device(X, out=out)
    torch.cuda.synchronize()
    # host time of a call that only enqueues (the stream is drained first)
    t0 = time.perf_counter()
    for _ in range(calls):
        fruit.transform_device(X, out=out)
    t_enqueue = (time.perf_counter() - t0) / calls
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / calls
    # device time of the same work with the host out of the way: one CUDA graph replay
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    t_graph = None
    try:
        with torch.cuda.stream(s):
            fruit.transform_device(X, out=out)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                fruit.transform_device(X, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(calls):
            g.replay()
        ev[1].record()
        torch.cuda.synchronize()
        t_graph = ev[0].elapsed_time(ev[1]) / calls
    except Exception as exc:      # noqa: BLE001
        print(f"graph capture failed: {type(exc).__name__}: {exc}")
    print(f"{name}: {X.shape[0]} series, {out.shape[1]} features: host enqueue {t_enqueue * 1e3:.3f} ms per call, "
          f"wall {t_total * 1e3:.3f} ms per call, device (graph replay) "
          f"{'n/a' if t_graph is None else f'{t_graph:.3f} ms'}", flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(calls):
        fruit.transform_device(X, out=out)
    pr.disable()
    torch.cuda.synchronize()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(22)
    print("\n".join(l for l in buf.getvalue().splitlines() if l.strip())[:6000])
