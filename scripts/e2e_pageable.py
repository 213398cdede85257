"""Fruit.transform end to end from ordinary (pageable) numpy arrays vs pinned
host buffers (development aid).

    python scripts/e2e_pageable.py [n_series]
"""
import os
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    X = np.random.default_rng(0).standard_normal((n, 3, 1024))
    fruit = specs.build_fruit(fruits, specs.SPECS["C5_sweep"])
    np.random.seed(0)
    fruit.fit(X[:64])
    for label, make in (("pageable in, fresh out", lambda: (X, None)),
                        ("pageable in, pageable out", lambda: (X, np.empty((n, 2225)))),
                        ("pinned in, pinned out", None)):
        if make is None:
            hx = torch.empty((n, 3, 1024), dtype=torch.float64, pin_memory=True)
            hx.copy_(torch.from_numpy(X))
            hf = torch.empty((n, 2225), dtype=torch.float64, pin_memory=True)
            xin, out = hx.numpy(), hf.numpy()
        else:
            xin, out = make()
        fruit.transform(xin, out=out) if out is not None else fruit.transform(xin)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            res = fruit.transform(xin, out=out) if out is not None else fruit.transform(xin)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"{label:28s}: {dt * 1e3:8.1f} ms  {n / dt / 1e6:6.3f} M series/s", flush=True)
