"""Per-slice transform time of a BASELINE.json configuration on one GPU, with
the route each slice takes (development aid).

    python scripts/slice_times.py C4_twi [n_series]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import fruits_b200 as fruits  # noqa: E402
import specs  # noqa: E402

if __name__ == "__main__":
    name = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else None
    X = torch.from_numpy(specs.make_input(name, n)).cuda()
    spec = specs.SPECS[name]
    for si, slc in enumerate(spec["slices"]):
        one = {"slices": [dict(slc, fit_sample_size=1)]}
        fruit = specs.build_fruit(fruits, one)
        np.random.seed(0)
        fruit.fit(X)
        out = fruit.transform_device(X)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(3):
            fruit.transform_device(X, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        route = getattr(fruit.get_slice(0), "_last_launch", ("?",))[0]
        print(f"{name} slice {si}: {X.shape[0]} series -> {out.shape[1]} features  {ms:8.2f} ms  "
              f"route={route}", flush=True)
