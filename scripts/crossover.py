"""Generic kernel vs plan-specialised kernel as a function of the batch size
(development aid for _jit.MIN_SERIES).

    python scripts/crossover.py [config]
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


def timed(fruit, X, out):
    fruit.transform_device(X, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        fruit.transform_device(X, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / 5


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "C5_sweep"
    shape = {"C5_sweep": (3, 1024), "C4_twi": (3, 2048), "C1_readme": (3, 100)}[name]
    for n in (256, 512, 1024, 2048, 3072, 4096, 8192):
        X = torch.randn((n,) + shape, dtype=torch.float64, device="cuda")
        res = {}
        for mode in ("0", "force"):
            os.environ["FRUITS_B200_JIT"] = mode
            fruit = specs.build_fruit(fruits, specs.SPECS[name])
            np.random.seed(0)
            fruit.fit(X)
            out = torch.empty((n, fruit.nfeatures()), dtype=torch.float64, device="cuda")
            res[mode] = timed(fruit, X, out)
        print(f"{name} n={n:6d}: generic {res['0']:8.3f} ms   generated {res['force']:8.3f} ms", flush=True)
