"""Cost model of the plan-specialised kernel: the C5 word set with different
sieve sets (development aid).  Prints ms and issue cycles per (warp, node, step).

    python scripts/jit_model.py --compile      (no GPU: fills the cubin cache)
    python scripts/jit_model.py --run N
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import copy  # noqa: E402

import specs  # noqa: E402

SETS = {
    "END": [["END", {}]],
    "NPI": [["NPI", {"q": [0.5, 1.0]}]],
    "PPV": [["PPV", {}]],
    "NPI+PPV": [["NPI", {"q": [0.5, 1.0]}], ["PPV", {}]],
    "MAX": [["MAX", {}]],
    "MAX+MIN": [["MAX", {}], ["MIN", {}]],
    "NPI0": [["NPI", {"q": [0.5, 1.0], "inc": 0}]],
    "all": specs.SPECS["C5_sweep"]["slices"][0]["sieves"],
}

for name, sv in SETS.items():
    spec = copy.deepcopy(specs.SPECS["C5_sweep"])
    spec["slices"][0]["sieves"] = sv
    specs.SPECS["M_" + name] = spec


def main() -> None:
    mode = sys.argv[1]
    names = [a for a in sys.argv[2:] if a in SETS] or list(SETS)
    if mode == "--compile":
        import jit_warm
        for name in names:
            jit_warm.SHAPES["M_" + name] = 3
            t0 = time.time()
            jit_warm.warm("M_" + name)
            print(f"compiled {name} in {time.time() - t0:.1f} s", flush=True)
        return
    n = int(sys.argv[2])
    import numpy as np
    import torch
    import fruits_b200 as fruits
    X = torch.randn((n, 3, 1024), dtype=torch.float64, device="cuda",
                    generator=torch.Generator("cuda").manual_seed(1234))
    for name in names:
        fruit = specs.build_fruit(fruits, specs.SPECS["M_" + name])
        np.random.seed(0)
        fruit.fit(X[:64])
        out = torch.empty((n, fruit.nfeatures()), dtype=torch.float64, device="cuda")
        fruit.transform_device(X, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        reps = 3
        for _ in range(reps):
            fruit.transform_device(X, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        # warp-level node-steps per SM sub-partition and cycle (1.965 GHz, 148 x 4)
        cyc = ms * 1e-3 * 1.965e9 * 592 / (n / 32 * 445 * 1024)
        print(f"[{name}] {ms:8.2f} ms  {n / ms * 1e3 / 1e6:6.3f} M series/s  "
              f"{cyc:5.2f} cycles per warp node-step", flush=True)


if __name__ == "__main__":
    main()
