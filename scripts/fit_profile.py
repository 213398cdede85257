"""Where does Fruit.fit spend its time?  cProfile of the host side plus the
device time of the same call (development aid).

    python scripts/fit_profile.py C3_general [n_series]
"""
import cProfile
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
    name = sys.argv[1] if len(sys.argv) > 1 else "C3_general"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else None
    X = torch.from_numpy(specs.make_input(name, n)).cuda()
    for rep in range(2):
        fruit = specs.build_fruit(fruits, specs.SPECS[name])
        np.random.seed(0)
        torch.cuda.synchronize()
        prof = cProfile.Profile()
        t0 = time.perf_counter()
        prof.enable()
        fruit.fit(X)
        prof.disable()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print(f"{name} fit #{rep}: host returned after {t_host:.3f} s, device idle after "
              f"{t_all:.3f} s", flush=True)
        if rep == 1:
            pstats.Stats(prof).sort_stats("cumulative").print_stats(35)
