"""Part-function calling conventions of the generated kernel (development aid):
    python scripts/abi_sweep.py C4_twi 0 [n]
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
    name, si = sys.argv[1], int(sys.argv[2])
    n = int(sys.argv[3]) if len(sys.argv) > 3 else None
    X = torch.from_numpy(specs.make_input(name, n)).cuda()
    one = {"slices": [dict(specs.SPECS[name]["slices"][si], fit_sample_size=1)]}
    for abi in (0, 1, 2, 3):
        # (an explicit option string switches the layout heuristic of _jit.generate off:
        # deep / sieve-heavy tries get their 255-register layout spelled out)
        wide = name in ("C4_twi", "C3_general", "C2_reduced")
        os.environ["FRUITS_B200_JIT_OPTS"] = f"abi={abi}" + (
            ",budget=150,ppc=1,gpc=8,minb=1,unroll=1" if wide else "")
        fruit = specs.build_fruit(fruits, one)
        np.random.seed(0)
        fruit.fit(X)
        out = fruit.transform_device(X)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            fruit.transform_device(X, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"{name} slice {si} abi={abi}: {ev[0].elapsed_time(ev[1]) / 5:8.2f} ms "
              f"route={fruit.get_slice(0)._last_launch[0]}", flush=True)
