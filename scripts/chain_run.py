"""One slice of a configuration through Fruit.transform_device a few times (for ncu).

    python scripts/chain_run.py C3_general 1 [n_series]
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
    fruit = specs.build_fruit(fruits, one)
    np.random.seed(0)
    fruit.fit(X)
    out = fruit.transform_device(X)
    for _ in range(3):
        fruit.transform_device(X, out=out)
    torch.cuda.synchronize()
    print(fruit.get_slice(0)._last_launch[0], float(out.sum()))
