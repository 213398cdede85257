"""Device time of Fruit.fit by kernel (torch.profiler / CUPTI; development aid).

    python scripts/fit_kernels.py C3_general [n_series]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import fruits_b200 as fruits  # noqa: E402
import specs  # noqa: E402

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "C3_general"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else None
    X = torch.from_numpy(specs.make_input(name, n)).cuda()
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    np.random.seed(0)
    fruit.fit(X)                      # warm: plans, allocator
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    np.random.seed(0)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fruit.fit(X)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25,
                                    max_name_column_width=70))
