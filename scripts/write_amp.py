"""DRAM traffic of the C5 feature kernel with the feature rows at their natural
stride (2225 doubles = 17,800 B, not a multiple of the 32-byte sector) and with
rows padded to whole sectors (2228 doubles) -- for ncu:

    ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,lts__t_sectors_op_write.sum \
        -k regex:fb_jit_slice python scripts/write_amp.py
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    fruit = specs.build_fruit(fruits, specs.SPECS["C5_sweep"])
    np.random.seed(0)
    fruit.fit(specs.make_input("C5_sweep", 64))
    X = torch.randn((n, 3, 1024), dtype=torch.float64, device="cuda")
    for ld in (2225, 2228):
        buf = torch.empty((n, ld), dtype=torch.float64, device="cuda")
        out = buf[:, :2225]
        fruit.transform_device(X, out=out)      # launch 1: natural stride, launch 2: padded rows
        torch.cuda.synchronize()
        print(ld, float(out[0, 0]))
