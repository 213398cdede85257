"""A few launches of two preparateur kernels (fb_time_mask keeping everything,
fb_lead_lag) on 65,536 x 3 x 1,024 series, for an ncu capture:

    ncu --set full --clock-control none -k regex:'time_mask|lead_lag' -c 2 \
        -o gpurun_out/r2_prep python scripts/prep_kernel_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fruits_b200 as fruits  # noqa: E402

P = fruits.preparation
X = torch.randn((65536, 3, 1024), dtype=torch.float64, device="cuda").cumsum(dim=2)
np.random.seed(0)
for prep in (P.PDD(), P.LAG()):
    prep._fit_device(X)
    for _ in range(2):
        out = prep._transform_device(X)
    torch.cuda.synchronize()
    print(prep, tuple(out.shape))
