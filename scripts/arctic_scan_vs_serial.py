"""ISS.transform of the arctic chains (C3 slice 1 words) through the block scan over T
(fb_arctic_word) and through the lane-per-node kernel, for a few batch shapes:
    python scripts/arctic_scan_vs_serial.py
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
    desc = specs.SPECS["C3_general"]["slices"][1]["iss"][0]
    for n, t in ((1, 1024), (1, 65536), (4, 16384), (16, 1024), (256, 1024), (2048, 1024)):
        X = torch.from_numpy(np.random.default_rng(1).standard_normal((n, 2, t)).cumsum(axis=2)).cuda()
        res = {}
        for mode in ("0", "1"):
            os.environ["FRUITS_B200_ARCTIC_SCAN"] = mode
            iss = specs.build_iss(fruits, desc)
            out = iss.transform(X)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(3):
                out = iss.transform(X)
            ev[1].record()
            torch.cuda.synchronize()
            res[mode] = (ev[0].elapsed_time(ev[1]) / 3, out)
        same = bool(torch.equal(res["0"][1], res["1"][1]))
        print(f"n={n:5d} T={t:6d}: lane-per-node {res['0'][0]:8.3f} ms, block scan over T "
              f"{res['1'][0]:8.3f} ms, identical={same}", flush=True)
