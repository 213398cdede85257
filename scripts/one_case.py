"""Development helper: run a single ISS golden case (for compute-sanitizer)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fruits_b200 as fr  # noqa: E402
import specs  # noqa: E402
from cases import ISS_CASES, make_iss_input  # noqa: E402

name = sys.argv[1]
desc, shape, kind = ISS_CASES[name]
X = make_iss_input(shape, kind)
res = specs.build_iss(fr, desc).transform(X)
g = np.load(os.path.join(ROOT, "tests", "golden", "iss.npz"))
print(name, "max abs diff", np.nanmax(np.abs(res - g[name])))
