"""Development helper: run ad-hoc ISS descriptions and compare with the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fruits_b200 as fr  # noqa: E402
import specs  # noqa: E402
from cases import make_iss_input  # noqa: E402
from oracle import pipeline as orc  # noqa: E402

desc = json.loads(sys.argv[1])
shape = tuple(json.loads(sys.argv[2]))
kind = sys.argv[3] if len(sys.argv) > 3 else "walk"
X = make_iss_input(shape, kind)
ref = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
res = specs.build_iss(fr, desc).transform(X)
scale = np.max(np.abs(ref), axis=-1, keepdims=True) + 1e-300
print("OK", sys.argv[1][:80], shape, "max rel", np.max(np.abs(res - ref) / scale))
