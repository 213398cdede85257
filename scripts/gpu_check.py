"""Development helper: run the CUDA path against the golden vectors and print
a mismatch summary for every case (does not stop at the first failure)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import fruits_b200 as fr  # noqa: E402
import specs  # noqa: E402
from cases import (ISS_CASES, PREP_CASES, SIEVE_CASES, make_iss_input,  # noqa: E402
                   make_prep_input, make_sieve_input)

GOLD = os.path.join(ROOT, "tests", "golden")


def cmp(name, got, ref, exact=True, rtol=1e-9):
    got, ref = np.asarray(got), np.asarray(ref)
    if got.shape != ref.shape:
        print(f"  FAIL {name}: shape {got.shape} vs {ref.shape}")
        return False
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    if same.all():
        print(f"  ok   {name}: bit-exact")
        return True
    fin = np.where(np.isfinite(ref), ref, 0.0)
    scale = np.maximum(np.abs(fin), np.max(np.abs(fin), axis=-1, keepdims=True))
    with np.errstate(invalid="ignore"):
        err = np.where(same, 0.0, np.abs(got - ref))
    rel = err / (scale + 1e-300)
    nbad = int((rel > rtol).sum())
    msg = f"{(~same).sum()} of {same.size} differ, max rel {rel.max():.3e}, {nbad} beyond {rtol}"
    if exact or nbad:
        idx = np.argwhere(~same)[:3].tolist()
        print(f"  FAIL {name}: {msg}; first {idx}")
        return False
    print(f"  ok   {name}: {msg}")
    return True


def main():
    print(torch.cuda.get_device_name(0))
    ok = True
    g = np.load(os.path.join(GOLD, "iss.npz"))
    print("[iss]")
    for name, (desc, shape, kind) in ISS_CASES.items():
        try:
            X = make_iss_input(shape, kind)
            res = specs.build_iss(fr, desc).transform(X)
            exact = desc.get("weighting") is None or desc.get("semiring") == "arctic"
            ok &= cmp(name, res, g[name], exact=exact, rtol=1e-9)
        except Exception:
            ok = False
            print(f"  EXC  {name}")
            traceback.print_exc()
    print("[preps]")
    g = np.load(os.path.join(GOLD, "preps.npz"))
    X = make_prep_input()
    for name, desc in PREP_CASES.items():
        try:
            if desc[0] == "NRM":
                continue
            p = specs._prep(fr, desc)
            p.fit(X)
            ok &= cmp(name, p.transform(X), g[name])
        except Exception:
            ok = False
            print(f"  EXC  {name}")
            traceback.print_exc()
    print("[sieves]")
    g = np.load(os.path.join(GOLD, "sieves.npz"))
    raw, Y = make_sieve_input()
    for name, desc in SIEVE_CASES.items():
        try:
            sv = specs._sieve(fr, desc)
            sv._cache = fr.cache.SharedSeedCache(raw)
            np.random.seed(3)
            sv.fit(Y)
            res = sv.transform(Y)
            thr = sv._q if desc[0] == "PPV" else sv._quantiles
            ok &= cmp(name + "_thr", np.array(thr, dtype=np.float64), g[name + "_thr"])
            ok &= cmp(name, res, g[name], exact=desc[0] not in ("MPI", "XPI"), rtol=1e-12)
        except Exception:
            ok = False
            print(f"  EXC  {name}")
            traceback.print_exc()
    print("[pipelines]")
    for name in ["C1_readme", "C5_sweep", "C2_reduced", "C4_twi", "C3_general"]:
        try:
            g = np.load(os.path.join(GOLD, f"pipeline_{name}.npz"))
            spec = specs.SPECS[name]
            X = specs.make_input(name, int(g["n"]))
            fruit = specs.build_fruit(fr, spec)
            np.random.seed(0)
            t0 = time.time()
            fruit.fit(X)
            t1 = time.time()
            res = fruit.transform(X)
            t2 = time.time()
            print(f"  {name}: fit {t1-t0:.3f}s transform {t2-t1:.3f}s")
            exact = name in ("C1_readme", "C5_sweep")
            thr = []
            for slc in fruit:
                for sieves in slc._sieves_extended:
                    for sv in sieves:
                        q = getattr(sv, "_quantiles", None)
                        if q is None:
                            q = getattr(sv, "_q", [])
                        thr.append(np.asarray(q, dtype=np.float64).ravel())
            thr = np.concatenate(thr) if thr else np.zeros(0)
            ok &= cmp(name + " thresholds", thr, g["thresholds"], exact=exact, rtol=1e-9)
            ok &= cmp(name + " features", res, g["features"], exact=exact, rtol=1e-9)
        except Exception:
            ok = False
            print(f"  EXC  {name}")
            traceback.print_exc()
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
