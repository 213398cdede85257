"""Pre-compile the plan-specialised kernels of the benchmark / test pipelines.

NVRTC needs no GPU, so this runs in the build container; the cubins land in
``fruits_b200/lib/jit/`` (git-ignored, shipped to the GPU box with the tree)
and the first ``transform`` on the box skips the compile.

    python scripts/jit_warm.py [config ...]     # default: every config of tests/specs.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fruits_b200 as fruits  # noqa: E402
from fruits_b200 import _jit, _jit_chain  # noqa: E402
from fruits_b200.iss.weighting import Indices, Plateaus  # noqa: E402
import specs  # noqa: E402

SHAPES = {"C1_readme": 3, "C2_reduced": 1, "C2_cos": 1, "C3_general": 6, "C3_cos": 6,
          "C4_twi": 3, "C5_sweep": 3}


def warm(name: str) -> None:
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    for si, slc in enumerate(fruit._slices):
        iss = slc._iss[0]
        feats, bhi, bmm = slc._fused_sieves()
        dims = slc._fused_dims(SHAPES[name])
        trie, n_shared = iss._jit_trie(len(dims))
        used = trie.used_dims()
        if any(dims[u][2] for u in used if u < len(dims)):
            dims = [(u, 0, 0) for u in range(len(dims))]
        jdims = [(dims[u][0], dims[u][1]) if u < len(dims) else ("row", u - len(dims))
                 for u in used]
        shared = iss.weighting is None or isinstance(iss.weighting, (Indices, Plateaus))
        t0 = time.time()
        sieves = _jit.SieveSet.make(feats, bhi, bmm)
        if (not n_shared and _jit_chain.suitable(trie, iss.semiring._code, iss._weight_mode())
                and _jit_chain.chain_like(trie)):
            genc = _jit_chain.generate(trie, iss.semiring._code, iss._weight_mode(), sieves, jdims)
            path = os.path.join(_jit.CACHE_DIR, genc.digest() + ".cubin")
            if not os.path.exists(path):
                os.makedirs(_jit.CACHE_DIR, exist_ok=True)
                with open(path, "wb") as f:
                    f.write(_jit._nvrtc(genc.source, "fb_jit_chain.cu", False, genc.max_regs))
            print(f"{name} slice {si}: {len(trie.nodes)} nodes, chain kernel, "
                  f"{len(genc.em.p.blocks)} blocks x {genc.em.p.rows} rows, "
                  f"{genc.em.spc} series/CTA, {time.time() - t0:.1f} s", flush=True)
            continue
        seen = set()
        for small in (False, True):           # the layout for < _jit.MIN_SERIES series too
            try:
                gen = _jit.generate(trie, iss.semiring._code, iss._weight_mode(),
                                    _jit.SieveSet.make(feats, bhi, bmm), jdims, shared,
                                    _jit.options(), n_shared,
                                    450 if getattr(iss, "_jit_only", False) else 0, small)
            except NotImplementedError as exc:
                print(f"{name} slice {si}: generic kernel ({exc})", flush=True)
                break
            if gen.digest() in seen:
                continue
            seen.add(gen.digest())
            t0 = time.time()
            cubin = _jit.build_cubin(gen)
            em = gen.em
            print(f"{name} slice {si}{' (small batches)' if small else ''}: {len(trie.nodes)} nodes, "
                  f"{len(gen.parts)} parts, {32 * em.ppc * em.gpc} threads/CTA, <= {gen.max_regs} "
                  f"registers, {len(cubin) // 1024} KB cubin, {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    for cfg in (sys.argv[1:] or list(SHAPES)):
        warm(cfg)
