"""Generator options of the thread-per-series kernel on one slice (development aid):
    python scripts/opts_sweep.py C3_cos 1 "budget=150,ppc=1,gpc=8,unroll=1,abi=1" "budget=70,ppc=2" ...
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
    X = torch.from_numpy(specs.make_input(name)).cuda()
    one = {"slices": [dict(specs.SPECS[name]["slices"][si], fit_sample_size=1)]}
    ref = None
    for opts in sys.argv[3:]:
        os.environ["FRUITS_B200_JIT_OPTS"] = opts
        fruit = specs.build_fruit(fruits, one)
        np.random.seed(0)
        fruit.fit(X)
        try:
            out = fruit.transform_device(X)
        except Exception as exc:          # noqa: BLE001
            print(f"{opts}: {type(exc).__name__}: {exc}", flush=True)
            continue
        torch.cuda.synchronize()
        # device time without the host in the way: replay one captured call
        reps, run = 5, (lambda: fruit.transform_device(X, out=out))
        if os.environ.get("SWEEP_GRAPH", "1") == "1":
            try:
                g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    run()
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g, stream=st):
                        run()
                torch.cuda.synchronize()
                reps, run = 20, g.replay
            except Exception as exc:      # noqa: BLE001
                print(f"(no graph: {type(exc).__name__})", flush=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            run()
        ev[1].record()
        torch.cuda.synchronize()
        route, _, kern = fruit.get_slice(0)._last_launch
        em = getattr(kern, "em", None)
        shape = f"parts={len(em.p.parts)} ppc={em.ppc} gpc={em.gpc} tt={em.tt}" if em is not None and hasattr(em, "ppc") else ""
        same = "" if ref is None else f" equal_to_first={bool(torch.equal(out, ref))}"
        if ref is None:
            ref = out.clone()
        print(f"{name} slice {si} [{opts}]: {ev[0].elapsed_time(ev[1]) / reps:8.3f} ms route={route} {shape}{same}",
              flush=True)
