"""Generator-option sweep for the plan-specialised kernel on the C5 pipeline.

    python scripts/jit_sweep.py --compile "budget=100,ppc=1" "budget=60,minb=2" ...   (no GPU: fills the cubin cache)
    python scripts/jit_sweep.py --run N "budget=100,ppc=1" ...                        (GPU: ms / series per second each)
    ... --config C4_twi                                                               (another configuration)
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _compile(opt: str, config: str = "C5_sweep"):
    os.environ["FRUITS_B200_JIT_OPTS"] = opt
    import jit_warm
    t0 = time.time()
    jit_warm.warm(config)
    return opt, time.time() - t0


def main() -> None:
    config = "C5_sweep"
    if "--config" in sys.argv:
        i = sys.argv.index("--config")
        config = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    mode = sys.argv[1]
    if mode == "--compile":
        opts = sys.argv[2:]
        with ProcessPoolExecutor(max_workers=min(8, len(opts))) as ex:
            for opt, dt in ex.map(_compile, opts, [config] * len(opts)):
                print(f"compiled [{opt}] in {dt:.1f} s", flush=True)
        return
    n = int(sys.argv[2])
    opts = sys.argv[3:]
    import numpy as np
    import torch
    import fruits_b200 as fruits
    import specs
    shape = {"C5_sweep": (3, 1024), "C4_twi": (3, 2048), "C3_general": (6, 1024), "C3_cos": (6, 1024)}[config]
    X = torch.randn((n,) + shape, dtype=torch.float64, device="cuda",
                    generator=torch.Generator("cuda").manual_seed(1234))
    if config != "C5_sweep":
        X = X.cumsum(dim=2)
    first = None
    for opt in opts:
        os.environ["FRUITS_B200_JIT_OPTS"] = opt
        if opt == "generic":
            os.environ["FRUITS_B200_JIT"] = "0"
            os.environ["FRUITS_B200_JIT_OPTS"] = ""
        else:
            os.environ["FRUITS_B200_JIT"] = "1"
        fruit = specs.build_fruit(fruits, specs.SPECS[config])
        np.random.seed(0)
        fruit.fit(X[:64])
        out = torch.empty((n, fruit.nfeatures()), dtype=torch.float64, device="cuda")
        fruit.transform_device(X, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        reps = 3
        for _ in range(reps):
            fruit.transform_device(X, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        if first is None:
            first = out.clone()
            same = True
        else:
            same = bool(torch.equal(first, out))
        print(f"[{opt}] {ms:8.2f} ms  {n / ms * 1e3 / 1e6:6.3f} M series/s  same_as_first={same}",
              flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    main()
