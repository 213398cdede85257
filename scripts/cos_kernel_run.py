"""One CosWISS materialisation of four-letter words (development aid for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import fruits_b200 as fruits  # noqa: E402
import specs  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    X = torch.from_numpy(specs.make_input("C3_cos", n)).cuda()
    words = [fruits.words.SimpleWord("[1][2][1][2]"), fruits.words.SimpleWord("[1][1][2]")]
    iss = fruits.CosWISS(words, freqs=[i / 20 for i in range(1, 11, 2)], exponent=2,
                         total_weighting=True)
    for _ in range(2):
        out = iss.materialize(X[:, :2].contiguous())
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    out = iss.materialize(X[:, :2].contiguous())
    ev[1].record()
    torch.cuda.synchronize()
    print(f"materialise {len(words)} words x 5 freqs, {n} series: {ev[0].elapsed_time(ev[1]):.2f} ms")
